"""The bootstrap all-gather of the multi-GPU exchange (shinestacker_b200/rendezvous.py), world_size 2 and 3 on the CPU:
separate processes, TCP on 127.0.0.1, several rounds, payloads of different sizes."""
import multiprocessing as mp
import os
import socket

import pytest

from shinestacker_b200 import rendezvous


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    try:
        rdv = rendezvous.TcpRendezvous(rank, world, "127.0.0.1", port, timeout=30)
        first = rdv.all_gather(bytes([rank]) * (384 + rank))
        rdv.barrier()
        second = rdv.all_gather(f"round2-{rank}".encode())
        rdv.close()
        q.put((rank, [len(p) for p in first], [p[:1] for p in first], second))
    except Exception as e:  # pylint: disable=broad-except
        q.put((rank, "error", repr(e), None))


@pytest.mark.parametrize("world", [2, 3])
def test_all_gather_between_processes(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = sorted(q.get(timeout=60) for _ in range(world))
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    for rank, lens, heads, second in results:
        assert lens == [384 + r for r in range(world)], (rank, lens, heads)
        assert heads == [bytes([r]) for r in range(world)]
        assert second == [f"round2-{r}".encode() for r in range(world)]


def test_from_env_single_rank(monkeypatch):
    monkeypatch.setenv("WORLD_SIZE", "1")
    rdv = rendezvous.TcpRendezvous.from_env()
    assert rdv.all_gather(b"x") == [b"x"] and rdv.world == 1
    rdv.barrier()
