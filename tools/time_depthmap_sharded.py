#!/usr/bin/env python
"""DepthMapStack with the frames of one stack sharded over the GPUs of a box: one process per GPU, launched like
bench.py (torchrun-style RANK / LOCAL_RANK / WORLD_SIZE / MASTER_ADDR / MASTER_PORT in the environment):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \\
        tools/time_depthmap_sharded.py --frames 64

Every rank generates the same synthetic stack and keeps its own run of the frames in pinned memory; the timed region is
upload + kernels + exchange + download on the root, bracketed by barriers.  The root compares its image with the one
a single handle produces from all frames (within 1 LSB: the sums over the stack are associated rank by rank)."""
import argparse
import ctypes
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=64)
    ap.add_argument("--height", type=int, default=3072)
    ap.add_argument("--width", type=int, default=4096)
    ap.add_argument("--dtype", default="uint8")
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--check", action="store_true", help="root also runs the whole stack on one handle and compares")
    a = ap.parse_args()
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", str(rank)))
    from shinestacker_b200 import DepthMapStack, synthetic, _native
    from shinestacker_b200 import rendezvous as rdvm
    from shinestacker_b200.sharded import ShardedDepthMapStack, shard_range
    from shinestacker_b200.pyramid import _check
    from shinestacker_b200.sharded import bind_to_gpu_numa_node
    bind_to_gpu_numa_node(local)                # pinned buffers and upload traffic on the GPU's own socket
    lib = _native.load()
    rdv = rdvm.TcpRendezvous.from_env(offset=7) if world > 1 else rdvm.Single()
    dt = np.dtype(a.dtype)
    frames = synthetic.make_stack(a.frames, a.height, a.width, dt, seed=2)
    f0, f1 = shard_range(a.frames, rank, world)
    n_local = f1 - f0
    algo = ShardedDepthMapStack(rank, world, device=local, rendezvous=rdv)
    dm = algo.algo
    dm._ensure_handle(frames[0].shape, dt, n_local)
    pin = ctypes.c_void_p()
    _check(lib.ss_host_alloc(ctypes.byref(pin), n_local * frames[0].nbytes))
    host = np.frombuffer((ctypes.c_uint8 * (n_local * frames[0].nbytes)).from_address(pin.value), dt).reshape(
        (n_local,) + frames[0].shape)
    host[:] = frames[f0:f1]
    pitch = a.width * 3 * dt.itemsize
    out, out_pin = None, ctypes.c_void_p()
    if rank == 0:                               # the root's result lands in pinned memory
        _check(lib.ss_host_alloc(ctypes.byref(out_pin), frames[0].nbytes))
        out_buf = np.frombuffer((ctypes.c_uint8 * frames[0].nbytes).from_address(out_pin.value), dt).reshape(frames[0].shape)
    else:
        out_buf = None

    def step():
        dm._ensure_handle(frames[0].shape, dt, n_local)        # resets the handle
        for i in range(n_local):
            _check(lib.ssdm_push_frame(dm._h, host[i].ctypes.data_as(ctypes.c_void_p), pitch))
        return algo.finish(n_local, frames[0].shape[:2], dt, out=out_buf)

    step()
    times, dev = [], []
    for _ in range(a.steps):
        rdv.barrier()
        t0 = time.perf_counter()
        out = step()
        rdv.barrier()
        times.append(time.perf_counter() - t0)
        dev.append(algo.last_kernel_ms)
    all_t = [np.frombuffer(b, np.float64)[0] for b in rdv.all_gather(np.float64(np.median(times)).tobytes())]
    all_d = [np.frombuffer(b, np.float64)[0] for b in rdv.all_gather(np.float64(np.median(dev)).tobytes())]
    if rank == 0:
        t = max(all_t)
        line = {"workload": f"DepthMapStack {a.frames} x {a.width}x{a.height} {a.dtype}, frames sharded over {world} GPU(s), "
                            "pinned host frames -> fused image on the root",
                "n_gpus": world, "e2e_ms_per_stack": t * 1e3, "e2e_mp_per_s": a.frames * a.height * a.width / 1e6 / t,
                "device_ms_first_push_to_end_per_rank": [round(v, 3) for v in all_d]}
        if a.check:
            ref = DepthMapStack(device=local).focus_stack_arrays(list(frames))
            d = np.abs(ref.astype(np.int64) - out.astype(np.int64))
            line["vs_one_handle"] = {"max_abs_diff_lsb": int(d.max()), "fraction_differing": float((d > 0).mean())}
        print(json.dumps(line))
    rdv.barrier()
    algo.close()
    lib.ss_host_free(pin)
    if out_pin.value:
        lib.ss_host_free(out_pin)
    rdv.close()


if __name__ == "__main__":
    main()
