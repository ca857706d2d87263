"""world_size-2 gloo test (CPU) of the multi-GPU exchange protocol in shinestacker_b200/sharded.py:
frames sharded into contiguous blocks, ONE max all-reduce over packed (value, global index) keys, losers
zeroed, ONE float32 sum-reduce of the coefficients to the root.  The per-rank pyramid state comes from the
CPU oracle here (on the GPU box the same protocol runs on the library's device buffers over NCCL); the
merged result must be bit-identical to fusing the whole stack in one go."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, n, h, w, seed, out_path):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import pyramid_oracle as po
    from shinestacker_b200 import synthetic
    from shinestacker_b200.sharded import shard_range, pack_keys, unpack_keys, exchange
    frames = synthetic.make_noise_stack(n, h, w, np.uint8, seed)
    # engineered exact ties across the shard boundary: frames 1 and 2 share a patch
    frames[2][:40, :40] = frames[1][:40, :40]
    f0, f1 = shard_range(n, rank, world)
    st = po.StreamingStack()
    for f in frames[f0:f1]:
        st.push(f)
    # local state with GLOBAL frame indices
    vals = [s["e"].ravel() for s in st.state] + [st.base["ve"].ravel(), st.base["vd"].ravel()]
    idxs = [s["idx"].ravel() + f0 for s in st.state] + [st.base["ie"].ravel() + f0, st.base["id"].ravel() + f0]
    coef = [s["fused"].reshape(3, -1) for s in st.state] + [st.base["pe"].reshape(3, -1), st.base["pd"].reshape(3, -1)]
    keys = torch.from_numpy(np.concatenate([pack_keys(v, i) for v, i in zip(vals, idxs)]))
    mine_before = np.concatenate(idxs)

    def mask():
        _, win = unpack_keys(keys.numpy())
        lost = win != mine_before
        off, parts = 0, []
        for c in coef:
            m = c.shape[1]
            c = c.copy()
            c[:, lost[off:off + m]] = 0.0
            parts.append(c.ravel())
            off += m
        return torch.from_numpy(np.concatenate(parts))

    coeffs = exchange(keys, mask, root=0)
    if rank == 0:
        val, win = unpack_keys(keys.numpy())
        off = coff = 0
        c = coeffs.numpy()
        for s in st.state:
            m = s["e"].size
            s["idx"][...] = win[off:off + m].reshape(s["idx"].shape)
            s["fused"][...] = c[coff:coff + 3 * m].reshape(s["fused"].shape)
            off += m
            coff += 3 * m
        m = st.base["ve"].size
        st.base["ie"][...] = win[off:off + m].reshape(st.base["ie"].shape)
        st.base["pe"][...] = c[coff:coff + 3 * m].reshape(st.base["pe"].shape)
        off += m
        coff += 3 * m
        st.base["id"][...] = win[off:off + m].reshape(st.base["id"].shape)
        st.base["pd"][...] = c[coff:coff + 3 * m].reshape(st.base["pd"].shape)
        out, info = st.finish()
        np.savez(out_path, out=out, best0=info["best"][0], best_e=info["best_e"], best_d=info["best_d"])
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n,world", [(5, 2), (4, 2)])
def test_sharded_exchange_matches_single_stack(tmp_path, oracle, n, world):
    from shinestacker_b200 import synthetic
    h, w, seed = 140, 200, 91
    out_path = str(tmp_path / "merged.npz")
    port = 29500 + (os.getpid() % 1000)
    mp.spawn(_worker, args=(world, port, n, h, w, seed, out_path), nprocs=world, join=True)
    frames = synthetic.make_noise_stack(n, h, w, np.uint8, seed)
    frames[2][:40, :40] = frames[1][:40, :40]
    ref, info = oracle.focus_stack(frames)
    z = np.load(out_path)
    assert np.array_equal(z["out"], ref)
    assert np.array_equal(z["best0"], info["best"][0])
    assert np.array_equal(z["best_e"], info["best_e"])
    assert np.array_equal(z["best_d"], info["best_d"])


def test_shard_range_and_key_packing():
    from shinestacker_b200.sharded import shard_range, pack_keys, unpack_keys
    for n in (1, 7, 8, 64, 255):
        for world in (1, 2, 3, 8):
            blocks = [shard_range(n, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(blocks, blocks[1:]))
    v = np.array([0.0, -0.0, 1.5, 1.5, 3e-39, 7.0], np.float32)
    i = np.array([3, 2, 9, 4, 0, 65535], np.int32)
    k = pack_keys(v, i)
    assert k[0] < k[1]                      # equal values (-0 == +0): the lower index wins the max
    assert k[3] > k[2]                      # tie on 1.5: index 4 beats index 9
    val, idx = unpack_keys(k)
    assert np.array_equal(idx, i) and np.array_equal(val, np.abs(v))
