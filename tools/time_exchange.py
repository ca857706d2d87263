"""Phase timing of the multi-GPU finish() (run under torchrun)."""
import ctypes, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from shinestacker_b200 import synthetic, _native
from shinestacker_b200.sharded import ShardedPyramidStack, shard_range, _DevArray, exchange
from shinestacker_b200.pyramid import _check, select_base

world = int(os.environ["WORLD_SIZE"]); rank = int(os.environ["RANK"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
lib = _native.load()
N, H, W = 64, 3072, 4096
f0, f1 = shard_range(N, rank, world); n_local = f1 - f0
sharp, blurred = synthetic.make_scene(H, W, np.uint8, 2)
sharp = np.clip(sharp, 0, 255).astype(np.uint8); blurred = np.clip(blurred, 0, 255).astype(np.uint8)
res = ShardedPyramidStack(rank, world, device=local, batch_frames=n_local)
frames = [synthetic.frame_from_scene(sharp, blurred, f0 + i, N, np.uint8) for i in range(n_local)]
algo = res.begin(frames[0], n_local, N, resident=True)
for f in frames:
    _check(lib.ss_push_frame(algo._h, f.ctypes.data_as(ctypes.c_void_p), W * 3))
_check(lib.ss_sync(algo._h))
out = np.empty((H, W, 3), np.uint8)
def sync():
    torch.cuda.synchronize(local)
T = {}
def tick(name, t0):
    sync(); T[name] = T.get(name, 0.0) + time.perf_counter() - t0
for it in range(8):
    if it == 3: T.clear()
    dist.barrier(); sync()
    t = time.perf_counter(); _check(lib.ss_rerun_resident(algo._h)); _check(lib.ss_sync(algo._h)); tick("chain", t)
    h = algo._h
    t = time.perf_counter(); select_base(lib, h, n_local, 256); tick("base", t)
    t = time.perf_counter(); ptr, cnt = ctypes.c_void_p(), ctypes.c_size_t(); _check(lib.ss_exchange_keys(h, ctypes.byref(ptr), ctypes.byref(cnt)))
    keys = torch.as_tensor(_DevArray(ptr.value, cnt.value, "<i8"), device=f"cuda:{local}"); tick("pack", t)
    t = time.perf_counter(); dist.all_reduce(keys, op=dist.ReduceOp.MAX); tick("allreduce_keys", t)
    t = time.perf_counter(); p2, c2 = ctypes.c_void_p(), ctypes.c_size_t(); _check(lib.ss_exchange_mask_losers(h, ctypes.byref(p2), ctypes.byref(c2)))
    co = torch.as_tensor(_DevArray(p2.value, c2.value, "<f4"), device=f"cuda:{local}"); tick("mask", t)
    t = time.perf_counter(); dist.reduce(co, dst=0, op=dist.ReduceOp.SUM); tick("reduce_coeffs", t)
    if rank == 0:
        t = time.perf_counter(); _check(lib.ss_exchange_adopt(h)); tick("adopt", t)
        t = time.perf_counter(); _check(lib.ss_finish_collapse(h, out.ctypes.data_as(ctypes.c_void_p), out.strides[0])); tick("collapse+d2h", t)
if rank == 0:
    print("world", world, {k: round(v / 5 * 1e3, 3) for k, v in T.items()}, "ms; keys MB", cnt.value * 8 / 1e6, "coeff MB", c2.value * 4 / 1e6)
dist.destroy_process_group()
