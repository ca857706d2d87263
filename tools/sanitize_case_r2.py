"""Round-2 code paths for compute-sanitizer (memcheck / racecheck / synccheck): the level kernel with tensor-memory
parking + owner-form S0 + energy skip on interior tiles (uint8, uint16 and the float32 levels), the fused peer exchange
(three handles on one device, phase by phase) and the sub-action passes."""
import ctypes, os, sys
import numpy as np
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from shinestacker_b200 import PyramidStack, synthetic, subactions, _native
from shinestacker_b200.pyramid import _check
from shinestacker_b200.sharded import shard_range
for gen, dtype, shape in ((synthetic.make_stack, np.uint8, (4, 288, 512)), (synthetic.make_noise_stack, np.uint16, (3, 224, 320)),
                          (synthetic.make_stack, np.uint8, (3, 300, 407))):
    frames = gen(shape[0], shape[1], shape[2], dtype, 7)
    out = PyramidStack(batch_frames=2).focus_stack_arrays(frames)
    print(dtype.__name__, shape, int(out.astype(np.int64).sum()))
lib = _native.load()
frames = synthetic.make_noise_stack(6, 200, 328, np.uint8, 9)
world, algos = 3, []
for r in range(world):
    f0, f1 = shard_range(len(frames), r, world)
    a = PyramidStack(frame_index_base=f0, keep_depth_maps=False, batch_frames=2)
    a._set_metadata(frames[0]); a._ensure_handle(frames[0].shape, frames[0].dtype, f1 - f0)
    algos.append((a, f0, f1))
ptrs = (ctypes.c_void_p * (6 * world))()
for r, (a, _, _) in enumerate(algos):
    mine = (ctypes.c_void_p * 6)(); _check(lib.ss_peer_ptrs(a._h, mine))
    for i in range(6): ptrs[6 * r + i] = mine[i]
for r, (a, _, _) in enumerate(algos): _check(lib.ss_peer_connect_ptrs(a._h, ptrs, world, r, 0))
for rep in range(2):
    for a, f0, f1 in algos:
        _check(lib.ss_reset(a._h)); ring = a._staging(frames[0].nbytes, 4)
        for i, f in enumerate(frames[f0:f1]): a._push(i, f, ring)
    for ph in (1, 2, 3):
        for a, _, _ in algos: _check(lib.ss_peer_phase(a._h, ph))
    out = np.empty(frames[0].shape, np.uint8)
    for r, (a, _, _) in enumerate(algos):
        _check(lib.ss_peer_wait(a._h, out.ctypes.data_as(ctypes.c_void_p) if r == 0 else None, out.strides[0] if r == 0 else 0))
print("peer", int(out.astype(np.int64).sum()), bool(np.array_equal(out, PyramidStack().focus_stack_arrays(frames))))
img = np.random.default_rng(3).integers(0, 256, (97, 131, 3)).astype(np.uint8)
f = subactions.DeviceFrame(img.shape[:2], np.uint8)
o = f.upload(img).apply_lut(np.random.default_rng(4).integers(0, 256, (3, 256)).astype(np.uint8)).apply_vignette((300.0, 0.02, 60.0), 150.0, 0.8, 2.0).download()
print("sub", int(o.astype(np.int64).sum()))
