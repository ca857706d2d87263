"""Parity tests proper: the CUDA path (through the ctypes / C-ABI boundary) against the golden vectors
produced by the live reference, against the CPU oracle on seeded inputs, and through size-independent
properties at larger sizes.  Bit-exact is the bar for the integer arg-max maps AND the fused image
(the north-star tolerates 1 LSB on the image; we do not use that slack)."""
import ctypes
import os

import numpy as np
import pytest

import golden_util
from shinestacker_b200 import PyramidStack, synthetic, _native
from shinestacker_b200 import ShapeError, BitDepthError, RunStopException

pytestmark = pytest.mark.gpu


def run_gpu(frames, batch_frames=8, **kwargs):
    algo = PyramidStack(batch_frames=batch_frames, **kwargs)
    out = algo.focus_stack_arrays(frames)
    return out, algo.depth_maps, algo


@pytest.mark.parametrize("name", golden_util.case_names())
def test_cuda_matches_golden(name):
    frames, kwargs, expect = golden_util.load_case(name)
    out, maps, _ = run_gpu(frames, batch_frames=3, **kwargs)
    golden_util.assert_matches(out, maps, expect, name)


@pytest.mark.parametrize("gen,n,h,w,dtype,seed,batch", [
    ("stack", 7, 333, 517, np.uint8, 31, 2),
    ("noise", 5, 129, 1000, np.uint8, 32, 8),
    ("noise", 4, 500, 750, np.uint16, 33, 3),      # tail columns at 750, 375, 188
    ("stack", 9, 64, 64, np.uint16, 34, 4),
    ("noise", 3, 67, 93, np.uint8, 35, 1),
    ("stack", 16, 768, 1024, np.uint8, 36, 8),
    ("noise", 2, 1025, 1537, np.uint16, 37, 2),
])
def test_cuda_matches_oracle(oracle, gen, n, h, w, dtype, seed, batch):
    f = synthetic.make_stack if gen == "stack" else synthetic.make_noise_stack
    frames = f(n, h, w, dtype, seed)
    ref, info = oracle.focus_stack(frames)
    out, maps, algo = run_gpu(frames, batch_frames=batch)
    golden_util.assert_matches(out, maps, dict(fused=ref, best=info["best"], best_e=info["best_e"],
                                               best_d=info["best_d"], levels=info["levels"]), "oracle")
    for a, b in zip(algo.get_fused_pyramid(), info["fused_pyramid"]):
        assert np.array_equal(a, b)


@pytest.mark.parametrize("n,h,w,dtype,seed,kwargs", [
    (3, 150, 200, np.uint8, 61, {"kernel_size": 13}),         # 169 values per window: one halving of numpy's pairwise sum
    (3, 300, 260, np.uint16, 62, {"kernel_size": 20}),        # 361
    (2, 280, 300, np.uint8, 63, {"kernel_size": 32}),         # 961: the widest window (side 31)
    (2, 150, 200, np.uint8, 64, {"kernel_size": 17, "float_type": "float-64"}),
    (2, 150, 200, np.uint8, 65, {"kernel_size": 12}),         # the widest window of the small instantiation
])
def test_wide_base_windows(oracle, n, h, w, dtype, seed, kwargs):
    frames = synthetic.make_noise_stack(n, h, w, dtype, seed)
    ref, info = oracle.focus_stack(frames, **kwargs)
    out, maps, _ = run_gpu(frames, **kwargs)
    assert np.array_equal(out, ref)
    assert np.array_equal(maps["best_e"], info["best_e"]) and np.array_equal(maps["best_d"], info["best_d"])


def test_window_wider_than_the_oracle_range_is_rejected():
    from shinestacker_b200 import InvalidOptionError
    with pytest.raises(InvalidOptionError):
        run_gpu(synthetic.make_noise_stack(2, 100, 100, np.uint8, 66), kernel_size=34)


def test_batch_size_does_not_change_the_result():
    frames = synthetic.make_noise_stack(11, 200, 264, np.uint8, 41)
    base, maps0, _ = run_gpu(frames, batch_frames=11)
    for b in (1, 2, 5, 8):
        out, maps, _ = run_gpu(frames, batch_frames=b)
        assert np.array_equal(out, base)
        for x, y in zip(maps["best"], maps0["best"]):
            assert np.array_equal(x, y)


def test_first_index_wins_on_exact_ties():
    # identical frames: every energy ties, np.argmax returns 0 everywhere (reference pyramid.py:51)
    f = synthetic.make_noise_stack(1, 160, 224, np.uint8, 42)[0]
    out, maps, _ = run_gpu([f, f.copy(), f.copy()])
    for m in maps["best"] + [maps["best_e"], maps["best_d"]]:
        assert not m.any()
    single, _, _ = run_gpu([f])
    assert np.array_equal(out, single)


def test_single_frame_reconstructs_the_input_within_rounding():
    f = synthetic.make_noise_stack(1, 256, 320, np.uint16, 43)[0]
    out, _, _ = run_gpu([f])
    assert np.abs(out.astype(np.int64) - f.astype(np.int64)).max() <= 1


def test_handle_reuse_like_focus_stack_bunch(oracle):
    # FocusStackBunch calls focus_stack repeatedly on one stack_algo (reference stack.py:94-97)
    algo = PyramidStack(batch_frames=2)
    for seed, n in ((51, 4), (52, 3), (53, 4)):
        frames = synthetic.make_stack(n, 150, 210, np.uint8, seed)
        ref, info = oracle.focus_stack(frames)
        assert np.array_equal(algo.focus_stack_arrays(frames), ref)
        assert np.array_equal(algo.depth_maps["best"][0], info["best"][0])


def test_file_interface_errors_and_callbacks(tmp_path, oracle):
    import cv2
    frames = synthetic.make_stack(4, 130, 170, np.uint8, 61)
    paths = []
    for i, f in enumerate(frames):
        p = str(tmp_path / f"f{i:02d}.tif")
        cv2.imwrite(p, f)
        paths.append(p)

    class Proc:
        id, name = 7, "stack"

        def __init__(self, stop_at=None):
            self.steps, self.stop_at, self.msgs = [], stop_at, []

        def sub_message_r(self, m):
            self.msgs.append(m)

        def callback(self, key, *a):
            if key == "after_step":
                self.steps.append(a[-1])
            if key == "check_running" and self.stop_at is not None and len(self.steps) >= self.stop_at:
                return False
            return None

    algo = PyramidStack()
    algo.process = Proc()
    algo.do_step_callback = True
    out = algo.focus_stack(paths)
    ref, _ = oracle.focus_stack(frames)
    assert np.array_equal(out, ref)
    assert algo.process.steps == list(range(8))                  # validation pass first, then processing (2 steps per frame)
    algo.process = Proc(stop_at=3)
    with pytest.raises(RunStopException):
        algo.focus_stack(paths)
    algo.process = Proc()
    assert np.array_equal(algo.focus_stack(paths), ref)           # usable after a cancelled run
    bad = str(tmp_path / "f99.tif")
    cv2.imwrite(bad, np.zeros((100, 170, 3), np.uint8))
    with pytest.raises(ShapeError):
        algo.focus_stack(paths + [bad])
    cv2.imwrite(bad, np.zeros((130, 170, 3), np.uint16))
    with pytest.raises(BitDepthError):
        algo.focus_stack(paths + [bad])


def test_collapse_kernel_is_expand_layer(oracle):
    lib = _native.load()
    rng = np.random.default_rng(71)
    k = oracle.make_kernel(0.4)
    for (h2, w2, h, w) in ((20, 30, 40, 60), (21, 47, 41, 93), (33, 94, 65, 188)):
        img = (rng.random((3, h2, w2), dtype=np.float32) * 255).astype(np.float32)
        lay = ((rng.random((3, h, w), dtype=np.float32) - .5) * 50).astype(np.float32)
        out = np.empty((3, h, w), np.float32)
        rc = lib.ss_test_collapse_step(img.ctypes.data_as(ctypes.c_void_p), h2, w2,
                                       lay.ctypes.data_as(ctypes.c_void_p), h, w, 0.4,
                                       out.ctypes.data_as(ctypes.c_void_p))
        assert rc == 0, _native.last_error()
        for c in range(3):
            assert np.array_equal(out[c], oracle.expand_plane(img[c], k, h, w) + lay[c])


@pytest.mark.parametrize("dtype,n", [(np.uint8, 12), (np.uint16, 6)])
def test_full_width_staircase_properties(dtype, n):
    """4096x3072 (BASELINE config 2/3 geometry, fewer frames): the level-0 depth map must be the
    staircase the generator built and the fused image must be the sharp scene."""
    h, w = 3072, 4096
    sharp, blurred = synthetic.make_scene(h, w, dtype, 81)
    frames = [synthetic.frame_from_scene(sharp, blurred, i, n, dtype) for i in range(n)]
    out, maps, _ = run_gpu(frames, batch_frames=4)
    d0 = maps["best"][0]
    for i in range(n):
        x0, x1 = (i * w) // n, ((i + 1) * w) // n
        band = d0[64:-64, x0 + 8:x1 - 8]
        assert (band == i).mean() > 0.999
    assert not d0[:14, :14].any()          # constant patch: exact ties -> frame 0
    # permutation check: the fused image is sharp everywhere => equals the sharp scene away from
    # band borders up to the pyramid's rounding
    maxv = np.iinfo(dtype).max
    ref = np.clip(sharp, 0, maxv).astype(dtype)
    x0, x1 = (3 * w) // n + 16, (4 * w) // n - 16
    err = np.abs(out[64:-64, x0:x1].astype(np.int64) - ref[64:-64, x0:x1].astype(np.int64))
    assert err.max() <= max(1, maxv // 128)   # coarse levels may come from another (blurred) frame


def _peer_stack_on_one_gpu(frames, world, root=0, **kw):
    """`world` native handles play the ranks of a box on one device: contiguous frame blocks, ss_peer_connect_ptrs
    (device pointers instead of CUDA IPC), every rank's part enqueued before any is awaited."""
    from shinestacker_b200.sharded import shard_range
    from shinestacker_b200.pyramid import _check
    lib = _native.load()
    n = len(frames)
    h, w = frames[0].shape[:2]
    algos = []
    for r in range(world):
        f0, f1 = shard_range(n, r, world)
        a = PyramidStack(device=0, frame_index_base=f0, keep_depth_maps=False, **kw)
        a._set_metadata(frames[0])
        a._ensure_handle(frames[0].shape, frames[0].dtype, f1 - f0)
        algos.append((a, f0, f1))
    ptrs = (ctypes.c_void_p * (6 * world))()
    for r, (a, _, _) in enumerate(algos):
        mine = (ctypes.c_void_p * 6)()
        _check(lib.ss_peer_ptrs(a._h, mine))
        for i in range(6):
            ptrs[6 * r + i] = mine[i]
    for r, (a, _, _) in enumerate(algos):
        _check(lib.ss_peer_connect_ptrs(a._h, ptrs, world, r, root))
    outs = []
    for rep in range(2):                        # twice: the second stack has to wait for the DONE flags of the first
        for a, f0, f1 in algos:
            _check(lib.ss_reset(a._h))
            ring = a._staging(frames[0].nbytes, 2 * max(a.batch_frames, 1))
            for i, f in enumerate(frames[f0:f1]):
                a._push(i, f, ring)
        for phase in (1, 2, 3):                 # phase by phase: the ranks share one device here
            for a, _, _ in algos:
                _check(lib.ss_peer_phase(a._h, phase))
        out = np.empty((h, w, 3), frames[0].dtype)
        for r, (a, _, _) in enumerate(algos):
            if r == root:
                _check(lib.ss_peer_wait(a._h, out.ctypes.data_as(ctypes.c_void_p), out.strides[0]))
            else:
                _check(lib.ss_peer_wait(a._h, None, 0))
        outs.append(out)
    maps = algos[root][0].get_depth_maps()
    for a, _, _ in algos:
        a.close()
    assert np.array_equal(outs[0], outs[1])
    return outs[1], maps


@pytest.mark.parametrize("world,n,h,w,dtype,seed", [
    (2, 7, 300, 420, np.uint8, 95),
    (3, 8, 260, 390, np.uint8, 96),
    (8, 19, 515, 750, np.uint16, 94),          # odd level sizes, bands of a few rows at the coarse levels
    (4, 5, 1030, 520, np.uint8, 93),           # more rows than columns; one rank gets a single frame
])
def test_fused_peer_exchange_on_one_gpu(oracle, world, n, h, w, dtype, seed):
    """Frames sharded over `world` ranks, merged by the fused peer exchange (ss_peer_*: epoch flags, banded merge out
    of the peers' state, banded collapse into the root's image): bit-identical to fusing the stack in one go."""
    frames = synthetic.make_noise_stack(n, h, w, dtype, seed)
    frames[4][:50, :60] = frames[2][:50, :60]          # exact ties across shard boundaries
    frames[n - 1][h - 40:, w - 70:] = frames[0][h - 40:, w - 70:]
    ref, info = oracle.focus_stack(frames)
    out, maps = _peer_stack_on_one_gpu(frames, world, batch_frames=2)
    golden_util.assert_matches(out, maps, dict(fused=ref, best=info["best"], best_e=info["best_e"],
                                               best_d=info["best_d"], levels=info["levels"]), f"peer world={world}")


def test_fused_peer_exchange_other_root(oracle):
    frames = synthetic.make_stack(6, 280, 384, np.uint8, 92)
    ref, _ = oracle.focus_stack(frames)
    out, _ = _peer_stack_on_one_gpu(frames, 3, root=2, batch_frames=3)
    assert np.array_equal(out, ref)


def test_config4_geometry_tail_columns_u16(oracle):
    """BASELINE config 4 geometry (6000x4000 RGB16: tail columns at 1500/750/375/188/94 and at the expand
    widths), few frames so that the oracle stays fast."""
    frames = synthetic.make_noise_stack(3, 4000, 6000, np.uint16, 97)
    ref, info = oracle.focus_stack(frames)
    out, maps, _ = run_gpu(frames, batch_frames=2)
    golden_util.assert_matches(out, maps, dict(fused=ref, best=info["best"], best_e=info["best_e"],
                                               best_d=info["best_d"], levels=info["levels"]), "config4")


def test_many_frames_u8_groups(oracle):
    """More frames than any frame-split group count, small image: exercises the group merge order."""
    frames = synthetic.make_noise_stack(70, 96, 160, np.uint8, 98)
    for i in (5, 17, 40, 69):
        frames[i] = frames[3].copy()          # exact ties far apart in the stack
    ref, info = oracle.focus_stack(frames)
    out, maps, _ = run_gpu(frames, batch_frames=70)
    golden_util.assert_matches(out, maps, dict(fused=ref, best=info["best"], best_e=info["best_e"],
                                               best_d=info["best_d"], levels=info["levels"]), "groups")


@pytest.mark.parametrize("gen,n,h,w,dtype,seed", [
    ("noise", 5, 300, 407, np.uint8, 101),          # tail columns in the float32 energy filter
    ("stack", 4, 515, 750, np.uint16, 102),
])
def test_float64_matches_oracle(oracle, gen, n, h, w, dtype, seed):
    """float_type='float-64' (reference base_stack_algo.py:15-16): float64 pyramids, float32 energy."""
    f = synthetic.make_stack if gen == "stack" else synthetic.make_noise_stack
    frames = f(n, h, w, dtype, seed)
    ref, info = oracle.focus_stack(frames, float_type="float-64")
    out, maps, _ = run_gpu(frames, float_type="float-64")
    golden_util.assert_matches(out, maps, dict(fused=ref, best=info["best"], best_e=info["best_e"],
                                               best_d=info["best_d"], levels=info["levels"]), "float-64")


def test_file_interface_maps_uncompressed_tiffs(tmp_path):
    # the files the reference's write_img produces (uncompressed TIFF) take the zero-decode path (fastio)
    import cv2
    from shinestacker_b200 import fastio
    frames = synthetic.make_stack(6, 200, 264, np.uint16, 51)
    paths = []
    for i, f in enumerate(frames):
        p = str(tmp_path / f"{i:03d}.tif")
        cv2.imwrite(p, f, [int(cv2.IMWRITE_TIFF_COMPRESSION), 1])
        assert fastio.open_plain_tiff(p) is not None
        paths.append(p)
    expect = PyramidStack().focus_stack_arrays(frames)
    for threads in (0, 4):
        assert np.array_equal(PyramidStack(decode_threads=threads, batch_frames=4).focus_stack(paths), expect)
