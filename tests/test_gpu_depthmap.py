"""GPU parity tests of the depth-map path (through ctypes -> C ABI, include/shinestack_b200_depthmap.h):
the CUDA kernels against the golden vectors of the live reference and against the numpy oracle.
DepthMapStack blends with soft weights, so the contract is the fused image within 1 LSB;
the focus energy of 8-bit stacks is integer-valued in float32 and must be bit-exact."""

import numpy as np
import pytest

import depthmap_util as du
from oracle import depthmap_oracle as dmo
from shinestacker_b200 import DepthMapStack, synthetic, ShapeError, RunStopException

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", du.case_names())
def test_cuda_matches_golden(name):
    frames, kwargs, expect = du.load_case(name)
    algo = DepthMapStack(**kwargs)
    out = algo.focus_stack_arrays(frames)
    du.assert_within_1lsb(out, expect["fused"], name)
    assert np.abs(algo.get_plane("weights", 0) - expect["weights0"]).max() < 2e-3   # float16 fixture
    assert algo.last_launches > 0 and algo.last_kernel_ms > 0


@pytest.mark.parametrize("n,h,w,dtype,seed,kwargs", [
    (5, 333, 517, np.uint8, 71, {}),
    (3, 257, 1000, np.uint16, 72, {"map_type": "max"}),
    (4, 129, 191, np.uint8, 73, {"energy": "sobel", "levels": 5}),
    (3, 600, 800, np.uint8, 74, {"smooth_size": 0, "levels": 4}),
    (2, 1025, 1537, np.uint16, 75, {"kernel_size": 7, "blur_size": 7, "smooth_size": 9}),
    (12, 96, 136, np.uint8, 76, {"map_type": "max", "temperature": 0.03}),
    (2, 150, 260, np.uint8, 77, {"kernel_size": 9, "blur_size": 15, "smooth_size": 5}),     # halo 11: > 48 KB of shared memory
    (2, 140, 200, np.uint16, 78, {"kernel_size": 31, "blur_size": 31, "smooth_size": 0, "levels": 2}),   # the widest kernels
])
def test_cuda_matches_oracle(n, h, w, dtype, seed, kwargs):
    frames = synthetic.make_noise_stack(n, h, w, np.dtype(dtype), seed)
    algo = DepthMapStack(**kwargs)
    out = algo.focus_stack_arrays(frames)
    ref, st = dmo.focus_stack(frames, return_stages=True, **kwargs)
    du.assert_within_1lsb(out, ref, "oracle")
    wts = np.array([algo.get_plane("weights", i) for i in range(n)])
    assert np.abs(wts - st["weights"]).max() < 5e-6
    assert np.abs(wts.sum(axis=0) - 1).max() < 1e-5


def test_energy_stage_is_exact_and_matches_the_host_build_of_the_kernels():
    frames = synthetic.make_noise_stack(3, 150, 203, np.uint8, 77)
    algo = DepthMapStack()
    algo._ensure_handle(frames[0].shape, frames[0].dtype, 3)
    ring = algo._staging(frames[0].nbytes, 4)
    for i, f in enumerate(frames):
        algo._push(i, f, ring)
    raw = np.array([algo.get_plane("energy", i) for i in range(3)])
    assert np.array_equal(raw, dmo.energy_maps(frames))
    out = algo._blend_and_finish(3, frames[0].shape[:2], frames[0].dtype)
    st = {}
    emu = du.emu_focus_stack(frames, stages=st)
    # same functors, same order of operations: the host build and the device must agree bit for bit
    assert np.array_equal(out, emu)
    assert np.array_equal(np.array([algo.get_plane("weights", i) for i in range(3)]), st["weights"])


def test_handle_reuse_and_rerun():
    a = synthetic.make_noise_stack(4, 120, 160, np.uint8, 78)
    b = synthetic.make_noise_stack(3, 120, 160, np.uint8, 79)
    algo = DepthMapStack()
    out_a = algo.focus_stack_arrays(a)
    out_b = algo.focus_stack_arrays(b)          # FocusStackBunch reuses one stack_algo (stack.py:94-97)
    assert np.array_equal(algo.focus_stack_arrays(a), out_a)
    assert np.array_equal(DepthMapStack().focus_stack_arrays(b), out_b)
    assert np.array_equal(DepthMapStack().focus_stack_arrays(b, stage=False), out_b)    # upload without staging
    import ctypes
    assert algo._lib.ssdm_rerun_resident(algo._h) == 0
    again = np.empty_like(out_a)
    assert algo._lib.ssdm_finish_end(algo._h, again.ctypes.data_as(ctypes.c_void_p), again.strides[0]) == 0
    assert np.array_equal(again, out_a)


def test_file_interface_callbacks_and_errors(tmp_path):
    import cv2
    from unittest.mock import MagicMock
    frames = synthetic.make_noise_stack(4, 100, 140, np.uint16, 80)
    paths = []
    for i, f in enumerate(frames):
        p = str(tmp_path / f"{i:03d}.tif")
        cv2.imwrite(p, f)
        paths.append(p)
    algo = DepthMapStack()
    algo.process = MagicMock()
    algo.process.callback.return_value = True
    out = algo.focus_stack(paths)
    assert np.array_equal(out, DepthMapStack().focus_stack_arrays(frames))
    steps = [c.args[3] for c in algo.process.callback.call_args_list if c.args[0] == "after_step"]
    assert steps == list(range(8))              # depth_map.py:75 and :113: i, then i + n
    # cancellation between frames
    algo.process.callback.side_effect = lambda what, *a: False if what == "check_running" else None
    with pytest.raises(RunStopException):
        algo.focus_stack(paths)
    algo.process.callback.side_effect = None
    cv2.imwrite(paths[2], frames[2][:50])
    with pytest.raises(ShapeError):
        algo.focus_stack(paths)
    cv2.imwrite(paths[2], frames[2])
    assert np.array_equal(algo.focus_stack(paths), out)   # the handle survives the failures


def test_full_size_identity_property():
    # 4096 x 3072: a stack of identical frames fuses to that frame (weights sum to 1, the pyramid telescopes)
    img = synthetic.make_noise_stack(1, 3072, 4096, np.uint8, 81)[0]
    out = DepthMapStack().focus_stack_arrays([img] * 3)
    du.assert_within_1lsb(out, img, "identical frames", max_frac=0.5)


def _sharded_on_one_gpu(frames, world, root, **kwargs):
    """The sharded call sequence with `world` handles on ONE device in one process (plain device pointers instead of
    IPC handles) -- the kernels and the order of the phases are those of a multi-GPU run."""
    import ctypes
    from shinestacker_b200.sharded import shard_range
    from shinestacker_b200.pyramid import _check
    algos = []
    for r in range(world):
        a, b = shard_range(len(frames), r, world)
        algo = DepthMapStack(**kwargs)
        algo._ensure_handle(frames[0].shape, frames[0].dtype, b - a)
        ring = algo._staging(frames[0].nbytes, 4)
        for i, f in enumerate(frames[a:b]):
            algo._push(i, f, ring)
        algos.append((algo, b - a))
    lib = algos[0][0]._lib
    maxima = []
    for algo, _ in algos:
        m = ctypes.c_float()
        _check(lib.ssdm_shard_local_max(algo._h, ctypes.byref(m)))
        maxima.append(m.value)
    count = lib.ssdm_shard_buffer_count(algos[0][0]._h)
    table = (ctypes.c_void_p * (world * count))()
    for r, (algo, _) in enumerate(algos):
        own = (ctypes.c_void_p * count)()
        _check(lib.ssdm_shard_ptrs(algo._h, own))
        table[r * count:(r + 1) * count] = list(own)
    for r, (algo, _) in enumerate(algos):
        _check(lib.ssdm_shard_connect_ptrs(algo._h, table, world, r, root))
        _check(lib.ssdm_shard_finish_local(algo._h, ctypes.c_float(max(maxima))))
    for algo, _ in algos:
        _check(lib.ssdm_sync(algo._h))
    for algo, _ in algos:
        _check(lib.ssdm_shard_softmax_rel(algo._h))
        _check(lib.ssdm_sync(algo._h))
    for algo, _ in algos:
        _check(lib.ssdm_shard_finish_weights(algo._h))
    weights = np.array([algo.get_plane("weights", i) for algo, n in algos for i in range(n)])
    for algo, n in algos:
        for i in range(n):
            _check(lib.ssdm_blend_frame(algo._h, i))
        _check(lib.ssdm_sync(algo._h))
    for algo, _ in algos:
        _check(lib.ssdm_shard_reduce_band(algo._h))
    for algo, _ in algos:
        _check(lib.ssdm_sync(algo._h))
    out = np.empty_like(frames[0])
    for r, (algo, _) in enumerate(algos):
        if r != root:
            _check(lib.ssdm_shard_finish_end(algo._h, None, 0))
    _check(lib.ssdm_shard_finish_end(algos[root][0]._h, out.ctypes.data_as(ctypes.c_void_p), out.strides[0]))
    for algo, _ in algos:
        algo.close()
    return out, weights


@pytest.mark.parametrize("world,root,n,h,w,dtype,kwargs", [
    (2, 0, 5, 150, 203, np.uint8, {}),
    (3, 2, 8, 130, 260, np.uint16, {"levels": 4}),
    (4, 1, 9, 100, 136, np.uint8, {"energy": "sobel"}),
    (3, 0, 7, 120, 168, np.uint8, {"map_type": "max"}),
])
def test_sharded_phases_equal_the_host_build_and_the_single_handle_result(world, root, n, h, w, dtype, kwargs):
    frames = synthetic.make_stack(n, h, w, np.dtype(dtype), seed=90 + world)
    out, weights = _sharded_on_one_gpu(frames, world, root, **kwargs)
    st = {}
    emu = du.emu_sharded_focus_stack(frames, world, root, stages=st, **kwargs)
    if kwargs.get("map_type") == "max":         # expf of the device and of the host's libm differ in the last bits
        assert np.abs(weights - st["weights"]).max() < 2e-6
        du.assert_within_1lsb(out, emu, "sharded softmax vs host build", max_frac=0.02)
    else:
        assert np.array_equal(out, emu) and np.array_equal(weights, st["weights"])      # same code, same order
    single = DepthMapStack(**kwargs)
    ref = single.focus_stack_arrays(frames)
    ref_w = np.array([single.get_plane("weights", i) for i in range(n)])
    assert np.abs(weights - ref_w).max() < 1e-6
    du.assert_within_1lsb(out, ref, "sharded vs one handle", max_frac=0.02)


def test_sharded_stack_of_one_rank_is_the_plain_stack():
    from shinestacker_b200.sharded import ShardedDepthMapStack
    frames = synthetic.make_stack(4, 120, 168, np.dtype(np.uint8), seed=95)
    algo = ShardedDepthMapStack(0, 1)
    assert np.array_equal(algo.focus_stack_arrays(frames, 4), DepthMapStack().focus_stack_arrays(frames))
    assert np.array_equal(algo.focus_stack_arrays(frames[:3], 3), DepthMapStack().focus_stack_arrays(frames[:3]))
    algo.close()


def test_two_processes_share_one_stack_through_cuda_ipc():
    # two ranks (both on device 0 when the box has one GPU) run the real multi-process path: TCP rendezvous, CUDA IPC
    # handles of the partial-sum plane and the blended pyramid, barriers between the phases
    import json
    import os
    import socket
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), LOCAL_RANK="0", WORLD_SIZE="2", MASTER_ADDR="127.0.0.1",
                   MASTER_PORT=str(port - 36))
        procs.append(subprocess.Popen([sys.executable, os.path.join(root, "tools", "time_depthmap_sharded.py"), "--frames", "7",
                                       "--height", "240", "--width", "320", "--steps", "1", "--check"],
                                      env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True))
    outs = []
    for p in procs:
        try:
            o, e = p.communicate(timeout=180)
        except subprocess.TimeoutExpired:
            for q in procs:
                q.kill()
            raise
        assert p.returncode == 0, e[-2000:]
        outs.append(o)
    line = json.loads(outs[0].strip().splitlines()[-1])
    assert line["n_gpus"] == 2 and line["vs_one_handle"]["max_abs_diff_lsb"] <= 1
    assert line["vs_one_handle"]["fraction_differing"] < 0.02
