"""Host-side logic of the widened rows without a GPU: DepthMapStack.focus_stack (file loop, callbacks,
cancellation, validation, option errors) with the native calls stubbed, and AlignFrames driven by the UNMODIFIED
reference host (StackJob + CombinedActions, reference stack_framework.py:246-302) with the device warp replaced by
the cv2 calls it is bit-identical to.  The pixel arithmetic itself is covered by the emulation and GPU tests."""
import os

import numpy as np
import pytest

import align_util as au
from oracle import ref_harness
from shinestacker_b200 import (DepthMapStack, AlignFrames, ShapeError, BitDepthError, ImageLoadError,
                               RunStopException, InvalidOptionError, synthetic)
from shinestacker_b200 import align as al
from test_host_logic import Proc, _write

cv2 = pytest.importorskip("cv2")


class FakeLib:
    """Records the ssdm_* calls of one stack."""

    def __init__(self):
        self.calls = []

    def __getattr__(self, name):
        def fn(*args):
            self.calls.append(name)
            return 0
        return fn


@pytest.fixture
def stubbed(monkeypatch):
    pushed = []

    def ensure(self, shape, dtype, n):
        self._check_options()
        self._geom = (int(shape[0]), int(shape[1]))
        self._frame = (tuple(int(v) for v in shape), np.dtype(dtype))
        self._h = 1

    def push_slot(self, i, slot):
        # what the upload would read from the pinned slot a worker thread staged the frame in
        shape, dtype = self._frame
        n = int(np.prod(shape)) * dtype.itemsize
        pushed.append((i, slot[1][:n].view(dtype).reshape(shape).copy()))

    monkeypatch.setattr(DepthMapStack, "_ensure_handle", ensure)
    monkeypatch.setattr(DepthMapStack, "_staging",
                        lambda self, nbytes, count: [(None, np.zeros(nbytes, np.uint8)) for _ in range(4)])
    monkeypatch.setattr(DepthMapStack, "_push_slot", push_slot)
    monkeypatch.setattr(DepthMapStack, "_sync_copies", lambda self: None)
    monkeypatch.setattr(DepthMapStack, "close", lambda self: None)
    return pushed


def _algo(**kw):
    a = DepthMapStack(**kw)
    a._lib = FakeLib()
    return a


@pytest.mark.parametrize("threads", [0, 3])
def test_depthmap_file_loop_order_and_callbacks(tmp_path, stubbed, threads):
    frames = synthetic.make_noise_stack(6, 40, 56, np.uint16, 1)
    paths = _write(tmp_path, frames)
    algo = _algo(decode_threads=threads)
    algo.process = Proc()
    out = algo.focus_stack(paths)
    assert out.shape == (40, 56, 3) and out.dtype == np.uint16
    assert [i for i, _ in stubbed] == list(range(6))
    for (_, img), f in zip(stubbed, frames):
        assert np.array_equal(img, f)
    assert algo.process.steps == list(range(12))           # depth_map.py:75 (i) and :113 (i + n), in order
    calls = [c for c in algo._lib.calls if c.startswith("ssdm_")]
    assert calls[:8] == ["ssdm_finish_begin"] + ["ssdm_blend_frame"] * 6 + ["ssdm_finish_end"]
    assert any("reading file (1/2) f000.tif" in m for m in algo.process.msgs)
    assert any("reading file (2/2) f005.tif" in m for m in algo.process.msgs)
    assert any("blend levels" in m for m in algo.process.msgs)


def test_depthmap_cancellation_and_errors(tmp_path, stubbed):
    frames = synthetic.make_noise_stack(4, 32, 48, np.uint8, 3)
    paths = _write(tmp_path, frames)
    algo = _algo()
    algo.process = Proc(stop_at=6)                          # during the blend pass
    with pytest.raises(RunStopException):
        algo.focus_stack(paths)
    assert "ssdm_reset" in algo._lib.calls and "ssdm_finish_end" not in algo._lib.calls
    bad = str(tmp_path / "zz_bad.tif")
    cv2.imwrite(bad, np.zeros((33, 48, 3), np.uint8))
    with pytest.raises(ShapeError):
        _algo().focus_stack(paths + [bad])
    cv2.imwrite(bad, np.zeros((32, 48, 3), np.uint16))
    with pytest.raises(BitDepthError):
        _algo().focus_stack(paths + [bad])
    with pytest.raises(RuntimeError):
        _algo().focus_stack(paths + [str(tmp_path / "missing.tif")])
    unsupported = str(tmp_path / "zz.bmp")
    cv2.imwrite(unsupported, frames[0])
    with pytest.raises(ImageLoadError):
        _algo().focus_stack(paths + [unsupported])
    with pytest.raises(ImageLoadError):
        _algo().focus_stack([])
    with pytest.raises(InvalidOptionError):
        _algo(energy="tenengrad").focus_stack(paths)
    with pytest.raises(InvalidOptionError):
        _algo(map_type="median").focus_stack(paths)


@pytest.mark.skipif(not ref_harness.reference_available(), reason="reference tree not present on this box")
@pytest.mark.parametrize("estimator", ["features"])
def test_reference_combined_actions_drive_our_align_frames(tmp_path, monkeypatch, estimator):
    ref_harness.import_reference()
    from shinestacker.algorithms.stack_framework import StackJob, CombinedActions

    img = au.textured_scene(300, 400, seed=12)
    frames = [au.moved_frame(img, 0.3 * (i - 1), 1 + 0.002 * (i - 1), 4.0 * (i - 1), -3.0 * (i - 1))[0] for i in range(3)]
    src = tmp_path / "proj" / "in"
    os.makedirs(src)
    for i, f in enumerate(frames):
        cv2.imwrite(str(src / f"img{i:02d}.tif"), f)

    warps = []

    class CpuWarper:            # stands in for the device warp with the cv2 calls it is bit-identical to
        def __init__(self, shape, dtype, device=0):
            pass

        def warp(self, image, m, transform, border_mode, border_value, border_blur, download=True):
            warps.append(np.array(m))
            return au.cv2_warp_stage(image, np.asarray(m), border_mode, tuple(border_value), border_blur)

        def close(self):
            pass

    monkeypatch.setattr(al, "Warper", CpuWarper)
    monkeypatch.setattr(al, "_WARPERS", {})
    action = AlignFrames(estimator=estimator, subsample=1)
    job = StackJob("job", str(tmp_path / "proj"), input_path="in", logger_name="test_b200_align")
    job.add_action(CombinedActions("align", [action], output_path="aligned"))
    job.run()
    out_dir = tmp_path / "proj" / "aligned"
    assert sorted(os.listdir(out_dir)) == ["img00.tif", "img01.tif", "img02.tif"]
    assert len(warps) == 2 and action.n_matches[0] > 0 and action.n_matches[2] > 0 and action.n_matches[1] == 0
    ref = cv2.imread(str(out_dir / "img01.tif"), cv2.IMREAD_UNCHANGED)
    assert np.array_equal(ref, frames[1])                   # the reference frame passes through untouched
    inner = (slice(30, -30), slice(30, -30))
    for name in ("img00.tif", "img02.tif"):
        got = cv2.imread(str(out_dir / name), cv2.IMREAD_UNCHANGED)
        assert np.abs(got[inner].astype(int) - img[inner].astype(int)).mean() < 3.0


def test_pipeline_sources_stage_pageable_frames_ahead_in_a_small_ring():
    # align_and_stack_arrays takes pinned frames as they are and copies pageable ones into a ring of pinned slots a few
    # positions ahead on worker threads; the ring is shorter than the stack, so slots are reused
    import ctypes
    from shinestacker_b200 import pipeline

    class Lib:
        def __init__(self, pinned_ids):
            self.pinned_ids, self.bufs, self.freed = pinned_ids, {}, 0

        def ss_host_is_pinned(self, ptr):
            return 1 if ptr.value in self.pinned_ids else 0

        def ss_host_alloc(self, ref, nbytes):
            buf = (ctypes.c_uint8 * nbytes)()
            addr = ctypes.addressof(buf)
            self.bufs[addr] = buf
            ctypes.cast(ref, ctypes.POINTER(ctypes.c_void_p))[0] = addr
            return 0

        def ss_host_free(self, ptr):
            self.freed += 1
            return 0

    frames = synthetic.make_noise_stack(11, 24, 40, np.uint16, 9)
    order = [i for i in range(11) if i != 5]                    # the moving frames (5 is the reference)
    lib = Lib({frames[2].ctypes.data, frames[7].ctypes.data})   # two of them already sit in page-locked memory
    src = pipeline._Sources(lib, frames)
    assert len(src.ring) == pipeline.STAGE_SLOTS
    try:
        for k, i in enumerate(order):
            got = src.get(order, k)
            assert np.array_equal(got, frames[i])
            assert (got is frames[i]) == (i in (2, 7))          # pinned frames are handed through, others are copies
            assert all(j <= k + pipeline.STAGE_SLOTS - 2 for j in src.pending)      # never further ahead than the ring allows
    finally:
        src.close()
    assert lib.freed == pipeline.STAGE_SLOTS and not src.ring
    all_pinned = pipeline._Sources(Lib({f.ctypes.data for f in frames}), frames)
    assert all_pinned.pool is None and not all_pinned.ring       # nothing to stage: no threads, no pinned ring
    all_pinned.close()
