"""GPU parity tests of the AlignFrames warp stage (through ctypes -> C ABI, include/shinestack_b200_align.h)
against cv2 itself -- the library the reference calls at align.py:228-251 -- bit-exact for uint8 and for
uint16 + rigid transforms, in all three border modes."""

import numpy as np
import pytest

import align_util as au
from oracle import align_oracle as ao
from shinestacker_b200 import align as al
from shinestacker_b200 import PyramidStack, synthetic
from shinestacker_b200.pyramid import _check

cv2 = pytest.importorskip("cv2")
pytestmark = pytest.mark.gpu

PINNED = [(dt, name) for dt in (np.uint8, np.uint16) for name, _ in au.TRANSFORMS
          if not (dt == np.uint16 and name == "homography")]


def _img(dtype, h, w, seed):
    return np.random.default_rng(seed).integers(0, np.iinfo(dtype).max + 1, (h, w, 3)).astype(dtype)


@pytest.mark.parametrize("dtype,name", PINNED)
@pytest.mark.parametrize("mode", list(au.MODES))
def test_cuda_warp_matches_cv2(dtype, name, mode):
    m = dict(au.TRANSFORMS)[name]
    img = _img(dtype, 333, 517, seed=len(name) + 7)
    bval = (17, 99, 250, 0) if mode == "BORDER_CONSTANT" else (0, 0, 0, 0)
    ref = au.cv2_warp_stage(img, m, mode, bval)
    wp = al.Warper(img.shape[:2], dtype)
    tr = al.ALIGN_HOMOGRAPHY if np.asarray(m).size == 9 else al.ALIGN_RIGID
    out = wp.warp(img, m, tr, mode, bval)
    assert np.array_equal(out, ref), f"{(out != ref).mean():.5f} of the samples differ"
    assert wp.last_kernel_ms > 0
    wp.close()


def test_uint16_homography_matches_the_float_model():
    yy, xx = np.mgrid[0:300, 0:420]
    img = np.stack([xx * 150 + yy * 17, xx * 5 + yy * 200, xx * 61 + yy * 122], -1).astype(np.uint16)
    m = au.homography(0.7, 1.01, 5.5, -8.25)
    wp = al.Warper(img.shape[:2], np.uint16)
    out = wp.warp(img, m, al.ALIGN_HOMOGRAPHY, "BORDER_REPLICATE")
    assert np.array_equal(out, ao.warp(img, m, ao.BORDER_REPLICATE))
    ref = cv2.warpPerspective(img, m, (420, 300), borderMode=cv2.BORDER_REPLICATE)
    inner = (slice(30, 270), slice(30, 390))
    assert np.abs(out[inner].astype(int) - ref[inner].astype(int)).max() <= 1


@pytest.mark.parametrize("dtype", [np.uint8, np.uint16])
def test_full_size_frame_and_other_blur_sigma(dtype):
    img = _img(dtype, 3072, 4096, seed=11)
    m = au.rigid(0.35, 1.002, 18.4, -11.7)
    wp = al.Warper(img.shape[:2], dtype)
    for sigma in (50, 4.0):
        out = wp.warp(img, m, al.ALIGN_RIGID, "BORDER_REPLICATE_BLUR", border_blur=sigma)
        assert np.array_equal(out, au.cv2_warp_stage(img, m, "BORDER_REPLICATE_BLUR", border_blur=sigma))
    wp.close()


def test_align_images_end_to_end_matches_the_reference_pipeline():
    # same estimator calls (cv2, host) + warp on the GPU == the reference's align_images, image for image
    from test_align_cpu import _scene
    img = _scene(768, 1024, seed=9)
    m_true = cv2.getRotationMatrix2D((512, 384), 1.2, 1.004)
    m_true[:, 2] += (9.0, -6.5)
    moved = cv2.warpAffine(img, m_true, (1024, 768), borderMode=cv2.BORDER_REPLICATE)
    n, m, out = al.align_images(img, moved)
    assert n > 100 and m is not None and m.dtype == np.float32      # estimated on the 2x sub-sampled frames
    assert np.array_equal(out, au.cv2_warp_stage(moved, m, "BORDER_REPLICATE_BLUR"))
    # the aligned frame is close to the reference frame again
    inner = (slice(40, -40), slice(40, -40))
    assert np.abs(out[inner].astype(int) - img[inner].astype(int)).mean() < 10.0   # the scene carries +-6 noise


def test_aligned_frames_feed_the_stacker_without_leaving_hbm():
    # config 5 of BASELINE.json in miniature: warp on the device, ss_push_frame_device, same fused image as
    # stacking the cv2-aligned frames
    frames = synthetic.make_stack(5, 300, 408, np.uint8, 13)
    mats = [au.rigid(0.1 * i, 1 + 0.001 * i, 1.5 * i, -0.75 * i) for i in range(5)]
    aligned = [au.cv2_warp_stage(f, m, "BORDER_REPLICATE_BLUR") for f, m in zip(frames, mats)]
    expect = PyramidStack().focus_stack_arrays(aligned)
    algo = PyramidStack(batch_frames=2)
    algo._set_metadata(frames[0])
    algo._ensure_handle(frames[0].shape, frames[0].dtype, 5)
    wp = al.Warper(frames[0].shape[:2], np.uint8)
    for f, m in zip(frames, mats):
        assert wp.warp(f, m, al.ALIGN_RIGID, "BORDER_REPLICATE_BLUR", download=False) is None
        ptr, pitch = wp.result_device()
        _check(algo._lib.ss_push_frame_device(algo._h, ptr, pitch))
        _check(algo._lib.ss_sync(algo._h))      # the warper's output buffer is reused by the next frame
    out = algo._finish(5, frames[0].shape[:2])
    assert np.array_equal(out, expect)


@pytest.mark.parametrize("dtype", [np.uint8, np.uint16])
def test_ecc_estimator_on_the_gpu(dtype):
    img = au.textured_scene(1536, 2048, seed=21)
    mov, m_true = au.moved_frame(img, -0.45, 1.006, 17.0, -12.0)
    if dtype == np.uint16:
        img, mov = img.astype(np.uint16) * 257, mov.astype(np.uint16) * 257
    wp = al.Warper(img.shape[:2], dtype)
    wp.ecc_set_reference(img)
    m, info = wp.ecc_estimate(mov)
    d_ang, d_scale, d_pos = au.transform_error(m, m_true, img.shape)
    assert d_ang < 0.005 and d_scale < 1e-4 and d_pos < 0.2, (d_ang, d_scale, d_pos, info)   # reference test_0031 bounds
    assert d_pos < 0.03 and info["rho"] > 0.98
    # same iteration as the host build of the kernels (different summation order only)
    emu = au.EmuWarper(img.shape[0], img.shape[1], dtype)
    emu.ecc_set_reference(img)
    m_emu, _ = emu.ecc_estimate(mov)
    assert au.transform_error(m, m_emu, img.shape)[2] < 5e-3
    # warp of the frame that is already on the device == cv2 with the same matrix
    out = wp.warp(None, m, al.ALIGN_RIGID, "BORDER_REPLICATE_BLUR")
    assert np.array_equal(out, au.cv2_warp_stage(mov, m, "BORDER_REPLICATE_BLUR"))
    wp.close()


def test_align_frames_sub_action_with_both_estimators():
    from unittest.mock import MagicMock
    img = au.textured_scene(768, 1024, seed=22)
    frames = [img] + [au.moved_frame(img, 0.2 * i, 1 + 0.002 * i, 3.0 * i, -2.0 * i)[0] for i in (1, 2)]
    for estimator in ("features", "ecc"):
        action = al.AlignFrames(estimator=estimator)
        process = MagicMock()
        process.ref_idx, process.counts = 0, 3
        process.img_ref.side_effect = lambda idx: frames[idx]
        action.begin(process)
        assert action.run_frame(0, 0, frames[0]) is frames[0]
        for i in (1, 2):
            out = action.run_frame(i, 0, frames[i])
            inner = (slice(40, -40), slice(40, -40))
            assert np.abs(out[inner].astype(int) - img[inner].astype(int)).mean() < 3.0
            assert action.n_matches[i] > 0
        if estimator == "ecc":
            assert process.img_ref.call_count == 1      # the reference pyramid stayed in HBM
        action.end()


def test_two_ecc_sub_actions_sharing_one_warper_never_align_to_a_stale_reference():
    """The warper (and the ECC reference pyramid inside it) is shared process-wide per geometry: an AlignFrames whose
    reference was replaced by somebody else must upload it again instead of trusting its own bookkeeping."""
    from unittest.mock import MagicMock
    scene_a, scene_b = au.textured_scene(512, 768, seed=31), au.textured_scene(512, 768, seed=32)
    mov_a, mov_b = au.moved_frame(scene_a, 0.3, 1.002, 5.0, -3.0)[0], au.moved_frame(scene_b, -0.2, 0.999, -4.0, 6.0)[0]
    acts = []
    for ref in (scene_a, scene_b):
        action, process = al.AlignFrames(estimator="ecc"), MagicMock()
        process.ref_idx, process.counts = 0, 3
        process.img_ref.side_effect = lambda idx, ref=ref: ref
        action.begin(process)
        acts.append((action, process))
    inner = (slice(40, -40), slice(40, -40))
    out = acts[0][0].run_frame(1, 0, mov_a)                      # A sets its reference
    assert np.abs(out[inner].astype(int) - scene_a[inner].astype(int)).mean() < 3.0
    out = acts[1][0].run_frame(1, 0, mov_b)                      # B replaces it in the shared warper
    assert np.abs(out[inner].astype(int) - scene_b[inner].astype(int)).mean() < 3.0
    out = acts[0][0].run_frame(2, 0, mov_a)                      # A must notice and upload scene_a again
    assert np.abs(out[inner].astype(int) - scene_a[inner].astype(int)).mean() < 3.0
    assert acts[0][1].img_ref.call_count == 2 and acts[1][1].img_ref.call_count == 1
    out = acts[0][0].run_frame(1, 0, mov_a)                      # ... and now it is current again
    assert acts[0][1].img_ref.call_count == 2


def test_fused_align_and_stack_pipeline_recovers_the_undisplaced_result():
    from shinestacker_b200.pipeline import align_and_stack_arrays
    from tools.time_pipeline import displaced_stack
    frames = synthetic.make_stack(7, 600, 800, np.uint8, 23)
    moved, truth, _ = displaced_stack(frames)
    out, ms = align_and_stack_arrays(list(moved), return_transforms=True)
    c = np.array([400.0, 300.0, 1.0])
    assert max(np.abs(m @ c - t @ c).max() for m, t in zip(ms, truth)) < 0.1
    # identical to aligning with cv2 (same matrices) and stacking the aligned files, which is what the reference job does
    aligned = [f if i == 3 else au.cv2_warp_stage(f, m, "BORDER_REPLICATE_BLUR") for i, (f, m) in enumerate(zip(moved, ms))]
    assert np.array_equal(out, PyramidStack().focus_stack_arrays(aligned))
    # and it resembles the fusion of the undisplaced stack away from the borders (each frame was resampled
    # twice and the scene is pixel-scale noise, so only a loose statistical statement holds)
    plain = PyramidStack().focus_stack_arrays(frames)
    inner = (slice(60, -60), slice(60, -60))
    a, b = out[inner].astype(np.float64).ravel(), plain[inner].astype(np.float64).ravel()
    assert np.corrcoef(a, b)[0, 1] > 0.8


def test_pipeline_gives_the_same_image_from_pageable_and_from_pinned_frames():
    # pageable frames are staged by worker threads into a ring of pinned slots, pinned frames upload from where they
    # are; both are prefetched one frame ahead (ssal_prefetch) -- same transforms, same image
    import ctypes
    from shinestacker_b200 import _native
    from shinestacker_b200.pipeline import align_and_stack_arrays
    from tools.time_pipeline import displaced_stack
    lib = _native.load()
    frames = synthetic.make_stack(11, 300, 400, np.uint8, 29)
    moved, _, _ = displaced_stack(frames)
    pin_moved, _, pin = displaced_stack(frames, lib)
    assert lib.ss_host_is_pinned(ctypes.c_void_p(pin_moved[0].ctypes.data)) == 1
    assert lib.ss_host_is_pinned(ctypes.c_void_p(moved[0].ctypes.data)) == 0
    try:
        a, ma = align_and_stack_arrays(list(moved), return_transforms=True)
        b, mb = align_and_stack_arrays(list(pin_moved), return_transforms=True)
        c, mc = align_and_stack_arrays(list(moved), ref_idx=0, return_transforms=True)      # reference frame first
    finally:
        _check(lib.ss_host_free(pin))
    # (the ECC sums are accumulated with atomics: the transforms agree to rounding, not necessarily to the bit)
    assert all(np.allclose(x, y, rtol=0, atol=1e-6) for x, y in zip(ma, mb))
    assert np.mean(a != b) < 1e-3
    assert c.shape == a.shape and np.array_equal(mc[0], np.array([[1.0, 0, 0], [0, 1.0, 0]]))


def test_pipeline_abort_leaves_a_shared_warper_usable():
    from shinestacker_b200.pipeline import align_and_stack_arrays
    from tools.time_pipeline import displaced_stack
    frames = synthetic.make_stack(5, 240, 320, np.uint8, 31)
    moved, _, _ = displaced_stack(frames)
    bad = list(moved)
    bad[1] = np.random.default_rng(1).integers(0, 256, moved[0].shape).astype(np.uint8)     # unrelated content
    wp = al.Warper(moved[0].shape[:2], np.uint8)
    algo = PyramidStack()
    try:
        from shinestacker_b200 import NativeLibraryError
        with pytest.raises((al.AlignmentError, NativeLibraryError)):
            align_and_stack_arrays(bad, stack_algo=algo, warper=wp)       # fails while frame 2 is being prefetched
        again = align_and_stack_arrays(list(moved), stack_algo=algo, warper=wp)
        fresh = align_and_stack_arrays(list(moved))
        assert np.mean(again != fresh) < 1e-3          # (the ECC sums use atomics: equal up to rounding of the fit)
    finally:
        wp.close()


def test_prefetched_frames_and_double_buffered_results():
    import ctypes
    from shinestacker_b200 import _native, NativeLibraryError
    lib = _native.load()
    frames = synthetic.make_noise_stack(3, 240, 320, np.uint8, 43)
    ref = frames[0]
    moving = [au.cv2_warp_stage(ref, au.rigid(0.3, 1.0, 2.5, -1.5), "BORDER_REPLICATE"),
              au.cv2_warp_stage(ref, au.rigid(-0.2, 1.0, -3.0, 2.0), "BORDER_REPLICATE")]
    pin = ctypes.c_void_p()
    _check(lib.ss_host_alloc(ctypes.byref(pin), 2 * ref.nbytes))
    buf = np.frombuffer((ctypes.c_uint8 * (2 * ref.nbytes)).from_address(pin.value), np.uint8).reshape((2,) + ref.shape)
    buf[:] = moving
    wp, plain = al.Warper(ref.shape[:2], np.uint8), al.Warper(ref.shape[:2], np.uint8)
    try:
        wp.ecc_set_reference(ref)
        plain.ecc_set_reference(ref)
        with pytest.raises(NativeLibraryError):
            wp.use_prefetched()                     # nothing prefetched yet
        wp.prefetch(buf[0])
        with pytest.raises(NativeLibraryError):
            wp.prefetch(buf[1])                     # one outstanding frame at most
        results, ms = [], []
        for k in range(2):
            wp.use_prefetched()
            if k == 0:
                wp.prefetch(buf[1])                 # uploads while frame 0 is fitted and warped
            m, _ = wp.ecc_estimate(None)
            m_plain, _ = plain.ecc_estimate(moving[k])
            assert np.allclose(m, m_plain, rtol=0, atol=1e-6)
            ms.append(m)
            wp.warp(None, m, al.ALIGN_RIGID, "BORDER_REPLICATE", download=False)
            results.append(wp.result_device())
            assert np.array_equal(plain.warp(moving[k], m, al.ALIGN_RIGID, "BORDER_REPLICATE"),
                                  au.cv2_warp_stage(moving[k], m, "BORDER_REPLICATE"))
        # both results are still in device memory (double-buffered outputs): feed them to a stacker and compare
        assert results[0][0].value != results[1][0].value
        algo = PyramidStack()
        algo._set_metadata(ref)
        algo._ensure_handle(ref.shape, ref.dtype, 2)
        for ptr, pitch in results:
            _check(lib.ss_push_frame_device(algo._h, ptr, pitch))
        fused = algo._finish(2, ref.shape[:2])
        expect = PyramidStack().focus_stack_arrays([plain.warp(moving[k], ms[k], al.ALIGN_RIGID, "BORDER_REPLICATE")
                                                    for k in range(2)])
        assert np.array_equal(fused, expect)
    finally:
        wp.close()
        plain.close()
        _check(lib.ss_host_free(pin))


def test_device_resident_sources():
    # ssal_warp_device / ssdm_push_frame_device: frames that already live in HBM (here: the output of another warp)
    import ctypes
    from shinestacker_b200 import DepthMapStack
    frames = synthetic.make_noise_stack(3, 120, 168, np.uint8, 41)
    m1, m2 = au.rigid(0.5, 1.01, 3.0, -2.0), au.rigid(-0.25, 0.995, -1.5, 4.0)
    a, b = al.Warper(frames[0].shape[:2], np.uint8), al.Warper(frames[0].shape[:2], np.uint8)
    dm, dm_ref = DepthMapStack(), DepthMapStack()
    dm._ensure_handle(frames[0].shape, frames[0].dtype, 3)
    staged = []
    for f in frames:
        first = a.warp(f, m1, al.ALIGN_RIGID, "BORDER_REPLICATE")
        ptr, pitch = a.result_device()
        out = np.empty_like(f)
        mat = np.ascontiguousarray(m2, np.float64)
        _check(b._lib.ssal_warp_device(b._h, ptr, pitch, mat.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), 0, 1, None,
                                       None, 0, out.ctypes.data_as(ctypes.c_void_p), out.strides[0]))
        assert np.array_equal(out, au.cv2_warp_stage(first, m2, "BORDER_REPLICATE"))
        _check(dm._lib.ssdm_push_frame_device(dm._h, ptr, pitch))
        _check(dm._lib.ssdm_sync_copies(dm._h))
        _check(dm._lib.ssdm_get_plane(dm._h, 2, len(staged), np.empty((120, 168), np.float32).ctypes.data_as(ctypes.c_void_p)))
        staged.append(first)
    fused = dm._blend_and_finish(3, frames[0].shape[:2], frames[0].dtype)
    assert np.array_equal(fused, dm_ref.focus_stack_arrays(staged))
