"""CPU tests of the AlignFrames warp stage: the integer oracle against cv2
(the library the reference calls, align.py:228-251) -- bit-exact for uint8 and for uint16 + rigid
transforms --, the product's kernels + launch sequence compiled for the host against both, and the host
mirror's estimator / option contract.  Nothing here needs a GPU."""
import numpy as np
import pytest

import align_util as au
from oracle import align_oracle as ao
from shinestacker_b200 import align as al
from shinestacker_b200 import InvalidOptionError, ShapeError, BitDepthError

cv2 = pytest.importorskip("cv2")
PINNED = [(dt, name) for dt in (np.uint8, np.uint16) for name, _ in au.TRANSFORMS
          if not (dt == np.uint16 and name == "homography")]


def _img(dtype, h=131, w=187, seed=0):
    return np.random.default_rng(seed).integers(0, np.iinfo(dtype).max + 1, (h, w, 3)).astype(dtype)


@pytest.mark.parametrize("dtype,name", PINNED)
@pytest.mark.parametrize("mode", list(au.MODES))
def test_oracle_and_host_build_of_the_kernels_match_cv2(dtype, name, mode):
    m = dict(au.TRANSFORMS)[name]
    img = _img(dtype, seed=len(name))
    bval = (17, 99, 250, 0) if mode == "BORDER_CONSTANT" else (0, 0, 0, 0)
    ref = au.cv2_warp_stage(img, m, mode, bval)
    k = ao.fixed_point_kernel(dtype, 21, 50)
    assert np.array_equal(ao.warp_and_fill(img, m, au.MODES[mode], bval, k), ref)
    emu = au.EmuWarper(img.shape[0], img.shape[1], dtype)
    assert np.array_equal(emu.warp(img, m, mode, bval, k), ref)
    emu.close()


def test_uint16_homography_is_pinned_where_cv2_defines_it():
    # cv2 4.13 routes uint16 warpPerspective through float kernels: bit-exact wherever the four taps lie in the image
    # (every border mode) and everywhere with BORDER_CONSTANT (a tap outside -> the border value).  With
    # BORDER_REPLICATE cv2 leaves the other pixels of the destination UNWRITTEN (checked with a sentinel): there the
    # model replicates the edge pixel.
    rng = np.random.default_rng(11)
    for trial in range(5):
        h, w = int(rng.integers(60, 200)), int(rng.integers(60, 260))
        img = rng.integers(0, 65536, (h, w, 3)).astype(np.uint16)
        m = au.homography(rng.uniform(-10, 10), rng.uniform(0.9, 1.1), rng.uniform(-15, 15), rng.uniform(-15, 15),
                          rng.uniform(-3e-5, 3e-5), rng.uniform(-3e-5, 3e-5))
        inside = ao.float_perspective_all_taps_inside(m, h, w)
        assert inside.mean() > 0.5
        emu = au.EmuWarper(h, w, np.uint16)
        bval = (1000.0, 20000.0, 65535.0, 0.0)
        ref = cv2.warpPerspective(img, m, (w, h), borderMode=cv2.BORDER_CONSTANT, borderValue=bval)
        out = emu.warp(img, m, "BORDER_CONSTANT", bval)
        assert np.array_equal(out, ref)
        assert np.array_equal(out, ao.warp(img, m, ao.BORDER_CONSTANT, bval))
        sentinel = np.full((h, w, 3), 12345, np.uint16)
        ref = cv2.warpPerspective(img, m, (w, h), dst=sentinel.copy(), borderMode=cv2.BORDER_REPLICATE)
        out = emu.warp(img, m, "BORDER_REPLICATE")
        assert np.array_equal(out, ao.warp(img, m, ao.BORDER_REPLICATE))
        assert np.array_equal(out[inside], ref[inside])
        assert (ref[~inside] == 12345).all()                     # cv2 does not write them
        emu.close()


@pytest.mark.parametrize("dtype", [np.uint8, np.uint16])
@pytest.mark.parametrize("sigma", [50, 3.0, 0.8])
def test_fixed_point_blur_model(dtype, sigma):
    img = _img(dtype, 64, 90, seed=3)
    k = al.fixed_point_blur_kernel(dtype, sigma)
    assert np.array_equal(k, ao.fixed_point_kernel(dtype, 21, sigma)) and k.sum() == (256 if dtype == np.uint8 else 65536)
    assert np.array_equal(ao.gaussian_blur_fixed(img, k), cv2.GaussianBlur(img, (21, 21), sigmaX=sigma))


def test_engine_rejects_bad_arguments():
    emu = au.EmuWarper(40, 50, np.uint8)
    img = _img(np.uint8, 40, 50)
    with pytest.raises(ValueError):
        emu.warp(img, au.rigid(1, 1, 0, 0), "BORDER_REPLICATE_BLUR", kernel=None)
    with pytest.raises(ValueError):
        emu.warp(img, au.rigid(1, 1, 0, 0), "BORDER_REPLICATE_BLUR", kernel=[100] * 21)   # does not sum to 256
    with pytest.raises(ValueError):
        emu.warp(img, np.zeros((3, 3)), "BORDER_REPLICATE")                              # singular homography
    # a degenerate affine matrix maps everything to the origin, exactly as cv2 does (D = 0 -> inverse 0)
    assert np.array_equal(emu.warp(img, np.zeros((2, 3)), "BORDER_REPLICATE"),
                          cv2.warpAffine(img, np.zeros((2, 3)), (50, 40), borderMode=cv2.BORDER_REPLICATE))


def _scene(h=384, w=512, seed=5):
    rng = np.random.default_rng(seed)
    img = np.zeros((h, w, 3), np.uint8)
    for _ in range(60):
        c = tuple(int(v) for v in rng.integers(60, 255, 3))
        p = (int(rng.integers(0, w)), int(rng.integers(0, h)))
        if rng.random() < 0.5:
            cv2.circle(img, p, int(rng.integers(5, 40)), c, int(rng.integers(1, 4)))
        else:
            q = (int(rng.integers(0, w)), int(rng.integers(0, h)))
            cv2.rectangle(img, p, q, c, int(rng.integers(1, 3)))
    return cv2.add(img, rng.integers(0, 12, img.shape).astype(np.uint8))


def test_estimator_recovers_a_known_rigid_transform_like_the_reference_test():
    # reference tests/test_0031_align_precision.py:62-65: angle < 0.005 deg, shift < 0.2 px, scale < 1e-4;
    # the estimator is the reference's own cv2 pipeline, so only loose sanity bounds are asserted here
    img = _scene()
    m_true = cv2.getRotationMatrix2D((256, 192), 2.0, 1.0)
    m_true[:, 2] += (6.0, -4.0)
    moved = cv2.warpAffine(img, m_true, (512, 384))
    kp_0, kp_1, good = al.detect_and_compute(img, moved)
    assert len(good) >= 20
    src = np.float32([kp_0[g.queryIdx].pt for g in good]).reshape(-1, 1, 2)
    dst = np.float32([kp_1[g.trainIdx].pt for g in good]).reshape(-1, 1, 2)
    m, _ = al.find_transform(src, dst)
    ang = np.degrees(np.arctan2(m[1, 0], m[0, 0]))
    ang_true = np.degrees(np.arctan2(m_true[1, 0], m_true[0, 0]))
    assert abs(ang - ang_true) < 0.05 and np.abs(m[:, 2] - m_true[:, 2]).max() < 0.5
    assert abs(np.hypot(m[0, 0], m[1, 0]) - 1.0) < 1e-3
    full = al.rescale_transform(np.float64(m), al.ALIGN_RIGID, 2, (768, 1024), (384, 512))
    assert full.dtype == np.float32 and np.allclose(full[:, 2], m[:, 2] * 2) and np.allclose(full[:, :2], m[:, :2])


def test_option_contract():
    with pytest.raises(ValueError):
        al.validate_align_config(al.DETECTOR_SIFT, al.DESCRIPTOR_ORB, al.MATCHING_KNN)
    with pytest.raises(ValueError):
        al.validate_align_config(al.DETECTOR_ORB, al.DESCRIPTOR_ORB, al.MATCHING_KNN)
    al.validate_align_config(al.DETECTOR_ORB, al.DESCRIPTOR_ORB, al.MATCHING_NORM_HAMMING)
    with pytest.raises(InvalidOptionError):
        al.find_transform(None, None, method="LSQ")
    with pytest.raises(InvalidOptionError):
        al.find_transform(None, None, transform="ALIGN_AFFINE")
    with pytest.raises(InvalidOptionError):
        al.get_good_matches(None, None, {"match_method": "L2"})
    img = np.zeros((32, 32, 3), np.uint8)
    with pytest.raises(InvalidOptionError):
        al.align_images(img, img, alignment_config={"border_mode": "BORDER_WRAP"})
    with pytest.raises(ShapeError):
        al.align_images(img, img[:16])
    with pytest.raises(BitDepthError):
        al.align_images(img, img.astype(np.uint16))
    a = al.AlignFrames(transform=al.ALIGN_HOMOGRAPHY, detector=al.DETECTOR_ORB, subsample=1, plot_summary=True)
    assert a.enabled and a.min_matches == 4 and a.feature_config["detector"] == "ORB"
    assert a.alignment_config["subsample"] == 1 and a.alignment_config["border_mode"] == "BORDER_REPLICATE_BLUR"


@pytest.mark.parametrize("h,w,dtype", [(10, 13, np.uint8), (25, 21, np.uint16), (64, 3, np.uint8), (1, 50, np.uint8)])
def test_frames_smaller_than_the_blur_kernel(h, w, dtype):
    img = _img(dtype, h, w, seed=h)
    k = ao.fixed_point_kernel(dtype, 21, 50)
    emu = au.EmuWarper(h, w, dtype)
    for _, m in au.TRANSFORMS[:2]:
        for mode in au.MODES:
            assert np.array_equal(emu.warp(img, m, mode, kernel=k), au.cv2_warp_stage(img, m, mode))
    emu.close()


def test_randomised_transforms_match_cv2():
    # large rotations, mirrors, frames mapped (almost) completely outside, arbitrary affine matrices, strong
    # homographies (uint8), random border values (saturated like cv2) and three blur sigmas
    rng = np.random.default_rng(123)
    kernels = {(dt, s): ao.fixed_point_kernel(dt, 21, s) for dt in (np.uint8, np.uint16) for s in (50, 5, 1.5)}
    for t in range(40):
        dt = [np.uint8, np.uint16][t % 2]
        h, w = int(rng.integers(8, 90)), int(rng.integers(8, 120))
        img = rng.integers(0, np.iinfo(dt).max + 1, (h, w, 3)).astype(dt)
        kind = t % 5
        if kind == 0:
            m = au.rigid(rng.uniform(-180, 180), rng.uniform(0.5, 2), rng.uniform(-w, w), rng.uniform(-h, h))
        elif kind == 1:
            m = au.rigid(rng.uniform(-5, 5), rng.uniform(0.9, 1.1), rng.uniform(-3 * w, 3 * w), rng.uniform(-3 * h, 3 * h))
        elif kind == 2:
            m = np.array([[-1.0, 0.0, w - 1.0], [0.0, 1.0, 0.0]]) + rng.normal(0, 1e-3, (2, 3))
        elif kind == 3:
            m = rng.normal(0, 1, (2, 3))
            m[:, 2] *= 20
        else:
            if dt == np.uint16:
                continue
            m = au.homography(rng.uniform(-20, 20), rng.uniform(0.8, 1.2), rng.uniform(-20, 20), rng.uniform(-20, 20),
                              rng.uniform(-1e-3, 1e-3), rng.uniform(-1e-3, 1e-3))
        sigma = float(rng.choice([50, 5, 1.5]))
        k = kernels[(dt, sigma)]
        emu = au.EmuWarper(h, w, dt)
        for mode in au.MODES:
            bval = tuple(rng.integers(0, 300, 4).tolist()) if mode == "BORDER_CONSTANT" else (0, 0, 0, 0)
            ref = au.cv2_warp_stage(img, m, mode, bval, sigma)
            assert np.array_equal(ao.warp_and_fill(img, m, au.MODES[mode], bval, k), ref), (t, mode)
            assert np.array_equal(emu.warp(img, m, mode, bval, k), ref), (t, mode)
        emu.close()
