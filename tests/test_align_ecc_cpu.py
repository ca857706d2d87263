"""CPU tests of the gradient-based (ECC) estimator of AlignFrames: the product's per-pixel device function and
the host-side iteration (ss_ecc_kernels.cuh / ss_ecc_engine.h), compiled for the host, on synthetic similarity
warps.  The reference has no ECC estimator (its align.py fits sparse features), so parity is "unpinned"; the
contract is the one of the reference's own precision test (tests/test_0031_align_precision.py:62-65: angle
< 0.005 deg, shift < 0.2 px, scale < 1e-4) plus agreement with the reference's feature pipeline."""
import numpy as np
import pytest

import align_util as au
from shinestacker_b200 import align as al

cv2 = pytest.importorskip("cv2")


@pytest.mark.parametrize("angle,scale,tx,ty", [
    (0.4, 1.003, 7.3, -4.6),
    (-0.5, 0.992, -19.0, 14.0),       # BASELINE config 5: rot <= 0.5 deg, shift <= 20 px, scale 1 +- 0.01
    (0.05, 1.0, 0.3, 0.2),
    (2.0, 1.01, 30.0, -25.0),
    (0.0, 1.0, 0.0, 0.0),
])
def test_ecc_recovers_similarity_within_the_reference_tolerances(angle, scale, tx, ty):
    img = au.textured_scene(480, 640)
    mov, m_true = au.moved_frame(img, angle, scale, tx, ty)
    emu = au.EmuWarper(480, 640, np.uint8)
    emu.ecc_set_reference(img)
    m, info = emu.ecc_estimate(mov)
    d_ang, d_scale, d_pos = au.transform_error(m, m_true, img.shape)
    assert d_ang < 0.005 and d_scale < 1e-4 and d_pos < 0.2, (d_ang, d_scale, d_pos, info)
    assert d_pos < 0.02                         # in fact an order of magnitude better
    assert info["rho"] > 0.98 and info["overlap"] > 0.85
    # the aligned frame is the reference frame again (bilinear resampling twice + noise)
    out = emu.warp(mov, m, "BORDER_REPLICATE")
    inner = (slice(40, -40), slice(40, -40))
    assert np.abs(out[inner].astype(int) - img[inner].astype(int)).mean() < 3.0
    emu.close()


def test_ecc_agrees_with_the_reference_feature_pipeline():
    img = au.textured_scene(600, 800, seed=4)
    mov, m_true = au.moved_frame(img, 0.8, 1.005, 11.0, -7.5)
    kp_0, kp_1, good = al.detect_and_compute(mov, img)
    src = np.float32([kp_0[g.queryIdx].pt for g in good]).reshape(-1, 1, 2)
    dst = np.float32([kp_1[g.trainIdx].pt for g in good]).reshape(-1, 1, 2)
    m_feat, _ = al.find_transform(src, dst)
    emu = au.EmuWarper(600, 800, np.uint8)
    emu.ecc_set_reference(img)
    m_ecc, _ = emu.ecc_estimate(mov)
    e_feat = au.transform_error(m_feat, m_true, img.shape)
    e_ecc = au.transform_error(m_ecc, m_true, img.shape)
    assert e_ecc[2] <= max(e_feat[2], 0.02) and e_ecc[0] <= max(e_feat[0], 0.002)
    assert au.transform_error(m_ecc, m_feat, img.shape)[2] < 0.5
    emu.close()


def test_ecc_uint16_half_resolution_and_initial_guess():
    img8 = au.textured_scene(480, 640, seed=6)
    mov8, m_true = au.moved_frame(img8, -0.3, 1.002, 5.0, 9.0)
    img, mov = img8.astype(np.uint16) * 257, mov8.astype(np.uint16) * 257
    emu = au.EmuWarper(480, 640, np.uint16)
    emu.ecc_set_reference(img)
    m_half, _ = emu.ecc_estimate(mov, finest_level=1)       # like the reference's default subsample = 2
    assert au.transform_error(m_half, m_true, img.shape)[2] < 0.2
    m_full, info = emu.ecc_estimate(mov, init_m=m_half)
    assert au.transform_error(m_full, m_true, img.shape)[2] < 0.02
    emu.close()


def test_ecc_failure_modes():
    emu = au.EmuWarper(200, 260, np.uint8)
    img = au.textured_scene(200, 260, seed=2)
    with pytest.raises(ValueError):
        emu.ecc_estimate(img)                    # no reference yet
    emu.ecc_set_reference(img)
    with pytest.raises(ValueError):
        emu.ecc_estimate(np.full_like(img, 7))   # flat frame: nothing to align
    emu.close()
    with pytest.raises(al.InvalidOptionError):
        al.align_images(img, img, alignment_config={"estimator": "phase"})
    with pytest.raises(al.InvalidOptionError):
        al._align_images_ecc(img, img, al._merged(al.DEFAULT_ALIGNMENT_CONFIG, {"transform": al.ALIGN_HOMOGRAPHY}), {}, 0)
