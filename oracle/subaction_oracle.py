"""CPU restatement (TEST INFRASTRUCTURE ONLY -- never imported by the product path) of the two elementwise passes that
the sub-action kernels of csrc/ss_sub_lib.cuh replace.  numpy only; pinned against the live reference by
tests/test_subactions_cpu.py and against golden vectors made from it (tests/golden/make_golden_subactions.py).

  apply_lut        reference src/shinestacker/algorithms/balance.py:27-48 (CorrectionMapBase.apply_lut / adjust)
  correct_vignette reference src/shinestacker/algorithms/vignetting.py:16-20,71-97 (sigmoid_model, correct_vignetting)
"""
import numpy as np


def apply_lut(img, luts):
    """``luts``: (1, num_values) -> the one table maps every channel (balance.py:35-36: cv2.LUT / np.take on the whole
    image); (3, num_values) -> table c maps channel c of the B, G, R image (balance.py:43-47)."""
    luts = np.asarray(luts)
    assert luts.ndim == 2 and luts.shape[0] in (1, 3) and luts.dtype == img.dtype
    out = np.empty_like(img)
    for c in range(3):
        out[..., c] = luts[c if luts.shape[0] == 3 else 0][img[..., c]]
    return out


def sigmoid_model(r, i0, k, r0):
    """vignetting.py:16-20 (float64)."""
    with np.errstate(over="ignore"):
        return i0 / (1.0 + np.exp(np.minimum(10, np.exp(np.clip(k * (r - r0), -10, 10)))))


def correct_vignette(img, params, v0=None, max_correction=1.0, black_threshold=1.0):
    """vignetting.py:71-97 with the model parameters given (the fit is host work on both sides)."""
    i0, k, r0 = (float(p) for p in params)
    if v0 is None:
        v0 = sigmoid_model(0, i0, k, r0)
    h, w = img.shape[:2]
    y, x = np.ogrid[:h, :w]
    r = np.sqrt((x - w / 2) ** 2 + (y - h / 2) ** 2)
    vign = np.clip(sigmoid_model(r, i0, k, r0) / v0, 1e-6, 1)
    if max_correction < 1:
        vign = (1.0 - max_correction) + vign * max_correction
    thr = black_threshold if img.dtype == np.uint8 else black_threshold * 256
    vign = np.repeat(vign[:, :, None], 3, axis=2)
    vign[np.min(img, axis=2) < thr, :] = 1
    return np.clip(img / vign, 0, 255 if img.dtype == np.uint8 else 65535).astype(img.dtype)
