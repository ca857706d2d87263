"""Generates tests/golden/subactions/*.npz by running the UNMODIFIED reference sub-actions ``BalanceFrames``
(src/shinestacker/algorithms/balance.py) and ``Vignetting`` / ``correct_vignetting`` (vignetting.py), imported through
tools/ref_loader.py, in the build container.  The reference's own tests for them (tests/test_0040_balance.py,
test_0050_vignetting.py) only assert "no exception", so these files pin the tables, the fit parameters and the
corrected frames that the oracle, the host mirrors and the CUDA passes are compared with.

    python tests/golden/make_golden_subactions.py
"""
import os
import sys
import tempfile
from unittest.mock import MagicMock

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from tools import ref_loader  # noqa: E402
from shinestacker_b200 import synthetic  # noqa: E402

OUT = os.path.join(HERE, "subactions")


def make_frames(n, h, w, dtype, seed):
    """A stack whose frames differ in exposure / contrast and carry a radial fall-off (something to balance / fit)."""
    dtype = np.dtype(dtype)
    maxv = 255 if dtype == np.uint8 else 65535
    frames = synthetic.make_stack(n, h, w, dtype, seed)
    rng = np.random.default_rng(seed)
    y, x = np.ogrid[:h, :w]
    r = np.sqrt((x - w / 2) ** 2 + (y - h / 2) ** 2) / np.sqrt((w / 2) ** 2 + (h / 2) ** 2)
    out = []
    for i, f in enumerate(frames):
        gain = 1.0 + 0.25 * (rng.random() - 0.5) * 2 if i else 1.0
        fall = 1.0 - 0.45 * r ** 2 * (0.8 + 0.4 * rng.random())
        g = f.astype(np.float64) * gain * fall[..., None] + maxv * 0.04
        out.append(np.clip(g, 0, maxv).astype(dtype))
    return out


BALANCE_CASES = [
    ("bal_u8_linear_lumi", 4, 120, 168, "uint8", 41, dict()),
    ("bal_u8_gamma_rgb", 3, 96, 150, "uint8", 42, dict(corr_map="GAMMA", channel="RGB", subsample=2)),
    ("bal_u8_match_lumi", 3, 90, 141, "uint8", 43, dict(corr_map="MATCH_HIST")),
    ("bal_u16_linear_rgb_mask", 3, 100, 160, "uint16", 44, dict(channel="RGB", mask_size=0.8, subsample=1)),
    ("bal_u16_gamma_lumi_fast", 3, 88, 136, "uint16", 45, dict(corr_map="GAMMA", fast_subsampling=True, subsample=4)),
    ("bal_u8_linear_hsv", 3, 80, 120, "uint8", 46, dict(channel="HSV", subsample=1)),
    ("bal_u8_match_hls_interval", 3, 80, 120, "uint8", 47, dict(corr_map="MATCH_HIST", channel="HLS",
                                                                 intensity_interval={"min": 5, "max": 250})),
]
VIGNETTE_CASES = [
    ("vig_u8_default", 2, 240, 360, "uint8", 51, dict()),
    ("vig_u16_default", 2, 200, 300, "uint16", 52, dict()),
    ("vig_u8_maxcorr_sub4", 2, 233, 301, "uint8", 53, dict(max_correction=0.6, subsample=4, black_threshold=12.0)),
]


def run_balance(frames, kwargs):
    import cv2
    from shinestacker.algorithms.balance import BalanceFrames
    with tempfile.TemporaryDirectory() as d:
        names = []
        for i, f in enumerate(frames):
            names.append(f"f{i:02d}.tif")
            cv2.imwrite(os.path.join(d, names[-1]), f, [int(cv2.IMWRITE_TIFF_COMPRESSION), 1])
        proc = MagicMock()
        proc.input_full_path, proc.filenames, proc.ref_idx, proc.counts = d, names, 0, len(frames)
        act = BalanceFrames(**kwargs)
        act.begin(proc)
        outs = [act.run_frame(i, 0, f.copy()) for i, f in enumerate(frames)]
        act.end()
        return np.array(outs), np.array(act.correction.corrections)


def run_vignette(frames, kwargs):
    from shinestacker.algorithms import vignetting as vg
    proc = MagicMock()
    proc.counts = len(frames)
    act = vg.Vignetting(**kwargs)
    act.begin(proc)
    outs, params = [], []
    for i, f in enumerate(frames):
        p = vg.compute_fit_parameters(f, act.r_steps, subsample=act.subsample, fast_subsampling=act.fast_subsampling)
        params.append(p)
        outs.append(act.run_frame(i, 0, f.copy()))
    return np.array(outs), np.array(params), np.array([np.asarray(c, dtype=float) for c in act.corrections])


def main():
    assert ref_loader.load() is not None, "reference not available"
    os.makedirs(OUT, exist_ok=True)
    for name, n, h, w, dtype, seed, kwargs in BALANCE_CASES:
        frames = make_frames(n, h, w, dtype, seed)
        outs, corr = run_balance(frames, kwargs)
        # frame 0 is the reference frame and passes through unchanged: only the corrected frames are stored
        assert np.array_equal(outs[0], frames[0])
        np.savez_compressed(os.path.join(OUT, name + ".npz"), out=outs[1:], corrections=corr, kwargs=np.array(repr(kwargs)),
                            recipe=np.array(repr((n, h, w, dtype, seed))))
        print(name, outs.shape, corr.ravel()[:4])
    for name, n, h, w, dtype, seed, kwargs in VIGNETTE_CASES:
        frames = make_frames(n, h, w, dtype, seed)
        outs, params, corr = run_vignette(frames, kwargs)
        np.savez_compressed(os.path.join(OUT, name + ".npz"), out=outs[:1], params=params, corrections=corr,
                            kwargs=np.array(repr(kwargs)), recipe=np.array(repr((n, h, w, dtype, seed))))
        print(name, outs.shape, params[0])


if __name__ == "__main__":
    main()
