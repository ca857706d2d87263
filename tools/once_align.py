#!/usr/bin/env python
"""One warp of each kind on a 4096x3072 frame (for an ncu launch list)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from shinestacker_b200 import align as al  # noqa: E402

for dt in (np.uint8, np.uint16):
    img = np.random.default_rng(1).integers(0, np.iinfo(dt).max + 1, (3072, 4096, 3)).astype(dt)
    a = np.deg2rad(0.35)
    m = np.array([[1.002 * np.cos(a), -1.002 * np.sin(a), 18.4], [1.002 * np.sin(a), 1.002 * np.cos(a), -11.7]])
    wp = al.Warper(img.shape[:2], dt)
    wp.warp(img, m, al.ALIGN_RIGID, "BORDER_REPLICATE_BLUR")
    wp.warp(img, np.vstack([m, [1e-6, -2e-6, 1.0]]), al.ALIGN_HOMOGRAPHY, "BORDER_REPLICATE_BLUR")
    wp.close()
