#!/usr/bin/env python
"""Device time of the AlignFrames warp stage (warp + mask + blur-fill) on one 4096x3072 frame, next to cv2 on the
host cores.

    python tools/time_align.py [--height 3072] [--width 4096] [--dtype uint8] [--steps 10]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def measure(height=3072, width=4096, dtype="uint8", steps=10):
    import cv2
    from shinestacker_b200 import align as al
    dt = np.dtype(dtype)
    img = np.random.default_rng(1).integers(0, np.iinfo(dt).max + 1, (height, width, 3)).astype(dt)
    a = np.deg2rad(0.35)
    m = np.array([[1.002 * np.cos(a), -1.002 * np.sin(a), 18.4], [1.002 * np.sin(a), 1.002 * np.cos(a), -11.7]])
    wp = al.Warper(img.shape[:2], dt)
    res = {}
    for mode in ("BORDER_REPLICATE", "BORDER_REPLICATE_BLUR"):
        wp.warp(img, m, al.ALIGN_RIGID, mode)
        dev, wall = [], []
        for _ in range(steps):
            t0 = time.perf_counter()
            wp.warp(img, m, al.ALIGN_RIGID, mode)
            wall.append(time.perf_counter() - t0)
            dev.append(wp.last_kernel_ms)
        res[mode] = {"device_ms": float(np.median(dev)), "host_call_ms": float(np.median(wall)) * 1e3}
    t0 = time.perf_counter()
    n_cpu = 3
    for _ in range(n_cpu):
        out = cv2.warpAffine(img, m, (width, height), borderMode=cv2.BORDER_REPLICATE)
        mask = cv2.cvtColor(cv2.warpAffine(np.ones_like(img, dtype=np.uint8), m, (width, height),
                                           borderMode=cv2.BORDER_CONSTANT, borderValue=0), cv2.COLOR_BGR2GRAY)
        bl = cv2.GaussianBlur(out, (21, 21), sigmaX=50)
        out[mask == 0] = bl[mask == 0]
    res["cv2_cpu_ms"] = (time.perf_counter() - t0) / n_cpu * 1e3
    res["cpu_threads"] = cv2.getNumThreads()
    res["workload"] = f"warp + mask + 21x21 blur fill, {width}x{height} {dtype}"
    res["megapixels"] = height * width / 1e6
    return res


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--height", type=int, default=3072)
    ap.add_argument("--width", type=int, default=4096)
    ap.add_argument("--dtype", default="uint8")
    ap.add_argument("--steps", type=int, default=10)
    a = ap.parse_args()
    print(json.dumps(measure(a.height, a.width, a.dtype, a.steps)))
