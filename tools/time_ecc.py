#!/usr/bin/env python
"""Time of the ECC estimator + warp for one 4096x3072 frame on the GPU, next to the reference's feature pipeline
(SIFT + FLANN + RANSAC on the 2x sub-sampled frames, cv2, host cores)."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))


def measure(height=3072, width=4096, steps=5, with_features=True):
    import align_util as au
    from shinestacker_b200 import align as al
    img = au.textured_scene(height, width, seed=31)
    mov, m_true = au.moved_frame(img, -0.45, 1.006, 17.0, -12.0)
    wp = al.Warper(img.shape[:2], np.uint8)
    wp.ecc_set_reference(img)
    wp.ecc_estimate(mov)
    dev, wall, its = [], [], []
    for _ in range(steps):
        t0 = time.perf_counter()
        m, info = wp.ecc_estimate(mov)
        wall.append(time.perf_counter() - t0)
        dev.append(wp.last_kernel_ms)
        its.append(info["iterations"])
    t0 = time.perf_counter()
    wp.warp(None, m, al.ALIGN_RIGID, "BORDER_REPLICATE_BLUR")
    warp_ms = (time.perf_counter() - t0) * 1e3
    res = {"workload": f"ECC estimate of one {width}x{height} uint8 frame (similarity, full resolution)",
           "device_ms_incl_upload": float(np.median(dev)), "host_call_ms": float(np.median(wall)) * 1e3,
           "iterations": int(np.median(its)), "rho": info["rho"],
           "error_vs_truth": dict(zip(("angle_deg", "scale", "centre_px"), au.transform_error(m, m_true, img.shape))),
           "warp_of_resident_frame_host_call_ms": warp_ms}
    if with_features:
        t0 = time.perf_counter()
        n, m_f, _ = al.align_images(img, mov)
        res["reference_feature_pipeline_ms"] = (time.perf_counter() - t0) * 1e3
        res["feature_error_vs_truth"] = dict(zip(("angle_deg", "scale", "centre_px"),
                                                 au.transform_error(np.float64(m_f), m_true, img.shape)))
        res["good_matches"] = n
    return res


if __name__ == "__main__":
    print(json.dumps(measure()))
