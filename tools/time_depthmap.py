#!/usr/bin/env python
"""Device time of the depth-map chain (DepthMapStack on the GPU) on a synthetic stack: frames resident in HBM
(ssdm_rerun_resident, CUDA events inside the library) and end to end from pinned host frames.

    python tools/time_depthmap.py [--frames 64] [--height 3072] [--width 4096] [--dtype uint8] [--steps 5]
"""
import argparse
import ctypes
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def measure(frames_n=64, height=3072, width=4096, dtype="uint8", steps=5, warmup=2, host=None, pinned=False, **kwargs):
    from shinestacker_b200 import DepthMapStack, synthetic, _native
    from shinestacker_b200.pyramid import _check
    lib = _native.load()
    dt = np.dtype(dtype)
    if host is None:
        sharp, blurred = synthetic.make_scene(height, width, dt, seed=2)
        maxv = 255 if dt == np.uint8 else 65535
        sharp = np.clip(sharp, 0, maxv).astype(dt)
        blurred = np.clip(blurred, 0, maxv).astype(dt)
        host = [synthetic.frame_from_scene(sharp, blurred, i, frames_n, dt) for i in range(frames_n)]
    algo = DepthMapStack(**kwargs)
    t0 = time.perf_counter()
    out = algo.focus_stack_arrays(host)          # includes handle creation (cudaMalloc of the resident stack)
    first = time.perf_counter() - t0
    h = algo._h
    ms = ctypes.c_float()
    for _ in range(warmup):
        _check(lib.ssdm_rerun_resident(h))
    dev, wall = [], []
    for _ in range(steps):
        t0 = time.perf_counter()
        _check(lib.ssdm_rerun_resident(h))
        wall.append(time.perf_counter() - t0)
        _check(lib.ssdm_kernel_time_ms(h, ctypes.byref(ms)))
        dev.append(ms.value)
    nl = ctypes.c_int64()
    _check(lib.ssdm_launch_count(h, ctypes.byref(nl)))
    again = np.empty_like(out)
    _check(lib.ssdm_finish_end(h, again.ctypes.data_as(ctypes.c_void_p), again.strides[0]))
    e2e = []
    for _ in range(max(2, steps // 2)):
        t0 = time.perf_counter()
        algo.focus_stack_arrays(host, stage=not pinned)   # pinned frames are uploaded from where they are
        e2e.append(time.perf_counter() - t0)
    mp = frames_n * height * width / 1e6
    return {"workload": f"DepthMapStack {frames_n} x {width}x{height} {dtype} {kwargs}",
            "device_ms_per_stack": float(np.median(dev)), "wall_ms_per_stack": float(np.median(wall) * 1e3),
            "resident_mp_per_s": mp / (float(np.median(wall))), "launches_per_stack": int(nl.value),
            "e2e_ms_per_stack": float(np.median(e2e) * 1e3), "e2e_mp_per_s": mp / float(np.median(e2e)),
            "e2e_frames_pinned": bool(pinned), "first_call_ms": first * 1e3, "rerun_matches_first_result": bool(np.array_equal(out, again))}


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=64)
    ap.add_argument("--height", type=int, default=3072)
    ap.add_argument("--width", type=int, default=4096)
    ap.add_argument("--dtype", default="uint8")
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--once", action="store_true", help="one stack only (for an ncu launch list)")
    a = ap.parse_args()
    if a.once:
        from shinestacker_b200 import DepthMapStack, synthetic
        fr = synthetic.make_stack(a.frames, a.height, a.width, np.dtype(a.dtype), seed=2)
        algo = DepthMapStack()
        algo.focus_stack_arrays(fr)
        print(json.dumps({"device_ms": algo.last_kernel_ms, "launches": algo.last_launches}))
        sys.exit(0)
    print(json.dumps(measure(a.frames, a.height, a.width, a.dtype, a.steps)))
