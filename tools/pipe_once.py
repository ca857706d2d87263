#!/usr/bin/env python
"""Probes of the ECC fit inside the align + stack pipeline (config 5) on a few 4096x3072 frames:
  1. host time of one fit + warp,
  2. cost and accuracy of the fit stopped at each pyramid level (finest_level = 0 .. 6),
  3. device time of the fit with and without an upload of the next frame in flight.
usage: python tools/pipe_once.py [n_frames]      (the numbers quoted in DESIGN.md section 11)"""
import sys, time, json
sys.path.insert(0, ".")
import numpy as np
from shinestacker_b200 import synthetic, PyramidStack, _native
from shinestacker_b200.align import Warper
from shinestacker_b200.pipeline import align_and_stack_arrays
from tools import time_pipeline
n = int(sys.argv[1]) if len(sys.argv) > 1 else 6
host = synthetic.make_stack(n, 3072, 4096, np.uint8, seed=2)
lib = _native.load()
frames, truth, pin = time_pipeline.displaced_stack(host, lib)
algo = PyramidStack(batch_frames=8)
wp = Warper((3072, 4096), np.uint8)
align_and_stack_arrays(frames, stack_algo=algo, warper=wp)
# per-call host timing of one moving frame
import ctypes
wp.ecc_set_reference(frames[n // 2])
t = {}
for rep in range(3):
    t0 = time.perf_counter(); wp.prefetch(frames[0]); wp.use_prefetched(); m, info = wp.ecc_estimate(None); t1 = time.perf_counter()
    wp.warp(None, m, download=False); t2 = time.perf_counter()
    t = dict(ecc_ms=(t1 - t0) * 1e3, warp_ms=(t2 - t1) * 1e3, iters=info["iterations"], kernel_ms=wp.last_kernel_ms)
print(json.dumps(t))
# cost per pyramid level: the same fit stopped at coarser finest levels
for fl in (0, 1, 2, 3, 4, 5, 6):
    ts = []
    for rep in range(3):
        wp.prefetch(frames[0]); wp.use_prefetched()
        t0 = time.perf_counter(); m, info = wp.ecc_estimate(None, fl); ts.append((time.perf_counter() - t0) * 1e3)
    c = np.array([2048.0, 1536.0, 1.0])
    print("finest", fl, "ms", round(min(ts), 3), "kernel_ms", round(wp.last_kernel_ms, 3), "iters", info["iterations"],
          "err_px", float(np.abs(m @ c - truth[0] @ c).max()))
# does an upload in flight slow the fit down?
for concurrent in (0, 1):
    ks, ws = [], []
    for rep in range(6):
        wp.prefetch(frames[0]); wp.use_prefetched()
        lib.ss_sync_copies(algo._h)
        import ctypes as _c
        time.sleep(0.003)                       # let the upload of frames[0] finish
        if concurrent:
            wp.prefetch(frames[1])
        t0 = time.perf_counter(); m, info = wp.ecc_estimate(None, 1); ws.append((time.perf_counter() - t0) * 1e3)
        ks.append(wp.last_kernel_ms)
        if concurrent:
            wp.use_prefetched()
    print("concurrent upload", concurrent, "wall ms", round(min(ws), 3), "device ms", round(min(ks), 3), round(float(np.median(ks)), 3))
