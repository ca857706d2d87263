#!/usr/bin/env python
"""BASELINE.json config 5 ("AlignFrames ECC + FocusStack end-to-end, 4096x3072 RGB8, 64 frames, 1xB200"):
every frame of the synthetic stack is displaced by a known small similarity (rotation <= 0.5 deg, shift <= 20 px,
scale 1 +- 0.01 i/N), then aligned (ECC fit + warp on the GPU) and fused
(shinestacker_b200.pipeline.align_and_stack_arrays).  Timed end to end from pinned host frames to the fused image
on the host."""
import ctypes
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def displaced_stack(host, lib=None):
    """Frames displaced by known similarities, written into pinned host memory when lib is given."""
    import cv2
    n, h, w = len(host), host[0].shape[0], host[0].shape[1]
    nbytes = host[0].nbytes
    if lib is not None:
        from shinestacker_b200.pyramid import _check
        pin = ctypes.c_void_p()
        _check(lib.ss_host_alloc(ctypes.byref(pin), n * nbytes))
        buf = np.frombuffer((ctypes.c_uint8 * (n * nbytes)).from_address(pin.value), dtype=host[0].dtype)
        out = buf.reshape((n,) + host[0].shape)
    else:
        pin, out = None, np.empty((n,) + host[0].shape, host[0].dtype)
    rng = np.random.default_rng(5)
    truth = []
    for i in range(n):
        if i == n // 2:
            out[i] = host[i]
            truth.append(np.array([[1.0, 0, 0], [0, 1.0, 0]]))
            continue
        ang, sc = rng.uniform(-0.5, 0.5), 1 + 0.01 * (i - n // 2) / n
        tx, ty = rng.uniform(-20, 20, 2)
        t = cv2.getRotationMatrix2D((w / 2.0, h / 2.0), ang, sc)
        t[:, 2] += (tx, ty)
        cv2.warpAffine(host[i], t, (w, h), dst=out[i], borderMode=cv2.BORDER_REPLICATE)
        truth.append(cv2.invertAffineTransform(t))
    return out, truth, pin


def measure(host, steps=2):
    from shinestacker_b200 import PyramidStack, _native
    from shinestacker_b200.align import Warper
    from shinestacker_b200.pipeline import align_and_stack_arrays
    lib = _native.load()
    frames, truth, pin = displaced_stack(host, lib)
    n, h, w = len(host), host[0].shape[0], host[0].shape[1]
    algo = PyramidStack(batch_frames=8)
    wp = Warper((h, w), host[0].dtype)
    out, ms = align_and_stack_arrays(frames, stack_algo=algo, warper=wp, return_transforms=True)   # warm-up
    times = []
    for _ in range(steps):
        t0 = time.perf_counter()
        out = align_and_stack_arrays(frames, stack_algo=algo, warper=wp)
        times.append(time.perf_counter() - t0)
    c = np.array([w / 2.0, h / 2.0, 1.0])
    err = max(float(np.abs(m @ c - t @ c).max()) for m, t in zip(ms, truth))
    t = float(np.median(times))
    res = {"workload": f"config 5: {n} x {w}x{h} {host[0].dtype} frames displaced by known similarities; ECC fit + warp + "
                       "pyramid fusion, pinned host frames -> fused image on the host",
           "e2e_ms_per_stack": t * 1e3, "e2e_mp_per_s": n * h * w / 1e6 / t,
           "max_centre_error_px_vs_known_transform": err,
           "h2d_bytes_per_stack": int(n * host[0].nbytes), "d2h_bytes_per_stack": int(host[0].nbytes)}
    wp.close()
    algo.close()
    lib.ss_host_free(pin)
    return res


if __name__ == "__main__":
    from shinestacker_b200 import synthetic
    n, h, w = (int(v) for v in (sys.argv[1:4] + [16, 1536, 2048][len(sys.argv) - 1:])[:3])
    print(json.dumps(measure(synthetic.make_stack(n, h, w, np.uint8, seed=2))))
