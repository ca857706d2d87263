#!/usr/bin/env python
"""T_e2e of SURVEY.md section 8d: TIFF files -> fused image through PyramidStack.focus_stack (decode + upload +
kernels + download), files in tmpfs, with the mapped-TIFF fast path (shinestacker_b200/fastio.py) and with
cv2.imread decoding only."""
import json
import os
import shutil
import sys
import tempfile
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def measure(host, steps=2, threads=None, sweep=()):
    import cv2
    from shinestacker_b200 import PyramidStack, fastio
    from shinestacker_b200.pyramid import default_decode_threads
    threads = default_decode_threads() if threads is None else threads
    need = int(sum(f.nbytes for f in host) * 1.1) + (64 << 20)
    root = None
    for cand in ("/dev/shm", tempfile.gettempdir()):
        try:
            if os.path.isdir(cand) and shutil.disk_usage(cand).free > need:
                root = cand
                break
        except OSError:
            pass
    if root is None:
        raise RuntimeError(f"no scratch directory with {need >> 20} MiB free")
    d = tempfile.mkdtemp(prefix="ssb200_", dir=root)
    try:
        paths = []
        for i, f in enumerate(host):
            p = os.path.join(d, f"{i:04d}.tif")
            if not cv2.imwrite(p, f, [int(cv2.IMWRITE_TIFF_COMPRESSION), 1]):   # what the reference's write_img does
                raise RuntimeError(f"cannot write {p}")
            paths.append(p)
        n, h, w = len(host), host[0].shape[0], host[0].shape[1]
        res = {"workload": f"{n} uncompressed TIFFs {w}x{h} {host[0].dtype} in {root} -> fused image, "
                           f"PyramidStack.focus_stack, {threads} decode threads"}
        outs = {}
        for label, enabled in (("mapped_tiff", True), ("cv2_imread", False)):
            fastio.ENABLED = enabled
            algo = PyramidStack(decode_threads=threads)
            algo.focus_stack(paths)                                         # warm-up (handle, page cache)
            ts = []
            for _ in range(steps):
                t0 = time.perf_counter()
                outs[label] = algo.focus_stack(paths)
                ts.append(time.perf_counter() - t0)
            algo.close()
            t = float(np.median(ts))
            res[label] = {"ms_per_stack": t * 1e3, "mp_per_s": n * h * w / 1e6 / t}
        fastio.ENABLED = True
        res["same_result"] = bool(np.array_equal(outs["mapped_tiff"], outs["cv2_imread"]))
        if sweep:       # worker threads / frames per kernel-chain launch (the pinned ring has 2 x batch slots)
            res["sweep_mapped_tiff_ms"] = {}
            for th, batch in sweep:
                algo = PyramidStack(decode_threads=th, batch_frames=batch)
                algo.focus_stack(paths)
                ts = []
                for _ in range(steps):
                    t0 = time.perf_counter()
                    algo.focus_stack(paths)
                    ts.append(time.perf_counter() - t0)
                algo.close()
                res["sweep_mapped_tiff_ms"][f"{th} threads, batch {batch}"] = float(np.median(ts)) * 1e3
        # the in-memory entry point: pageable numpy frames -> fused image (worker threads stage them into pinned memory)
        algo = PyramidStack(decode_threads=threads)
        frames = [np.array(f) for f in host]
        algo.focus_stack_arrays(frames)
        ts = []
        for _ in range(steps):
            t0 = time.perf_counter()
            out = algo.focus_stack_arrays(frames)
            ts.append(time.perf_counter() - t0)
        algo.close()
        t = float(np.median(ts))
        res["arrays"] = {"ms_per_stack": t * 1e3, "mp_per_s": n * h * w / 1e6 / t,
                         "same_result": bool(np.array_equal(out, outs["mapped_tiff"]))}
        return res
    finally:
        shutil.rmtree(d, ignore_errors=True)


if __name__ == "__main__":
    from shinestacker_b200 import synthetic
    n, h, w = (int(v) for v in (sys.argv[1:4] + [16, 1536, 2048][len(sys.argv) - 1:])[:3])
    thr = int(sys.argv[4]) if len(sys.argv) > 4 else 6
    print(json.dumps(measure(synthetic.make_stack(n, h, w, np.uint8, seed=2), steps=3, threads=thr)))
