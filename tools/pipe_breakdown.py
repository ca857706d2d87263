#!/usr/bin/env python
"""Host-side time per stage of align_and_stack_arrays on the config-5 workload (pinned frames)."""
import ctypes
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from shinestacker_b200 import synthetic, PyramidStack, _native   # noqa: E402
from shinestacker_b200 import pipeline, align                     # noqa: E402
from shinestacker_b200.align import Warper                        # noqa: E402
from tools import time_pipeline                                   # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
host = synthetic.make_stack(n, 3072, 4096, np.uint8, seed=2)
lib = _native.load()
frames, truth, pin = time_pipeline.displaced_stack(host, lib)
algo = PyramidStack(batch_frames=8)
wp = Warper((3072, 4096), np.uint8)
pipeline.align_and_stack_arrays(frames, stack_algo=algo, warper=wp)

acc, extra = {}, {}


def timed(name, fn):
    def w(*a, **k):
        t0 = time.perf_counter()
        r = fn(*a, **k)
        acc[name] = acc.get(name, 0.0) + (time.perf_counter() - t0) * 1e3
        return r
    return w


wp.use_prefetched = timed("use_prefetched", wp.use_prefetched)
wp.prefetch = timed("prefetch", wp.prefetch)
_ecc = wp.ecc_estimate


def ecc_and_kernel_time(*a, **k):
    r = _ecc(*a, **k)
    extra["ecc_device_ms_between_events"] = extra.get("ecc_device_ms_between_events", 0.0) + wp.last_kernel_ms
    return r


wp.ecc_estimate = timed("ecc_estimate", ecc_and_kernel_time)
wp.warp = timed("warp", wp.warp)
wp.result_device = timed("result_device", wp.result_device)
wp.ecc_set_reference = timed("ecc_set_reference", wp.ecc_set_reference)
algo._finish = timed("finish", algo._finish)


class L:
    def __init__(self, lib):
        self._l = lib

    def __getattr__(self, k):
        f = getattr(self._l, k)
        if k in ("ss_sync_copies", "ss_push_frame_device", "ss_push_frame"):
            return timed(k, f)
        return f


algo._lib = L(algo._lib)
t0 = time.perf_counter()
pipeline.align_and_stack_arrays(frames, stack_algo=algo, warper=wp)
total = (time.perf_counter() - t0) * 1e3
acc["total"] = total
acc["unaccounted"] = total - sum(v for k, v in acc.items() if k != "total")
acc.update(extra)
print(json.dumps({k: round(v, 2) for k, v in acc.items()}))
