"""Loader for the golden vectors in tests/golden (see make_golden.py)."""
import ast
import glob
import os

import numpy as np

from shinestacker_b200 import synthetic

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def case_names():
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "*.npz")))


def load_case(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    kwargs = ast.literal_eval(str(z["kwargs"]))
    if "frames" in z.files:
        frames = [np.ascontiguousarray(f) for f in z["frames"]]
    else:
        gen, n, h, w, dtype, seed = ast.literal_eval(str(z["recipe"]))
        f = synthetic.make_stack if gen == "stack" else synthetic.make_noise_stack
        frames = f(n, h, w, np.dtype(dtype), seed)
    levels = int(z["levels"])
    expect = dict(fused=z["fused"], best=[z[f"best_{l}"].astype(np.int32) for l in range(levels)],
                  best_e=z["best_e"].astype(np.int32), best_d=z["best_d"].astype(np.int32), levels=levels)
    return frames, kwargs, expect


def assert_matches(out, maps, expect, what):
    assert out.dtype == expect["fused"].dtype and out.shape == expect["fused"].shape
    assert maps["levels"] == expect["levels"], f"{what}: level count"
    for l, (a, b) in enumerate(zip(maps["best"], expect["best"])):
        assert np.array_equal(a, b), f"{what}: arg-max map of level {l}: {(a != b).sum()} mismatches"
    assert np.array_equal(maps["best_e"], expect["best_e"]), f"{what}: base entropy arg-max"
    assert np.array_equal(maps["best_d"], expect["best_d"]), f"{what}: base deviation arg-max"
    # bit-exact is the bar; the north-star's tolerance for the fused image is 1 LSB per channel
    diff = np.abs(out.astype(np.int64) - expect["fused"].astype(np.int64))
    assert diff.max() == 0, f"{what}: fused image differs in {(diff > 0).sum()} samples (max {diff.max()})"
