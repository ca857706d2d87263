"""Generates the golden vectors in tests/golden/depthmap/ by running the UNMODIFIED reference
``DepthMapStack`` (lucalista/shinestacker, src/shinestacker/algorithms/depth_map.py, imported through
oracle/ref_harness.py) in the build container.  The reference's own test (tests/test_0061_depth_map.py)
only asserts shapes and dtypes, so these files are the pinned reference outputs that the oracle, the CPU
emulation of the kernels and the CUDA path are compared with where /root/reference does not exist.

    python tests/golden/make_golden_depthmap.py

Each .npz holds the constructor kwargs, the generator recipe (synthetic stacks are rebuilt from
shinestacker_b200.synthetic) or the input frames (crops of the reference's example stacks), the reference's
fused image and, as float16, the focus map of frame 0 (a coarse cross-check of the soft weights).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import ref_harness  # noqa: E402
from shinestacker_b200 import synthetic  # noqa: E402

OUT = os.path.join(HERE, "depthmap")

CASES = [
    # name, generator, n, h, w, dtype, seed, kwargs
    ("dm_noise_u8_4x120x161", "noise", 4, 120, 161, "uint8", 21, {}),
    ("dm_noise_u16_3x90x140", "noise", 3, 90, 140, "uint16", 22, {}),
    ("dm_noise_u8_3x90x140_max", "noise", 3, 90, 140, "uint8", 23, {"map_type": "max"}),
    ("dm_noise_u16_3x90x141_sobel_l4", "noise", 3, 90, 141, "uint16", 24, {"energy": "sobel", "levels": 4}),
    ("dm_noise_u8_3x64x64_nosmooth_l1", "noise", 3, 64, 64, "uint8", 25, {"smooth_size": 0, "levels": 1}),
    ("dm_noise_u8_3x77x93_b3_s5_l2", "noise", 3, 77, 93, "uint8", 26, {"blur_size": 3, "smooth_size": 5, "levels": 2}),
    ("dm_noise_u8_3x77x93_k7_b9", "noise", 3, 77, 93, "uint8", 27, {"kernel_size": 7, "blur_size": 9}),
    ("dm_syn_u8_6x200x264", "stack", 6, 200, 264, "uint8", 28, {}),
    ("dm_syn_u16_5x131x203_max_t05", "stack", 5, 131, 203, "uint16", 29, {"map_type": "max", "temperature": 0.5}),
    ("dm_noise_u8_32x48x72", "noise", 32, 48, 72, "uint8", 30, {}),
]


def make_frames(gen, n, h, w, dtype, seed):
    f = synthetic.make_stack if gen == "stack" else synthetic.make_noise_stack
    return f(n, h, w, np.dtype(dtype), seed)


def save(name, frames, kwargs, extra):
    out, rec = ref_harness.reference_depthmap_arrays(frames, **kwargs)
    assert np.isfinite(rec["weights"]).all()
    d = dict(fused=out, weights0=rec["weights"][0].astype(np.float16), kwargs=np.array(repr(kwargs)))
    d.update(extra)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **d)
    print(name, out.shape, out.dtype)


def main():
    os.makedirs(OUT, exist_ok=True)
    for name, gen, n, h, w, dtype, seed, kwargs in CASES:
        frames = make_frames(gen, n, h, w, dtype, seed)
        save(name, frames, kwargs, dict(recipe=np.array(repr((gen, n, h, w, dtype, seed)))))
    import cv2
    ex = "/root/reference/examples/input"
    for sub, flag, y0, x0, hh, ww, kw in (("img-jpg", cv2.IMREAD_COLOR, 500, 800, 192, 256, {}),
                                          ("img-tif", cv2.IMREAD_UNCHANGED, 520, 760, 150, 200, {"energy": "sobel"})):
        files = sorted(os.listdir(os.path.join(ex, sub)))
        frames = [np.ascontiguousarray(cv2.imread(os.path.join(ex, sub, f), flag)[y0:y0 + hh, x0:x0 + ww])
                  for f in files]
        save("dm_example_" + sub.replace("-", "_"), frames, kw, dict(frames=np.stack(frames)))


if __name__ == "__main__":
    main()
