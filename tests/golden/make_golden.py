"""Generates the golden vectors in tests/golden/ by running the UNMODIFIED reference
(lucalista/shinestacker, /root/reference/src, imported through oracle/ref_harness.py) in the build
container.  The reference itself stores no golden vectors for this path (its tests/test_0060_stack.py
only asserts "no exception"), so these files are the pinned reference outputs the oracle and the CUDA
path are compared with on boxes where /root/reference does not exist.

    python tests/golden/make_golden.py

Each .npz holds: the constructor kwargs, either the generator recipe (synthetic stacks are rebuilt from
shinestacker_b200.synthetic with the stored seed) or the input frames themselves (crops of the
reference's example stacks), and the reference outputs: fused image, per-level arg-max maps (int16),
base-level arg-max maps.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import ref_harness  # noqa: E402
from shinestacker_b200 import synthetic  # noqa: E402

CASES = [
    # name, generator, n, h, w, dtype, seed, kwargs
    ("syn_u8_5x300x407", "stack", 5, 300, 407, "uint8", 1, {}),
    ("syn_u16_4x260x376", "stack", 4, 260, 376, "uint16", 2, {}),
    ("syn_u8_8x512x512", "stack", 8, 512, 512, "uint8", 3, {}),
    ("noise_u8_3x131x200", "noise", 3, 131, 200, "uint8", 4, {}),
    ("noise_u16_3x140x133_k03", "noise", 3, 140, 133, "uint16", 5, {"gen_kernel": 0.3}),
    ("syn_u8_3x100x90_min8", "stack", 3, 100, 90, "uint8", 6, {"min_size": 8}),
    ("syn_u8_1x64x64", "stack", 1, 64, 64, "uint8", 7, {}),
    ("syn_u8_4x96x150_ks3", "stack", 4, 96, 150, "uint8", 8, {"kernel_size": 3}),
    ("noise_u8_6x257x385", "noise", 6, 257, 385, "uint8", 9, {}),
    ("syn_u8_2x40x50_nolevels", "stack", 2, 40, 50, "uint8", 10, {}),
    ("noise_u8_64x96x136", "noise", 64, 96, 136, "uint8", 11, {}),
    ("f64_noise_u8_4x150x210", "noise", 4, 150, 210, "uint8", 12, {"float_type": "float-64"}),
    ("f64_syn_u16_3x133x260", "stack", 3, 133, 260, "uint16", 13, {"float_type": "float-64"}),
    ("f64_syn_u8_2x40x50_nolevels", "stack", 2, 40, 50, "uint8", 14, {"float_type": "float-64", "kernel_size": 3}),
]


def make_frames(gen, n, h, w, dtype, seed):
    f = synthetic.make_stack if gen == "stack" else synthetic.make_noise_stack
    return f(n, h, w, np.dtype(dtype), seed)


def save(name, frames, kwargs, extra):
    out, info = ref_harness.reference_stack_arrays(frames, **kwargs)
    d = dict(fused=out, best_e=info["best_e"].astype(np.int16), best_d=info["best_d"].astype(np.int16),
             levels=np.int32(info["levels"]), kwargs=np.array(repr(kwargs)))
    for l, b in enumerate(info["best"]):
        d[f"best_{l}"] = b.astype(np.int16)
    d.update(extra)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **d)
    print(name, out.shape, out.dtype, "levels", info["levels"])


def main():
    for name, gen, n, h, w, dtype, seed, kwargs in CASES:
        frames = make_frames(gen, n, h, w, dtype, seed)
        save(name, frames, kwargs, dict(recipe=np.array(repr((gen, n, h, w, dtype, seed)))))
    # crops of the reference's own example stacks (inputs stored, the GPU box has no /root/reference)
    import cv2
    ex = "/root/reference/examples/input"
    for sub, flag, y0, x0, hh, ww in (("img-jpg", cv2.IMREAD_COLOR, 500, 800, 256, 384),
                                      ("img-tif", cv2.IMREAD_UNCHANGED, 520, 760, 200, 264),
                                      ("amber-flies", cv2.IMREAD_COLOR, 300, 500, 192, 272)):
        files = sorted(os.listdir(os.path.join(ex, sub)))
        frames = [np.ascontiguousarray(cv2.imread(os.path.join(ex, sub, f), flag)[y0:y0 + hh, x0:x0 + ww])
                  for f in files]
        save("example_" + sub.replace("-", "_"), frames, {}, dict(frames=np.stack(frames)))


if __name__ == "__main__":
    main()
