"""Small depth-map stacks, warps and one ECC fit through the product path for compute-sanitizer."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from shinestacker_b200 import DepthMapStack, synthetic
from shinestacker_b200 import align as al
import align_util as au
# (3, 120, 168) and (10, 140, 200): interior tiles of the tiled energy kernel (four-pixel path), two blend batches;
# the other two: generic radii / scalar path, Sobel, direct bilateral
for dtype, shape, kw in ((np.uint8, (3, 120, 168), {}), (np.uint16, (2, 90, 131), {"map_type": "max", "energy": "sobel", "levels": 4}),
                         (np.uint8, (2, 70, 93), {"smooth_size": 5, "levels": 1}), (np.uint16, (10, 140, 200), {}),
                         (np.uint8, (2, 100, 136), {"kernel_size": 7, "blur_size": 3})):
    frames = synthetic.make_noise_stack(shape[0], shape[1], shape[2], dtype, 7)
    out = DepthMapStack(**kw).focus_stack_arrays(frames)
    print("depthmap", dtype.__name__, shape, int(out.astype(np.int64).sum()))
for dtype in (np.uint8, np.uint16):
    img = np.random.default_rng(3).integers(0, np.iinfo(dtype).max + 1, (150, 211, 3)).astype(dtype)
    wp = al.Warper(img.shape[:2], dtype)
    for name, m in au.TRANSFORMS:
        tr = al.ALIGN_HOMOGRAPHY if np.asarray(m).size == 9 else al.ALIGN_RIGID
        for mode in au.MODES:
            out = wp.warp(img, m, tr, mode)
    print("warp", dtype.__name__, int(out.astype(np.int64).sum()))
    wp.close()
img = au.textured_scene(300, 400, seed=3)
mov, _ = au.moved_frame(img, 0.3, 1.002, 4.0, -3.0)
wp = al.Warper(img.shape[:2], np.uint8)
wp.ecc_set_reference(img)
m, info = wp.ecc_estimate(mov)
print("ecc", np.round(m, 4).tolist(), info["iterations"])
# pipelined align + stack: prefetch behind the fit, double-buffered warp outputs, pageable frames through the staging ring
from shinestacker_b200.pipeline import align_and_stack_arrays
from tools.time_pipeline import displaced_stack
moved, _, _ = displaced_stack(synthetic.make_stack(6, 240, 320, np.uint8, 23))
out = align_and_stack_arrays(list(moved))
print("pipeline", int(out.astype(np.int64).sum()))
