"""Small stack through the product path for compute-sanitizer (memcheck / racecheck / initcheck)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from shinestacker_b200 import PyramidStack, synthetic
for dtype, shape in ((np.uint8, (3, 200, 328)), (np.uint16, (2, 150, 210)), (np.uint8, (2, 288, 512))):
    frames = synthetic.make_noise_stack(shape[0], shape[1], shape[2], dtype, 7)
    out = PyramidStack(batch_frames=2).focus_stack_arrays(frames)
    print(dtype.__name__, shape, int(out.astype(np.int64).sum()))
out = PyramidStack(float_type="float-64").focus_stack_arrays(synthetic.make_noise_stack(2, 100, 140, np.uint8, 8))
print("f64", int(out.astype(np.int64).sum()))
