"""The UNMODIFIED reference host (StackJob / FocusStack / FocusStackBunch from baseline/_ref, reference stack.py:14-113,
stack_framework.py:13-24) drives the GPU stacker end to end on a real device: files in, the reference's own write_img
out -- the path of the reference's tests/test_0060_stack.py:7-44 with `PyramidStack()` replaced by ours.  The written
image must equal the oracle's (and, for the plain job, what the reference's own PyramidStack writes)."""
import os

import numpy as np
import pytest

from tools import ref_loader

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(ref_loader.reference_root() is None, reason="baseline/_ref not installed")]
cv2 = pytest.importorskip("cv2")


def _write_stack(folder, frames, ext, params=()):
    os.makedirs(folder)
    for i, f in enumerate(frames):
        cv2.imwrite(os.path.join(folder, f"img{i:02d}.{ext}"), f, list(params))


@pytest.mark.parametrize("ext,dtype", [("tif", np.uint8), ("tif", np.uint16), ("jpg", np.uint8)])
def test_reference_stack_job_runs_the_gpu_algo(tmp_path, oracle, ext, dtype):
    ref_loader.load()
    from shinestacker.algorithms.pyramid import PyramidStack as RefPyramidStack
    from shinestacker_b200 import PyramidStack, synthetic
    frames = synthetic.make_stack(5, 200, 328, dtype, 4)
    work = str(tmp_path / "proj")
    _write_stack(os.path.join(work, "in"), frames, ext,
                 [int(cv2.IMWRITE_JPEG_QUALITY), 100] if ext == "jpg" else [int(cv2.IMWRITE_TIFF_COMPRESSION), 1])
    decoded = [cv2.imread(os.path.join(work, "in", f"img{i:02d}.{ext}"), cv2.IMREAD_UNCHANGED) for i in range(5)]
    algo = PyramidStack()
    ours = ref_loader.run_stack_job(work, "in", algo, output_subdir="out_gpu", name="gpu")
    theirs = ref_loader.run_stack_job(work, "in", RefPyramidStack(), output_subdir="out_ref", name="ref")
    assert os.path.basename(ours) == f"stack_img00.{ext}" and algo.do_step_callback is True
    expect, _ = oracle.focus_stack(decoded)
    img_ours = cv2.imread(ours, cv2.IMREAD_UNCHANGED)
    if ext == "tif":                                   # lossless container: the written file is the fused image
        assert np.array_equal(img_ours, expect)
    assert np.array_equal(img_ours, cv2.imread(theirs, cv2.IMREAD_UNCHANGED))


def test_reference_bunches_reuse_one_gpu_algo(tmp_path, oracle):
    ref_loader.load()
    from shinestacker.algorithms.stack_framework import StackJob
    from shinestacker.algorithms.stack import FocusStackBunch, get_bunches
    from shinestacker_b200 import PyramidStack, synthetic
    frames = synthetic.make_noise_stack(7, 96, 160, np.uint8, 8)
    work = str(tmp_path / "proj")
    _write_stack(os.path.join(work, "in"), frames, "tif", [int(cv2.IMWRITE_TIFF_COMPRESSION), 1])
    job = StackJob("bunches", work, input_path="in", logger_name="b200_ref_arm")
    job.add_action(FocusStackBunch("bunch", PyramidStack(), output_path="bunches", frames=3, overlap=1, plot_stack=False))
    job.run()
    names = sorted(f for f in os.listdir(os.path.join(work, "bunches")) if f.endswith(".tif"))
    chunks = get_bunches([f"img{i:02d}.tif" for i in range(7)], 3, 1)
    assert len(names) == len(chunks)
    for chunk in chunks:
        expect, _ = oracle.focus_stack([frames[int(n[3:5])] for n in chunk])
        got = cv2.imread(os.path.join(work, "bunches", "stack_" + chunk[0]), cv2.IMREAD_UNCHANGED)
        assert np.array_equal(got, expect)
