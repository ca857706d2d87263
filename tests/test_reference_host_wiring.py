"""Drop-in check against the UNMODIFIED reference host objects (StackJob / FocusStack / FocusStackBunch,
reference stack.py:14-113, stack_framework.py:13-24): our PyramidStack is accepted as `stack_algo`, is
driven through exactly the protocol documented in INTEGRATION.md, and its result is written by the
reference's own write_img.  Needs the reference tree (build container only); the pixel arithmetic itself
is stubbed here because this box has no GPU -- the GPU parity tests cover it."""
import os

import numpy as np
import pytest

from oracle import ref_harness

pytestmark = pytest.mark.skipif(not ref_harness.reference_available(),
                                reason="reference tree not present on this box")


def test_focus_stack_and_bunch_drive_our_algo(tmp_path, monkeypatch):
    import cv2
    ref_harness.import_reference()
    from shinestacker.algorithms.stack_framework import StackJob
    from shinestacker.algorithms.stack import FocusStack, FocusStackBunch
    from shinestacker_b200 import PyramidStack, synthetic

    src = tmp_path / "proj" / "in"
    os.makedirs(src)
    frames = synthetic.make_stack(5, 64, 96, np.uint8, 3)
    for i, f in enumerate(frames):
        cv2.imwrite(str(src / f"img{i:02d}.tif"), f)

    calls = []

    def fake_focus_stack(self, filenames):
        # what the reference host hands over: absolute, sorted paths (stack.py:28-29)
        assert all(os.path.isabs(p) and os.path.isfile(p) for p in filenames)
        assert list(filenames) == sorted(filenames)
        assert self.process is not None and hasattr(self.process, "callback")
        calls.append(len(filenames))
        return cv2.imread(filenames[0], cv2.IMREAD_UNCHANGED)

    monkeypatch.setattr(PyramidStack, "focus_stack", fake_focus_stack)
    algo = PyramidStack()
    job = StackJob("job", str(tmp_path / "proj"), input_path="in", logger_name="test_b200")
    job.add_action(FocusStack("stack", algo, output_path="out", plot_stack=False))
    job.run()
    assert calls == [5]
    assert algo.do_step_callback is True and algo.process is not None
    out = [f for f in os.listdir(tmp_path / "proj" / "out") if f.endswith(".tif")]
    assert out == ["stack_img00.tif"]
    # bunches reuse one algo object for several sequential calls (stack.py:94-97)
    calls.clear()
    algo2 = PyramidStack()
    job2 = StackJob("job2", str(tmp_path / "proj"), input_path="in", logger_name="test_b200")
    job2.add_action(FocusStackBunch("bunch", algo2, output_path="bunches", frames=3, overlap=1, plot_stack=False))
    job2.run()
    assert calls == [3, 3] and algo2.do_step_callback is False
