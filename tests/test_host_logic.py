"""Host-side logic of shinestacker_b200.PyramidStack.focus_stack (file loop, validation, callbacks,
cancellation, decode-ahead ordering) with the native calls stubbed out -- runs without a GPU.  The pixel
arithmetic itself is covered by the GPU parity tests."""
import numpy as np
import pytest

from shinestacker_b200 import PyramidStack, ShapeError, BitDepthError, ImageLoadError, RunStopException, synthetic


class Proc:
    id, name = 3, "stack"

    def __init__(self, stop_at=None):
        self.steps, self.msgs, self.stop_at = [], [], stop_at

    def sub_message_r(self, m):
        self.msgs.append(m)

    def callback(self, key, *a):
        if key == "after_step":
            self.steps.append(a[-1])
        if key == "check_running" and self.stop_at is not None and len(self.steps) >= self.stop_at:
            return False
        return None


@pytest.fixture
def stubbed(monkeypatch):
    pushed = []

    def ensure(self, shape, dtype, n, resident=False):
        self._h = None

    def push_slot(self, i, slot):
        # what the upload would read from the pinned slot the worker thread staged the frame in
        n = int(np.prod(self._frame_shape)) * np.dtype(self.dtype).itemsize
        pushed.append((i, slot[1][:n].view(self.dtype).reshape(self._frame_shape).copy()))

    def finish(self, n, shape):
        return np.zeros((shape[0], shape[1], 3), self.dtype)

    monkeypatch.setattr(PyramidStack, "_ensure_handle", ensure)
    monkeypatch.setattr(PyramidStack, "_staging",
                        lambda self, nbytes, count: [(None, np.zeros(nbytes, np.uint8)) for _ in range(4)])
    monkeypatch.setattr(PyramidStack, "_push_slot", push_slot)
    monkeypatch.setattr(PyramidStack, "_sync_copies", lambda self: None)
    monkeypatch.setattr(PyramidStack, "_finish", finish)
    return pushed


def _write(tmp_path, frames, ext="tif"):
    import cv2
    paths = []
    for i, f in enumerate(frames):
        p = str(tmp_path / f"f{i:03d}.{ext}")
        cv2.imwrite(p, f)
        paths.append(p)
    return paths


@pytest.mark.parametrize("threads", [0, 3])
def test_frames_are_pushed_in_order_with_two_steps_each(tmp_path, stubbed, threads):
    frames = synthetic.make_noise_stack(7, 40, 56, np.uint16, 1)
    paths = _write(tmp_path, frames)
    algo = PyramidStack(decode_threads=threads)
    algo.process = Proc()
    algo.do_step_callback = True
    out = algo.focus_stack(paths)
    assert out.shape == (40, 56, 3) and out.dtype == np.uint16
    assert [i for i, _ in stubbed] == list(range(7))
    for (_, img), f in zip(stubbed, frames):
        assert np.array_equal(img, f)
    # the reference's order (pyramid.py:155-177): every file is validated (steps 0..n-1) before any is processed
    assert algo.process.steps == list(range(14))                  # steps_per_frame = 2
    msgs = [m for m in algo.process.msgs if "file" in m]
    assert all("validating" in m for m in msgs[:7]) and all("processing" in m for m in msgs[7:])
    assert any("validating file f000.tif" in m for m in algo.process.msgs)
    assert any("processing file f006.tif" in m for m in algo.process.msgs)


def test_validation_errors_match_the_reference_types(tmp_path, stubbed):
    import cv2
    frames = synthetic.make_noise_stack(3, 40, 56, np.uint8, 2)
    paths = _write(tmp_path, frames)
    bad = str(tmp_path / "zz_bad.tif")
    cv2.imwrite(bad, np.zeros((41, 56, 3), np.uint8))
    with pytest.raises(ShapeError):
        PyramidStack().focus_stack(paths + [bad])
    cv2.imwrite(bad, np.zeros((40, 56, 3), np.uint16))
    with pytest.raises(BitDepthError):
        PyramidStack().focus_stack(paths + [bad])
    with pytest.raises(RuntimeError):                              # reference utils.py:12-13
        PyramidStack().focus_stack(paths + [str(tmp_path / "missing.tif")])
    unsupported = str(tmp_path / "zz.bmp")                         # read_img returns None for this extension
    cv2.imwrite(unsupported, frames[0])
    with pytest.raises(ImageLoadError):
        PyramidStack().focus_stack(paths + [unsupported])
    gray = str(tmp_path / "a_gray.tif")
    cv2.imwrite(gray, np.zeros((40, 56), np.uint8))
    with pytest.raises(ImageLoadError):
        PyramidStack().focus_stack([gray] + paths)


def test_shape_error_surfaces_before_any_frame_is_uploaded(tmp_path, stubbed):
    """reference pyramid.py:155-169: the validation pass covers ALL files before the first pyramid is built."""
    import cv2
    frames = synthetic.make_noise_stack(4, 40, 56, np.uint8, 5)
    paths = _write(tmp_path, frames)
    cv2.imwrite(paths[2], np.zeros((44, 56, 3), np.uint8))
    algo = PyramidStack()
    algo.process = Proc()
    algo.do_step_callback = True
    with pytest.raises(ShapeError):
        algo.focus_stack(paths)
    assert stubbed == [] and algo.process.steps == [0, 1]


def test_cancellation_raises_run_stop(tmp_path, stubbed):
    paths = _write(tmp_path, synthetic.make_noise_stack(5, 32, 48, np.uint8, 3))
    algo = PyramidStack()
    algo.process = Proc(stop_at=4)
    algo.do_step_callback = True
    with pytest.raises(RunStopException):
        algo.focus_stack(paths)
    assert len(stubbed) < 5
