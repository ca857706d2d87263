"""BalanceFrames / Vignetting on the GPU, through ctypes -> the sssub_* C ABI: the table
look-up is bit-exact against the oracle and the reference's goldens; the radial divide agrees within 1 LSB (float64
exp of CUDA vs the host's; tolerance stated below); the unmodified reference's CombinedActions drives both sub-actions;
a corrected frame can be handed to the stacker without leaving HBM."""
import ctypes
import os

import numpy as np
import pytest

from oracle import subaction_oracle as so
from shinestacker_b200 import BalanceFrames, Vignetting, PyramidStack, subactions, synthetic, vignetting as vig
from shinestacker_b200.pyramid import _check
from tools import ref_loader

import subaction_util as su

pytestmark = pytest.mark.gpu
cv2 = pytest.importorskip("cv2")


@pytest.mark.parametrize("dtype", [np.uint8, np.uint16])
@pytest.mark.parametrize("shape", [(37, 53), (200, 328), (1, 5), (515, 750)])
@pytest.mark.parametrize("n_lut", [1, 3])
def test_table_look_up_is_bit_exact(dtype, shape, n_lut):
    rng = np.random.default_rng(shape[0] + n_lut)
    nv = 256 if dtype == np.uint8 else 65536
    img = rng.integers(0, nv, shape + (3,)).astype(dtype)
    luts = rng.integers(0, nv, (n_lut, nv)).astype(dtype)
    f = subactions.DeviceFrame(shape, dtype)
    out = f.upload(img).apply_lut(luts).download()
    assert np.array_equal(out, so.apply_lut(img, luts)) and f.launches == 1
    f.close()


@pytest.mark.parametrize("name", su.cases("bal_"))
def test_balance_frames_reproduces_the_reference(tmp_path, name):
    d, kwargs, frames = su.load(name)
    names = []
    for i, f in enumerate(frames):
        names.append(f"f{i}.tif")
        cv2.imwrite(str(tmp_path / names[-1]), f, [int(cv2.IMWRITE_TIFF_COMPRESSION), 1])
    act = BalanceFrames(**kwargs)
    act.begin(su.Proc(str(tmp_path), names))
    outs = [act.run_frame(i, 0, f.copy()) for i, f in enumerate(frames)]
    act.end()
    assert np.array_equal(outs[0], frames[0])                       # the reference frame passes through
    assert np.array_equal(np.array(outs[1:]), d["out"])
    assert np.allclose(act.corrections, d["corrections"], rtol=1e-12, atol=0)


def _close_to_1_lsb(got, expect):
    diff = np.abs(got.astype(np.int64) - expect.astype(np.int64))
    assert diff.max() <= 1, f"max difference {diff.max()} LSB"
    assert (diff != 0).mean() < 1e-4, f"{(diff != 0).mean():.2e} of the samples differ"


@pytest.mark.parametrize("name", su.cases("vig_"))
def test_vignetting_reproduces_the_reference_within_1_lsb(name):
    d, kwargs, frames = su.load(name)
    act = Vignetting(**kwargs)
    act.begin(su.Proc("", ["a"] * len(frames)))
    out = act.run_frame(0, 0, frames[0].copy())
    assert np.allclose(act.last_params, d["params"][0], rtol=1e-9, atol=0)
    assert np.allclose(np.array(act.corrections, dtype=float)[:, 0], d["corrections"][:, 0], rtol=1e-6, atol=1e-9)
    _close_to_1_lsb(out, d["out"][0])


@pytest.mark.parametrize("dtype,kw", [(np.uint8, {}), (np.uint16, dict(max_correction=0.5, black_threshold=20.0))])
def test_vignette_divide_against_the_oracle_full_frame(dtype, kw):
    f = su.make_frames(1, 3072, 4096, dtype, 79)[0]
    p = np.array([300.0 if dtype == np.uint8 else 70000.0, 1.4e-3, 2300.0])
    v0 = vig.sigmoid_model(0, *p)
    frame = subactions.DeviceFrame(f.shape[:2], dtype)
    out = frame.upload(f).apply_vignette(p, v0, kw.get("max_correction", 1.0), kw.get("black_threshold", 1.0)).download()
    _close_to_1_lsb(out, so.correct_vignette(f, p, v0, kw.get("max_correction", 1.0), kw.get("black_threshold", 1.0)))
    assert frame.kernel_ms < 5.0
    frame.close()


@pytest.mark.skipif(ref_loader.reference_root() is None, reason="baseline/_ref not installed")
def test_reference_combined_actions_drive_both_sub_actions(tmp_path):
    """reference stack_framework.py:246-302: CombinedActions reads each frame, runs the sub-actions, writes it."""
    ref_loader.load()
    import logging
    from shinestacker.algorithms.stack_framework import StackJob, CombinedActions
    from shinestacker.algorithms.balance import BalanceFrames as RefBalance
    from shinestacker.algorithms.vignetting import Vignetting as RefVignetting
    for lname in (None, "tqdm"):
        if not logging.getLogger(lname).handlers:
            logging.getLogger(lname).addHandler(logging.NullHandler())
    frames = su.make_frames(4, 160, 240, np.uint8, 80)
    work = str(tmp_path / "proj")
    os.makedirs(os.path.join(work, "in"))
    for i, f in enumerate(frames):
        cv2.imwrite(os.path.join(work, "in", f"img{i:02d}.tif"), f, [int(cv2.IMWRITE_TIFF_COMPRESSION), 1])

    def run(actions, out):
        job = StackJob("sub", work, input_path="in", logger_name="b200_ref_arm")
        job.add_action(CombinedActions("combo", actions, output_path=out))
        job.run()
        return [cv2.imread(os.path.join(work, out, f"img{i:02d}.tif"), cv2.IMREAD_UNCHANGED) for i in range(4)]

    ours = run([Vignetting(), BalanceFrames(channel="RGB")], "ours")
    theirs = run([RefVignetting(), RefBalance(channel="RGB")], "theirs")
    for a, b in zip(ours, theirs):
        # vignetting within 1 LSB; the balance tables that follow are fitted on those frames, so a sample may
        # move by the local slope of the table as well
        assert np.abs(a.astype(int) - b.astype(int)).max() <= 2
        assert (a != b).mean() < 1e-3


def test_corrected_frame_goes_to_the_stacker_without_leaving_hbm(oracle):
    frames = su.make_frames(4, 200, 328, np.uint8, 81)
    act = BalanceFrames(channel="RGB", subsample=2)
    act.begin_with_reference(frames[0], len(frames))
    balanced = [frames[0]] + [so.apply_lut(f, act.look_up_tables(f)) for f in frames[1:]]
    expect, _ = oracle.focus_stack(balanced)
    algo = PyramidStack(batch_frames=2)
    algo._set_metadata(frames[0])
    algo._ensure_handle(frames[0].shape, frames[0].dtype, len(frames))
    dev = subactions.DeviceFrame(frames[0].shape[:2], frames[0].dtype)
    for i, f in enumerate(frames):
        dev.upload(f)
        if i:
            dev.apply_lut(act.look_up_tables(f))
        d_ptr, pitch = dev.device_frame()
        _check(algo._lib.ss_push_frame_device(algo._h, d_ptr, pitch))
        _check(algo._lib.ss_sync_copies(algo._h))              # the device frame is reused for the next upload
    out = algo._finish(len(frames), frames[0].shape[:2])
    assert np.array_equal(out, expect)
    dev.close()
