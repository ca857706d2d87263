"""The product's fused level-kernel tile code (shinestacker_b200/csrc/ss_kernels.cuh) compiled for the
host (tests/emu/emu_level.cpp, -DSS_HOST_EMU) against the oracle: checks the halo / REFLECT101 fix-up /
tail-column index arithmetic on a CPU-only box.  The emulation is never used by the product."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from shinestacker_b200 import synthetic

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def emu():
    src = os.path.join(HERE, "emu", "emu_level.cpp")
    so = os.path.join(HERE, "emu", "libemu_level.so")
    hdr = os.path.join(HERE, "..", "shinestacker_b200", "csrc", "ss_kernels.cuh")
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.run(["g++", "-std=c++17", "-O1", "-mfma", "-ffp-contract=off", "-shared", "-fPIC",
                        "-o", so, src], check=True)
    return ctypes.CDLL(so)


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def run_levels(emu, oracle, frames, split=None):
    n = len(frames)
    h, w = frames[0].shape[:2]
    k = oracle.make_kernel(0.4)
    levels = oracle.num_levels(h, w)
    bps = frames[0].dtype.itemsize
    # level-0 frames with a 16-byte aligned row pitch, like the library's frame store
    pitch = (w * 3 * bps + 15) // 16 * 16
    cur = np.zeros((n, h, pitch), np.uint8)
    for i, f in enumerate(frames):
        cur[i, :, :w * 3 * bps] = f.reshape(h, -1).view(np.uint8)
    in_pitch = pitch
    res = []
    for l in range(levels):
        h2, w2 = (h + 1) // 2, (w + 1) // 2
        gp = (w2 + 3) // 4 * 4
        gnext = np.full((n, 3, h2, gp), np.nan, np.float32)
        be = np.full((h, w), np.nan, np.float32)
        bi = np.full((h, w), -1, np.int32)
        fu = np.full((3, h, w), np.nan, np.float32)
        batches = [(0, n)] if not split else [(a, min(a + split, n)) for a in range(0, n, split)]
        for a, b in batches:
            emu.emu_level(_p(cur[a:b]), bps if l == 0 else 4, in_pitch, b - a, h, w, _p(k), _p(gnext[a:b]), gp,
                          _p(be), _p(bi), _p(fu), a, int(a == 0))
        res.append((bi, fu))
        cur, h, w, in_pitch = gnext, h2, w2, gp * 4
    return res, cur[:, :, :, :w]


@pytest.mark.parametrize("gen,n,h,w,dtype,seed,split", [
    ("stack", 3, 128, 192, np.uint8, 1, None),
    ("stack", 3, 131, 203, np.uint8, 2, None),
    ("noise", 4, 97, 150, np.uint16, 3, 2),
    ("noise", 5, 300, 407, np.uint8, 4, 3),
    ("noise", 2, 70, 66, np.uint8, 6, None),
    ("noise", 3, 64, 375, np.uint8, 7, 1),
    ("noise", 2, 288, 512, np.uint8, 8, None),      # interior (fast-path) tiles at levels 0 and 1
    ("stack", 3, 300, 448, np.uint16, 9, 2),
    ("noise", 2, 384, 700, np.uint8, 10, None),     # interior tile + tail columns (700 % 8 == 4)
    ("stack", 5, 224, 320, np.uint8, 11, None),     # interior tiles, 5 frames through the prefetch pipeline
    ("stack", 4, 224, 320, np.uint8, 12, 3),        # ... and a second batch that continues the running state
])
def test_emulated_level_kernel_matches_oracle(emu, oracle, gen, n, h, w, dtype, seed, split):
    f = synthetic.make_stack if gen == "stack" else synthetic.make_noise_stack
    frames = f(n, h, w, dtype, seed)
    _, info = oracle.focus_stack(frames)
    res, base = run_levels(emu, oracle, frames, split)
    assert len(res) == info["levels"]
    for l, (bi, fu) in enumerate(res):
        assert np.array_equal(bi, info["best"][l]), f"level {l} arg-max"
        assert np.array_equal(fu.transpose(1, 2, 0), info["fused_pyramid"][l]), f"level {l} coefficients"
    gp = oracle.gaussian_pyramid(oracle.split_planes(frames[-1]), len(res), oracle.make_kernel(0.4))
    assert np.array_equal(gp[-1], base[-1])


def test_energy_stage_skip_triggers_and_is_exact(emu, oracle, monkeypatch):
    """The exact skip of the energy stage (ss_skip_factor): on a staircase stack the frames after the sharp one are
    skipped in the interior tiles, and the result is unchanged (bit-exact against the oracle, with and without)."""
    frames = synthetic.make_stack(6, 224, 384, np.uint8, 13)      # bands of 64 columns = one tile column each
    _, info = oracle.focus_stack(frames)
    emu.emu_skipped_threads.restype = ctypes.c_longlong
    emu.emu_skipped_threads(1)
    res, _ = run_levels(emu, oracle, frames)
    skipped = emu.emu_skipped_threads(1)
    assert skipped > 0 and skipped % 32 == 0                      # whole warps only
    for l, (bi, fu) in enumerate(res):
        assert np.array_equal(bi, info["best"][l]) and np.array_equal(fu.transpose(1, 2, 0), info["fused_pyramid"][l])
    monkeypatch.setenv("SS_EMU_NO_SKIP", "1")
    res2, _ = run_levels(emu, oracle, frames)
    assert emu.emu_skipped_threads(1) == 0
    for (bi, fu), (bi2, fu2) in zip(res, res2):
        assert np.array_equal(bi, bi2) and np.array_equal(fu, fu2)
