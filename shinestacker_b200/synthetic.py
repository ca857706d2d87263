"""Deterministic synthetic focus stacks (SURVEY.md §8d generator) shared by tests and bench.py.

Frame ``i`` of ``n`` is a blurred copy of a common scene that is kept sharp only inside the vertical
band ``x in [i*W/n, (i+1)*W/n)`` (so the depth map is a staircase), plus a constant patch in the
top-left corner of every frame (exact energy ties -> exercises the first-index rule of
``np.argmax``, reference pyramid.py:51).  Pure numpy: identical bytes on every box.
"""
import numpy as np


def _box_blur(img, radius, passes):
    """Separable running-mean blur with edge replication (approximates a Gaussian)."""
    out = img
    k = 2 * radius + 1
    for _ in range(passes):
        for axis in (0, 1):
            pad = [(0, 0)] * out.ndim
            pad[axis] = (radius + 1, radius)
            c = np.cumsum(np.pad(out, pad, mode="edge"), axis=axis, dtype=np.float64)
            n = out.shape[axis]
            hi = np.take(c, np.arange(k, k + n), axis=axis)
            lo = np.take(c, np.arange(0, n), axis=axis)
            out = ((hi - lo) / k).astype(np.float32)
    return out


def make_scene(h, w, dtype=np.uint8, seed=0):
    """Returns (sharp, blurred) float32 HxWx3 scenes in [0, MAX]."""
    maxv = 255.0 if np.dtype(dtype) == np.uint8 else 65535.0
    rng = np.random.default_rng(1000 + seed)
    sharp = _box_blur(rng.random((h, w, 3), dtype=np.float32) * maxv, 1, 1)
    # stretch contrast back after the blur so the sharp band carries clear detail
    sharp = np.clip((sharp - maxv / 2) * 2.0 + maxv / 2, 0, maxv).astype(np.float32)
    blurred = _box_blur(sharp, 2, 2)
    return sharp, blurred


def frame_from_scene(sharp, blurred, i, n, dtype=np.uint8):
    h, w, _ = sharp.shape
    maxv = 255 if np.dtype(dtype) == np.uint8 else 65535
    f = blurred.copy()
    x0, x1 = (i * w) // n, ((i + 1) * w) // n
    f[:, x0:x1] = sharp[:, x0:x1]
    ph, pw = min(20, h // 4), min(20, w // 4)
    f[:ph, :pw] = maxv // 3
    return np.clip(f, 0, maxv).astype(dtype)


def make_stack(n, h, w, dtype=np.uint8, seed=0):
    """List of ``n`` HxWx3 BGR frames of the given dtype."""
    sharp, blurred = make_scene(h, w, dtype, seed)
    # quantise the two scenes once; every frame is then a byte-level composition of them (identical
    # to frame_from_scene on the float scenes, clip/astype being elementwise)
    maxv = 255 if np.dtype(dtype) == np.uint8 else 65535
    sharp = np.clip(sharp, 0, maxv).astype(dtype)
    blurred = np.clip(blurred, 0, maxv).astype(dtype)
    return [frame_from_scene(sharp, blurred, i, n, dtype) for i in range(n)]


def make_noise_stack(n, h, w, dtype=np.uint8, seed=0):
    """iid uniform noise frames: worst case for any data-dependent shortcut (every frame competes)."""
    rng = np.random.default_rng(2000 + seed)
    maxv = 255 if np.dtype(dtype) == np.uint8 else 65535
    return [rng.integers(0, maxv + 1, size=(h, w, 3), dtype=np.uint16).astype(dtype)
            for _ in range(n)]
