"""Shared helpers of the AlignFrames tests: the reference's warp stage spelled with cv2 (align.py:228-251),
transform generators and the ctypes wrapper of the CPU emulation of the kernels (tests/emu/emu_align.cpp)."""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
MODES = {"BORDER_CONSTANT": 0, "BORDER_REPLICATE": 1, "BORDER_REPLICATE_BLUR": 2}


def cv2_warp_stage(img, m, border_mode, border_value=(0, 0, 0, 0), border_blur=50):
    """Literally the reference's calls (align.py:228-251) with the transform given."""
    import cv2
    h, w = img.shape[:2]
    cvmode = cv2.BORDER_CONSTANT if border_mode == "BORDER_CONSTANT" else cv2.BORDER_REPLICATE
    f = cv2.warpPerspective if np.asarray(m).size == 9 else cv2.warpAffine
    out = f(img, m, (w, h), borderMode=cvmode, borderValue=list(border_value))
    if border_mode == "BORDER_REPLICATE_BLUR":
        mask = f(np.ones_like(img, dtype=np.uint8), m, (w, h), borderMode=cv2.BORDER_CONSTANT, borderValue=0)
        mask = cv2.cvtColor(mask, cv2.COLOR_BGR2GRAY)
        blurred = cv2.GaussianBlur(out, (21, 21), sigmaX=border_blur)
        out[mask == 0] = blurred[mask == 0]
    return out


def rigid(angle_deg, scale, tx, ty, dtype=np.float64):
    a = np.deg2rad(angle_deg)
    return np.array([[scale * np.cos(a), -scale * np.sin(a), tx], [scale * np.sin(a), scale * np.cos(a), ty]], dtype)


def homography(angle_deg, scale, tx, ty, p0=1.2e-5, p1=-2.3e-5):
    return np.vstack([rigid(angle_deg, scale, tx, ty), [p0, p1, 1.0]])


TRANSFORMS = [
    ("rigid_small", rigid(0.4, 1.003, 7.3, -4.6)),
    ("rigid_large", rigid(-3.0, 0.98, -20.2, 13.9)),
    ("rigid_f32", rigid(0.1, 1.0005, 0.5, 0.25, np.float32)),      # what the sub-sampled estimator returns
    ("rigid_identity", rigid(0, 1, 0, 0)),
    ("homography", homography(0.7, 1.01, 5.5, -8.25)),
]


def transform_error(m, m_true, shape):
    """(angle difference in degrees, scale difference, displacement difference at the image centre in pixels):
    the three quantities the reference's precision test bounds (tests/test_0031_align_precision.py:50-65)."""
    ang = np.degrees(np.arctan2(m[1, 0], m[0, 0]) - np.arctan2(m_true[1, 0], m_true[0, 0]))
    scale = np.hypot(m[0, 0], m[1, 0]) - np.hypot(m_true[0, 0], m_true[1, 0])
    c = np.array([shape[1] / 2.0, shape[0] / 2.0, 1.0])
    return abs(ang), abs(scale), float(np.abs(m @ c - m_true @ c).max())


def textured_scene(h, w, seed=9, blur=1.0):
    """Shapes + noise, slightly blurred (a photographic frame is band-limited too)."""
    import cv2
    rng = np.random.default_rng(seed)
    img = np.zeros((h, w, 3), np.uint8)
    for _ in range(max(40, h * w // 6000)):
        c = tuple(int(v) for v in rng.integers(60, 255, 3))
        p = (int(rng.integers(0, w)), int(rng.integers(0, h)))
        if rng.random() < 0.5:
            cv2.circle(img, p, int(rng.integers(5, 40)), c, int(rng.integers(1, 4)))
        else:
            q = (int(rng.integers(0, w)), int(rng.integers(0, h)))
            cv2.rectangle(img, p, q, c, int(rng.integers(1, 3)))
    img = cv2.add(img, rng.integers(0, 12, img.shape).astype(np.uint8))
    return cv2.GaussianBlur(img, (0, 0), blur) if blur > 0 else img


def moved_frame(img, angle_deg, scale, tx, ty):
    """Returns (moving frame, forward transform that aligns it back onto img)."""
    import cv2
    h, w = img.shape[:2]
    t = cv2.getRotationMatrix2D((w / 2.0, h / 2.0), angle_deg, scale)
    t[:, 2] += (tx, ty)
    return cv2.warpAffine(img, t, (w, h), borderMode=cv2.BORDER_REPLICATE), cv2.invertAffineTransform(t)


class EmuWarper:
    _lib = None

    @classmethod
    def lib(cls):
        if cls._lib is None:
            src = os.path.join(HERE, "emu", "emu_align.cpp")
            so = os.path.join(HERE, "emu", "libemu_align.so")
            csrc = os.path.join(HERE, "..", "shinestacker_b200", "csrc")
            deps = [src, os.path.join(HERE, "emu", "emu_backend.h")] + \
                [os.path.join(csrc, f) for f in ("ss_al_kernels.cuh", "ss_al_engine.h", "ss_dm_kernels.cuh", "ss_dm_engine.h")]
            if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(d) for d in deps):
                subprocess.run(["g++", "-std=c++17", "-O2", "-mfma", "-ffp-contract=off", "-shared", "-fPIC",
                                "-o", so, src], check=True)
            lib = ctypes.CDLL(so)
            vp, i, sz = ctypes.c_void_p, ctypes.c_int, ctypes.c_size_t
            lib.emu_al_create.restype = vp
            lib.emu_al_create.argtypes = [i, i, i]
            lib.emu_al_destroy.argtypes = [vp]
            lib.emu_al_destroy.restype = None
            lib.emu_al_error.restype = ctypes.c_char_p
            lib.emu_al_error.argtypes = [vp]
            lib.emu_al_warp.argtypes = [vp, vp, sz, vp, i, i, vp, vp, i, vp, sz]
            lib.emu_al_ecc_set_reference.argtypes = [vp, vp, sz]
            lib.emu_al_ecc_estimate.argtypes = [vp, vp, sz, i, i, ctypes.c_double, vp, vp, vp]
            cls._lib = lib
        return cls._lib

    def __init__(self, h, w, dtype):
        self.lib_ = self.lib()
        self.p = self.lib_.emu_al_create(w, h, np.dtype(dtype).itemsize)
        assert self.p

    def warp(self, img, m, border_mode, border_value=(0, 0, 0, 0), kernel=None):
        img = np.ascontiguousarray(img)
        m = np.ascontiguousarray(np.asarray(m, np.float64))
        bval = np.asarray(list(border_value) + [0] * 4, np.float64)[:4].copy()
        out = np.empty_like(img)
        k = np.ascontiguousarray(kernel, np.int32) if kernel is not None else None
        rc = self.lib_.emu_al_warp(self.p, img.ctypes.data, img.strides[0], m.ctypes.data, int(m.size == 9),
                                   MODES[border_mode], bval.ctypes.data, k.ctypes.data if k is not None else None,
                                   len(k) if k is not None else 0, out.ctypes.data, out.strides[0])
        if rc:
            raise ValueError(self.lib_.emu_al_error(self.p).decode())
        return out

    def ecc_set_reference(self, img):
        img = np.ascontiguousarray(img)
        if self.lib_.emu_al_ecc_set_reference(self.p, img.ctypes.data, img.strides[0]):
            raise ValueError(self.lib_.emu_al_error(self.p).decode())

    def ecc_estimate(self, img, finest_level=0, max_iters=30, eps_px=1e-3, init_m=None):
        img = np.ascontiguousarray(img)
        m, info = np.zeros(6), np.zeros(4)
        init = None if init_m is None else np.ascontiguousarray(np.asarray(init_m, np.float64).reshape(6))
        rc = self.lib_.emu_al_ecc_estimate(self.p, img.ctypes.data, img.strides[0], finest_level, max_iters, eps_px,
                                           init.ctypes.data if init is not None else None, m.ctypes.data, info.ctypes.data)
        if rc:
            raise ValueError(self.lib_.emu_al_error(self.p).decode())
        return m.reshape(2, 3), dict(rho=info[0], iterations=int(info[1]), overlap=info[2], last_step=info[3])

    def close(self):
        if self.p:
            self.lib_.emu_al_destroy(self.p)
            self.p = None

    def __del__(self):
        self.close()
