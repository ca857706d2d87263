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

    def close(self):
        if self.p:
            self.lib_.emu_al_destroy(self.p)
            self.p = None

    def __del__(self):
        self.close()
