"""Device frame for the per-frame sub-actions: ctypes shim over the ``sssub_*`` C ABI
(include/shinestack_b200_subactions.h).  One frame is uploaded, BalanceFrames' table look-up and / or Vignetting's
radial divide run on it in HBM, and it is either downloaded (sub-action protocol of the reference's CombinedActions,
stack_framework.py:269-297) or handed on by device pointer.  No CPU fallback."""
import ctypes

import numpy as np

from . import _native
from .pyramid import _check
from .exceptions import ImageLoadError

_FRAMES = {}


class DeviceFrame:
    def __init__(self, shape, dtype, device=0):
        self._lib = _native.load()
        self.shape = (int(shape[0]), int(shape[1]), 3)
        self.dtype = np.dtype(dtype)
        if self.dtype not in (np.uint8, np.uint16):
            raise ImageLoadError("<array>", "expected an 8 or 16 bit image")
        h = ctypes.c_void_p()
        _check(self._lib.sssub_create(self.shape[1], self.shape[0], self.dtype.itemsize, int(device), ctypes.byref(h)))
        self._h = h
        self._src = None

    def close(self):
        if getattr(self, "_h", None) is not None:
            self._lib.sssub_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:  # pylint: disable=broad-except
            pass

    def upload(self, img):
        if img.ndim != 3 or img.shape != self.shape or img.dtype != self.dtype:
            raise ImageLoadError("<array>", f"expected a {self.shape} {self.dtype} frame")
        self._src = np.ascontiguousarray(img)            # keeps the host memory alive while the copy runs
        _check(self._lib.sssub_upload(self._h, self._src.ctypes.data_as(ctypes.c_void_p), self._src.strides[0]))
        return self

    def apply_lut(self, luts):
        """``luts``: (1 | 3, num_values) array of the frame's dtype (reference balance.py:27-48)."""
        luts = np.ascontiguousarray(luts, dtype=self.dtype)
        nv = 256 if self.dtype == np.uint8 else 65536
        if luts.ndim != 2 or luts.shape[0] not in (1, 3) or luts.shape[1] != nv:
            raise ValueError(f"look-up tables must have shape (1 | 3, {nv})")
        _check(self._lib.sssub_apply_lut(self._h, luts.ctypes.data_as(ctypes.c_void_p), luts.shape[0]))
        return self

    def apply_vignette(self, params, v0, max_correction=1.0, black_threshold=1.0):
        """reference vignetting.py:71-97 with the fitted ``params`` = (i0, k, r0)."""
        p = (ctypes.c_double * 3)(*[float(v) for v in params])
        thr = float(black_threshold) if self.dtype == np.uint8 else float(black_threshold) * 256
        _check(self._lib.sssub_apply_vignette(self._h, p, float(v0), float(max_correction), thr))
        return self

    def download(self, out=None):
        if out is None:
            out = np.empty(self.shape, self.dtype)
        _check(self._lib.sssub_download(self._h, out.ctypes.data_as(ctypes.c_void_p), out.strides[0]))
        self._src = None
        return out

    def device_frame(self):
        """(device pointer, row pitch) for ss_push_frame_device & co; waits for the queued passes first."""
        _check(self._lib.sssub_sync(self._h))
        d, pitch = ctypes.c_void_p(), ctypes.c_size_t()
        _check(self._lib.sssub_device_frame(self._h, ctypes.byref(d), ctypes.byref(pitch)))
        return d, pitch.value

    @property
    def kernel_ms(self):
        ms = ctypes.c_float()
        _check(self._lib.sssub_kernel_time_ms(self._h, ctypes.byref(ms)))
        return ms.value

    @property
    def launches(self):
        n = ctypes.c_int64()
        _check(self._lib.sssub_launch_count(self._h, ctypes.byref(n)))
        return n.value


def frame_for(img, device=0):
    """A cached DeviceFrame of the image's geometry (sub-actions see frame after frame of one size)."""
    key = (img.shape[0], img.shape[1], np.dtype(img.dtype).str, int(device))
    f = _FRAMES.get(key)
    if f is None or f._h is None:  # pylint: disable=protected-access
        if len(_FRAMES) >= 4:
            _, old = _FRAMES.popitem()
            old.close()
        f = _FRAMES[key] = DeviceFrame(img.shape[:2], img.dtype, device)
    return f
