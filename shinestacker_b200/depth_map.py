"""Host-side mirror of the reference ``DepthMapStack`` (reference
src/shinestacker/algorithms/depth_map.py:10-123) running on a B200.

Same constructor keywords and defaults (reference config/constants.py:144-157), same stack-algo protocol
(``process``, ``name()``, ``steps_per_frame()``, ``focus_stack(filenames)``; reference stack.py:23,29,47,108),
same progress callbacks (``after_step`` i and i + n, ``check_running`` after each, depth_map.py:75-77,113-115)
and the same exception types.  All pixel arithmetic -- gray, Sobel / blurred-Laplacian focus energy, global
normalisation, bilateral smoothing, focus map, weighted Laplacian-pyramid blend, collapse -- runs in
libshinestack_b200.so through the C ABI of include/shinestack_b200_depthmap.h.  Each file is decoded once
(the reference decodes it twice); the stack stays resident in HBM between the two passes.

There is no CPU implementation here: without the library or an sm_100 GPU every call raises
``NativeLibraryError``.
"""
import ctypes

import numpy as np

from . import _native
from .exceptions import InvalidOptionError, ImageLoadError, RunStopException
from . import fastio
from .pyramid import read_img, _check, stage_frame, default_decode_threads, PyramidStack, FLOAT_32, FLOAT_64

# reference src/shinestacker/config/constants.py:144-157
DM_ENERGY_LAPLACIAN = "laplacian"
DM_ENERGY_SOBEL = "sobel"
DM_MAP_AVERAGE = "average"
DM_MAP_MAX = "max"
DEFAULT_DM_FLOAT = FLOAT_32
DEFAULT_DM_MAP = DM_MAP_AVERAGE
DEFAULT_DM_ENERGY = DM_ENERGY_LAPLACIAN
DEFAULT_DM_KERNEL_SIZE = 5
DEFAULT_DM_BLUR_SIZE = 5
DEFAULT_DM_SMOOTH_SIZE = 15
DEFAULT_DM_TEMPERATURE = 0.1
DEFAULT_DM_LEVELS = 3


class DepthMapStack:
    """Drop-in replacement of the reference ``DepthMapStack`` (depth_map.py:10-123).

    Extra keywords: ``device`` CUDA ordinal, ``decode_threads`` files decoded ahead of the upload.
    ``float_type='float-64'`` is accepted by the constructor like the reference but rejected by the device
    path (``InvalidOptionError`` when the stack starts): only float-32 is implemented.
    """

    def __init__(self, map_type=DEFAULT_DM_MAP, energy=DEFAULT_DM_ENERGY,
                 kernel_size=DEFAULT_DM_KERNEL_SIZE, blur_size=DEFAULT_DM_BLUR_SIZE,
                 smooth_size=DEFAULT_DM_SMOOTH_SIZE, temperature=DEFAULT_DM_TEMPERATURE,
                 levels=DEFAULT_DM_LEVELS, float_type=DEFAULT_DM_FLOAT, device=0, decode_threads=None):
        # reference base_stack_algo.py:10-22
        self._name = "depth map"
        self._steps_per_frame = 2
        self.process = None
        self.do_step_callback = False
        if float_type == FLOAT_32:
            self.float_type = np.float32
        elif float_type == FLOAT_64:
            self.float_type = np.float64
        else:
            raise InvalidOptionError("float_type", float_type,
                                     details=" valid values are FLOAT_32 and FLOAT_64")
        self.map_type = map_type
        self.energy = energy
        self.kernel_size = kernel_size
        self.blur_size = blur_size
        self.smooth_size = smooth_size
        self.temperature = temperature
        self.levels = levels
        self.device = int(device)
        self.decode_threads = default_decode_threads() if decode_threads is None else int(decode_threads)
        self.last_kernel_ms = None
        self.last_launches = None
        self._lib = _native.load()
        self._h = None
        self._geom = None
        self._max_frames = 0
        self._stage = []
        self._stage_bytes = 0

    def name(self):
        return self._name

    def steps_per_frame(self):
        return self._steps_per_frame

    print_message = PyramidStack.print_message
    _staging = PyramidStack._staging
    _validate = staticmethod(PyramidStack._validate)
    _stream = PyramidStack._stream       # worker threads stage frames into the pinned ring, this thread uploads

    def _push_slot(self, _i, slot):
        _check(self._lib.ssdm_push_frame(self._h, slot[0], self._pitch))

    def _sync_copies(self):
        _check(self._lib.ssdm_sync_copies(self._h))

    def _callbacks_after_push(self, i):
        if self._report_steps:
            self._callbacks(i)

    # ---- native handle ------------------------------------------------------------------------------
    def _check_options(self):
        """The reference raises these inside focus_stack (depth_map.py:83-87, :61-62)."""
        if self.energy not in (DM_ENERGY_SOBEL, DM_ENERGY_LAPLACIAN):
            raise InvalidOptionError('energy', self.energy, details=f" valid values are "
                                     f"{DM_ENERGY_SOBEL} and {DM_ENERGY_LAPLACIAN}.")
        if self.map_type not in (DM_MAP_AVERAGE, DM_MAP_MAX):
            raise InvalidOptionError("map_type", self.map_type, details=f" valid values are "
                                     f"{DM_MAP_AVERAGE} and {DM_MAP_MAX}.")

    def _ensure_handle(self, shape, dtype, n_frames):
        self._check_options()
        h, w = int(shape[0]), int(shape[1])
        bps = np.dtype(dtype).itemsize
        geom = (h, w, bps, self.map_type, self.energy, int(self.kernel_size), int(self.blur_size),
                int(self.smooth_size), float(self.temperature), int(self.levels))
        if self._h is not None and self._geom == geom and n_frames <= self._max_frames:
            _check(self._lib.ssdm_reset(self._h))
            return
        self.close()
        cfg = _native.SsdmConfig()
        cfg.width, cfg.height, cfg.bytes_per_sample = w, h, bps
        cfg.map_type = 1 if self.map_type == DM_MAP_MAX else 0
        cfg.energy = 1 if self.energy == DM_ENERGY_SOBEL else 0
        cfg.kernel_size, cfg.blur_size = int(self.kernel_size), int(self.blur_size)
        cfg.smooth_size = int(self.smooth_size)
        cfg.temperature = float(self.temperature)
        cfg.levels = int(self.levels)
        cfg.float_bits = 64 if self.float_type == np.float64 else 32
        cfg.device = self.device
        cfg.max_frames = max(int(n_frames), 1)
        hp = ctypes.c_void_p()
        _check(self._lib.ssdm_create(ctypes.byref(cfg), ctypes.byref(hp)))
        self._h = hp
        self._geom = geom
        self._max_frames = cfg.max_frames

    def close(self):
        if getattr(self, "_h", None) is not None:
            self._lib.ssdm_destroy(self._h)
            self._h = None
        for ptr, _ in getattr(self, "_stage", []):
            self._lib.ss_host_free(ptr)
        self._stage = []
        self._stage_bytes = 0

    def __del__(self):
        try:
            self.close()
        except Exception:  # pylint: disable=broad-except
            pass

    # ---- reference protocol ---------------------------------------------------------------------------
    def _callbacks(self, step):
        # depth_map.py:75-77 and :113-115: after_step is unconditional in this class
        if self.process is None:
            return
        self.process.callback('after_step', self.process.id, self.process.name, step)
        if self.process.callback('check_running', self.process.id, self.process.name) is False:
            raise RunStopException(self.name())

    def _push(self, i, img, ring):
        """One frame through the pinned ring from the calling thread (tests; the stack entry points use _stream)."""
        slot = i % len(ring)
        if i >= len(ring) and slot == 0:
            _check(self._lib.ssdm_sync_copies(self._h))
        ptr, view = ring[slot]
        stage_frame(img, view)
        _check(self._lib.ssdm_push_frame(self._h, ptr, img.shape[1] * 3 * img.dtype.itemsize))

    def _blend_and_finish(self, n, shape, dtype, names=None):
        lib, h = self._lib, self._h
        _check(lib.ssdm_finish_begin(h))
        for i in range(n):
            if names is not None:
                self.print_message(f": reading file (2/2) {names[i].split('/')[-1]}")
            _check(lib.ssdm_blend_frame(h, i))
            self._callbacks(i + n)
        if names is not None:
            self.print_message(': blend levels')
        out = np.empty((shape[0], shape[1], 3), dtype=dtype)
        _check(lib.ssdm_finish_end(h, out.ctypes.data_as(ctypes.c_void_p), out.strides[0]))
        ms = ctypes.c_float()
        _check(lib.ssdm_kernel_time_ms(h, ctypes.byref(ms)))
        self.last_kernel_ms = ms.value
        nl = ctypes.c_int64()
        _check(lib.ssdm_launch_count(h, ctypes.byref(nl)))
        self.last_launches = nl.value
        return out

    def get_plane(self, what, frame):
        """Test hook: 'energy' (raw, before the blend), 'normalized' or 'weights' of one frame (H x W float32)."""
        code = {"normalized": 0, "weights": 1, "energy": 2}[what]
        out = np.empty(self._geom[:2], dtype=np.float32)
        _check(self._lib.ssdm_get_plane(self._h, code, int(frame), out.ctypes.data_as(ctypes.c_void_p)))
        return out

    def focus_stack(self, filenames):
        """reference depth_map.py:64-123: list of image paths -> fused H x W x 3 BGR array of the input dtype.
        Files are decoded and copied into pinned memory by ``decode_threads`` workers while this thread uploads."""
        filenames = list(filenames)
        n = len(filenames)
        if n == 0:
            raise ImageLoadError("", "empty file list")
        try:
            self.print_message(f": reading file (1/2) {filenames[0].split('/')[-1]}")
            first = fastio.read_frame(filenames[0], read_img)
            if first is None:
                raise ImageLoadError(filenames[0])
            if first.ndim != 3 or first.shape[2] != 3 or first.dtype not in (np.uint8, np.uint16):
                raise ImageLoadError(filenames[0], "expected a 3-channel 8 or 16 bit image")
            shape, dtype = tuple(first.shape), np.dtype(first.dtype)
            self._ensure_handle(shape, dtype, n)
            self._pitch = shape[1] * 3 * dtype.itemsize
            self._report_steps = True
            ring = self._staging(shape[0] * shape[1] * 3 * dtype.itemsize, 2 * max(self.decode_threads, 2))

            def load_into(i, view):
                img = first if i == 0 else fastio.read_frame(filenames[i], read_img)
                if img is not None and tuple(img.shape) == shape and img.dtype == dtype:
                    stage_frame(img, view)
                return img

            def on_frame(i, img):
                if i:
                    self.print_message(f": reading file (1/2) {filenames[i].split('/')[-1]}")
                if img is None:
                    raise ImageLoadError(filenames[i])
                self._validate(img, shape[:2], dtype)

            self._stream(n, load_into, ring, on_frame)
            return self._blend_and_finish(n, shape[:2], dtype, filenames)
        except BaseException:
            if self._h is not None:
                self._lib.ssdm_reset(self._h)
            raise

    def focus_stack_arrays(self, frames, stage=True):
        """In-memory entry point: iterable of H x W x 3 BGR uint8/uint16 arrays -> fused image.
        ``stage=False`` uploads straight from the arrays' own memory (C-contiguous, ideally pinned, e.g.
        allocated with ss_host_alloc) instead of copying each frame into the pinned staging ring first."""
        frames = list(frames)
        if not frames:
            raise ImageLoadError("", "empty frame list")
        img0 = frames[0]
        if img0.ndim != 3 or img0.shape[2] != 3 or img0.dtype not in (np.uint8, np.uint16):
            raise ImageLoadError("<array>", "expected a 3-channel 8 or 16 bit image")
        self._ensure_handle(img0.shape, img0.dtype, len(frames))
        self._pitch = img0.shape[1] * 3 * img0.dtype.itemsize
        self._report_steps = False
        shape2, dtype = img0.shape[:2], img0.dtype

        def load_into(i, view):
            img = frames[i]
            if img is not None and img.shape == img0.shape and img.dtype == dtype:
                stage_frame(img, view)         # pageable -> pinned on a worker thread (numpy releases the GIL)
            return img

        try:
            if stage:
                ring = self._staging(img0.nbytes, 2 * max(self.decode_threads, 2))
                self._stream(len(frames), load_into, ring, lambda _i, img: self._validate(img, shape2, dtype))
            else:
                for img in frames:
                    self._validate(img, shape2, dtype)
                    if not img.flags.c_contiguous:
                        raise ValueError("stage=False needs C-contiguous frames")
                    _check(self._lib.ssdm_push_frame(self._h, img.ctypes.data_as(ctypes.c_void_p), img.strides[0]))
            return self._blend_and_finish(len(frames), img0.shape[:2], img0.dtype)
        except BaseException:
            if self._h is not None:
                self._lib.ssdm_reset(self._h)
            raise
