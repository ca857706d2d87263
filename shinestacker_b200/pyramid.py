"""Host-side mirror of the reference stack-algo interface for the pyramid path.

``PyramidStack`` here has the constructor, attributes and methods that
``FocusStack`` / ``FocusStackBunch`` (reference src/shinestacker/algorithms/stack.py:23,29,47,76,108)
use on a ``stack_algo`` object -- ``process``, ``do_step_callback``, ``name()``, ``steps_per_frame()``,
``focus_stack(filenames)`` -- and the same constructor keywords, progress callbacks and exception
types as the reference class (reference src/shinestacker/algorithms/pyramid.py:10-22,114-179 and
base_stack_algo.py:9-42).  All pixel arithmetic runs in libshinestack_b200.so (hand-written sm_100a
CUDA kernels) through the C ABI declared in include/shinestack_b200.h; this module only decodes
files (cv2.imread, as the reference does), moves bytes and computes the base-level log-probability
table with numpy (so that it is bit-identical to the reference's ``np.log``).

There is deliberately no CPU implementation here: without the shared library or without an sm_100
GPU every call raises ``NativeLibraryError``.
"""
import ctypes
import os
import threading
from concurrent.futures import ThreadPoolExecutor

import numpy as np

from . import _native, fastio
from .exceptions import (InvalidOptionError, ImageLoadError, BitDepthError, ShapeError,
                         RunStopException, NativeLibraryError)

# reference src/shinestacker/config/constants.py:159-162, :8-11
DEFAULT_PY_MIN_SIZE = 32
DEFAULT_PY_KERNEL_SIZE = 5
DEFAULT_PY_GEN_KERNEL = 0.4
FLOAT_32 = 'float-32'
FLOAT_64 = 'float-64'
DEFAULT_PY_FLOAT = FLOAT_32


def read_img(file_path):
    """reference src/shinestacker/algorithms/utils.py:11-20."""
    import cv2  # pylint: disable=import-outside-toplevel
    if not os.path.isfile(file_path):
        raise RuntimeError("File does not exist: " + file_path)
    ext = file_path.split(".")[-1]
    img = None
    if ext in ['jpeg', 'jpg']:
        img = cv2.imread(file_path)
    elif ext in ['tiff', 'tif', 'png']:
        img = cv2.imread(file_path, cv2.IMREAD_UNCHANGED)
    return img


def log_prob_table(counts, dtype=np.float32):
    """np.log(probabilities) of PyramidBase.entropy (reference pyramid.py:67-69,77) for every frame.

    ``counts``: (n_frames, num_values) int64 histogram of each frame's truncated gray base level.
    The reference evaluates ``counts.astype(float32) / counts.sum()`` in float64 (float32 array /
    int64 scalar) and rounds to float32 when storing into ``probabilities``; values that do not
    occur keep probability 0 (-inf after the log, never indexed).
    """
    counts = np.asarray(counts, dtype=np.int64)
    total = counts.sum(axis=1, keepdims=True)
    prob = (counts.astype(dtype) / total).astype(dtype)
    with np.errstate(divide="ignore"):
        return np.ascontiguousarray(np.log(prob))


def log_prob_per_pixel(gray, dtype=np.float32):
    """16-bit variant of log_prob_table: for every base pixel the value np.log(probabilities)[its gray value].

    ``gray``: (n_frames, hL, wL) int32 truncated gray base levels.  Per frame this is literally the
    reference's expression (pyramid.py:67-69): np.unique(..., return_counts=True), counts.astype(float32) /
    counts.sum() (evaluated in float64, stored as float32), np.log.
    """
    gray = np.asarray(gray)
    out = np.empty(gray.shape, dtype=dtype)
    for f in range(gray.shape[0]):
        _, inverse, counts = np.unique(gray[f].ravel(), return_inverse=True, return_counts=True)
        prob = np.zeros(counts.shape, dtype=dtype)
        prob[...] = counts.astype(dtype) / counts.sum()
        out[f] = np.log(prob)[inverse].reshape(gray.shape[1:])
    return out


def log_prob_lut(n_base, dtype=np.float32):
    """np.log(probabilities) of PyramidBase.entropy (reference pyramid.py:67-69,77) for every value a probability
    can take: ``probabilities[v] = counts.astype(float32)[v] / counts.sum()`` with ``counts.sum() == n_base`` (every
    base pixel is counted once), evaluated like the reference (float32 array / int64 scalar -> float64, stored as
    float32) and passed through numpy's own ``log``.  Entry ``c`` belongs to a value that occurs ``c`` times."""
    c = np.arange(n_base + 1, dtype=np.int64)
    prob = (c.astype(dtype) / np.int64(n_base)).astype(dtype)
    with np.errstate(divide="ignore"):
        return np.ascontiguousarray(np.log(prob))


def select_base(lib, h, n_frames, num_values, float_type=np.float32):
    """Base-level arg-max on the device.  float-32: one asynchronous call (the device indexes the log table set by
    ``_ensure_handle`` with its own histogram); float-64: histogram to the host, numpy log table back."""
    f64 = float_type == np.float64
    if not f64:
        _check(lib.ss_finish_base_device(h))
        return
    if num_values <= 256:
        counts = np.empty((n_frames, num_values), dtype=np.int64)
        _check(lib.ss_finish_begin(h, counts.ctypes.data_as(ctypes.c_void_p)))
        table = log_prob_table(counts, float_type)
        if f64:
            _check(lib.ss_finish_select_base_f64(h, table.ctypes.data_as(ctypes.c_void_p), 0))
        else:
            _check(lib.ss_finish_select_base(h, table.ctypes.data_as(ctypes.c_void_p)))
    else:
        _check(lib.ss_finish_begin(h, None))
        L = lib.ss_num_levels(h)
        hh, ww = ctypes.c_int32(), ctypes.c_int32()
        _check(lib.ss_level_size(h, L, ctypes.byref(hh), ctypes.byref(ww)))
        gray = np.empty((n_frames, hh.value, ww.value), dtype=np.int32)
        _check(lib.ss_get_base_gray(h, gray.ctypes.data_as(ctypes.c_void_p)))
        table = log_prob_per_pixel(gray, float_type)
        if f64:
            _check(lib.ss_finish_select_base_f64(h, table.ctypes.data_as(ctypes.c_void_p), 1))
        else:
            _check(lib.ss_finish_select_base_px(h, table.ctypes.data_as(ctypes.c_void_p)))


def default_decode_threads():
    return max(1, min(12, (os.cpu_count() or 4) - 2))


def stage_frame(img, view):
    """Copy one decoded frame -- an ndarray from cv2.imread or a mapped plain TIFF (fastio) -- into a pinned
    staging buffer (flat uint8 view)."""
    if isinstance(img, fastio.PlainTiff):
        img.into_bgr(view[:img.nbytes].view(img.dtype).reshape(img.shape))
        img.close()
    else:
        img = np.ascontiguousarray(img)
        view[:img.nbytes] = img.reshape(-1).view(np.uint8)


def _check(rc):
    if rc == _native.SS_OK:
        return
    msg = _native.last_error()
    if rc == _native.SS_ERR_INVALID_OPTION:
        raise InvalidOptionError("native", rc, details=msg)
    raise NativeLibraryError(f"libshinestack_b200 error {rc}: {msg}")


class PyramidBase:
    """Constructor contract of the reference PyramidBase/BaseStackAlgo (pyramid.py:10-22,
    base_stack_algo.py:9-31)."""

    def __init__(self, min_size=DEFAULT_PY_MIN_SIZE, kernel_size=DEFAULT_PY_KERNEL_SIZE,
                 gen_kernel=DEFAULT_PY_GEN_KERNEL, float_type=DEFAULT_PY_FLOAT):
        self._name = "pyramid"
        self._steps_per_frame = 2
        self.process = None
        if float_type == FLOAT_32:
            self.float_type = np.float32
        elif float_type == FLOAT_64:
            self.float_type = np.float64
        else:
            raise InvalidOptionError(
                "float_type", float_type,
                details=" valid values are FLOAT_32 and FLOAT_64")
        self.min_size = min_size
        self.kernel_size = kernel_size
        self.pad_amount = (kernel_size - 1) // 2
        self.do_step_callback = False
        self.gen_kernel_param = float(gen_kernel)
        kernel = np.array([0.25 - gen_kernel / 2.0, 0.25,
                           gen_kernel, 0.25, 0.25 - gen_kernel / 2.0])
        self.gen_kernel = np.outer(kernel, kernel)

    def name(self):
        return self._name

    def steps_per_frame(self):
        return self._steps_per_frame

    def print_message(self, msg):
        if self.process is None:
            return
        try:
            from shinestacker.core.colors import color_str  # type: ignore
            from shinestacker.config.constants import constants  # type: ignore
            msg = color_str(msg, constants.LOG_COLOR_LEVEL_3)
        except Exception:  # pylint: disable=broad-except
            pass
        self.process.sub_message_r(msg)


class PyramidStack(PyramidBase):
    """Drop-in replacement of the reference ``PyramidStack`` (pyramid.py:114-179) running on a B200.

    Extra keywords (all optional): ``device`` CUDA ordinal, ``batch_frames`` frames fused per kernel
    chain launch, ``keep_depth_maps`` download the integer arg-max maps after each stack,
    ``decode_threads`` files decoded ahead of the upload.
    """

    def __init__(self, min_size=DEFAULT_PY_MIN_SIZE, kernel_size=DEFAULT_PY_KERNEL_SIZE,
                 gen_kernel=DEFAULT_PY_GEN_KERNEL, float_type=DEFAULT_PY_FLOAT,
                 device=0, batch_frames=8, keep_depth_maps=True, frame_index_base=0, decode_threads=None):
        super().__init__(min_size, kernel_size, gen_kernel, float_type)
        self.offset = np.arange(-self.pad_amount, self.pad_amount + 1)
        self.dtype = None
        self.num_pixel_values = None
        self.max_pixel_value = None
        self.device = int(device)
        self.batch_frames = int(batch_frames)
        self.keep_depth_maps = keep_depth_maps
        self.frame_index_base = int(frame_index_base)
        # workers that decode / copy frames into the pinned ring ahead of the upload (0 = inline); default: the host's
        # cores minus two, at most 12 (B200 box, 16 cores, 64 x 37.7 MB TIFFs: 6 -> 100 ms, 8 -> 87 ms, 12 -> 85 ms)
        self.decode_threads = default_decode_threads() if decode_threads is None else int(decode_threads)
        self._depth_maps = None    # dict(best=[level 0..L-1], best_e, best_d) after focus_stack, fetched on first use
        self._depth_maps_ready = False
        self.last_kernel_ms = None
        self.last_launches = None
        self._lib = _native.load()
        self._h = None
        self._geom = None
        self._stage = []
        self._stage_bytes = 0
        self._frame_shape = None
        self._rgb_frames = set()
        self.device_channel_swap = True     # mapped RGB TIFFs: upload the raw payload, swap R <-> B on the device

    # ---- native handle management ---------------------------------------------------------------
    def _ensure_handle(self, shape, dtype, n_frames, resident=False):
        h, w = int(shape[0]), int(shape[1])
        bps = np.dtype(dtype).itemsize
        # everything ss_create bakes into the native handle: changing a public option between two calls must not
        # silently keep the old configuration
        geom = (h, w, bps, int(self.min_size), int(self.kernel_size), float(self.gen_kernel_param),
                64 if self.float_type == np.float64 else 32, self.device, self.batch_frames, self.frame_index_base)
        if self._h is not None and self._geom == geom and n_frames <= self._max_frames and \
                (not resident or self._resident >= n_frames):
            self._depth_maps_ready = False       # maps not fetched by now belong to the stack being replaced
            _check(self._lib.ss_reset(self._h))
            return
        self.close()
        cfg = _native.SsConfig()
        cfg.width, cfg.height, cfg.bytes_per_sample = w, h, bps
        cfg.min_size, cfg.kernel_size = int(self.min_size), int(self.kernel_size)
        cfg.gen_kernel = self.gen_kernel_param
        cfg.float_bits = 64 if self.float_type == np.float64 else 32
        cfg.device = self.device
        cfg.max_frames = max(int(n_frames), 1)
        cfg.batch_frames = self.batch_frames
        cfg.frame_index_base = self.frame_index_base
        cfg.resident_frames = int(n_frames) if resident else 0
        hp = ctypes.c_void_p()
        _check(self._lib.ss_create(ctypes.byref(cfg), ctypes.byref(hp)))
        self._h = hp
        if cfg.float_bits == 32:
            L = self._lib.ss_num_levels(hp)
            hh, ww = ctypes.c_int32(), ctypes.c_int32()
            _check(self._lib.ss_level_size(hp, L, ctypes.byref(hh), ctypes.byref(ww)))
            lut = log_prob_lut(hh.value * ww.value)
            _check(self._lib.ss_set_log_lut(hp, lut.ctypes.data_as(ctypes.c_void_p), lut.size))
        self._geom = geom
        self._max_frames = cfg.max_frames
        self._resident = cfg.resident_frames

    def close(self):
        if getattr(self, "_h", None) is not None:
            if getattr(self, "_depth_maps_ready", False) and self._depth_maps is None:
                try:
                    self._depth_maps = self.get_depth_maps()      # keep the last stack's maps readable after close
                except Exception:  # pylint: disable=broad-except
                    pass
            self._depth_maps_ready = False
            self._lib.ss_destroy(self._h)
            self._h = None
        for ptr, _ in getattr(self, "_stage", []):
            self._lib.ss_host_free(ptr)
        self._stage = []

    def __del__(self):
        try:
            self.close()
        except Exception:  # pylint: disable=broad-except
            pass

    def _staging(self, nbytes, count):
        """Ring of pinned host buffers the decoded frames are copied into before the async upload."""
        if self._stage_bytes != nbytes or len(self._stage) < count:
            for ptr, _ in self._stage:
                self._lib.ss_host_free(ptr)
            self._stage = []
            for _ in range(count):
                p = ctypes.c_void_p()
                _check(self._lib.ss_host_alloc(ctypes.byref(p), nbytes))
                buf = (ctypes.c_uint8 * nbytes).from_address(p.value)
                self._stage.append((p, np.frombuffer(buf, dtype=np.uint8)))
            self._stage_bytes = nbytes
        return self._stage

    # ---- reference protocol ---------------------------------------------------------------------
    def _set_metadata(self, img):
        self.dtype = img.dtype
        self.num_pixel_values = 256 if self.dtype == np.uint8 else 65536
        self.max_pixel_value = 255 if self.dtype == np.uint8 else 65535

    @staticmethod
    def _validate(img, shape, dtype):
        # reference utils.py:56-63
        if img is None:
            raise RuntimeError("Image is None")
        if img.shape[:2] != shape[:2]:
            raise ShapeError(shape, img.shape[:2])
        if img.dtype != dtype:
            raise BitDepthError(dtype, img.dtype)

    def _callbacks(self, step):
        if self.process is None:
            return
        if self.do_step_callback:
            self.process.callback('after_step', self.process.id, self.process.name, step)
        if self.process.callback('check_running', self.process.id, self.process.name) is False:
            raise RunStopException(self.name())

    def _push(self, i, img, ring):
        slot = i % len(ring)
        if i >= len(ring) and slot == 0:
            _check(self._lib.ss_sync_copies(self._h))
        ptr, view = ring[slot]
        stage_frame(img, view)
        _check(self._lib.ss_push_frame(self._h, ptr, img.shape[1] * 3 * img.dtype.itemsize))

    def _ring_slots(self):
        """Pinned slots: two per frame of a kernel-chain batch and two per worker, so that neither waits for the other."""
        return 2 * max(self.batch_frames, min(self.decode_threads, 12), 1)

    def _stream(self, n, load_into, ring, on_frame):
        """Decode / convert on worker threads straight into the pinned ring, upload from the calling thread.

        ``load_into(i, view)`` (worker thread) fills the flat uint8 ``view`` of a pinned slot with frame ``i`` and
        returns what ``on_frame(i, info)`` (calling thread, stack order, before the upload) should see -- or raises.
        Slot ``i % R`` is handed to frame ``i`` only once frame ``i - R`` has reached the device.  The calling
        thread does nothing but validate, push and run the callbacks, so decode, host copies, PCIe and the kernel
        chain all overlap."""
        slots = len(ring)
        if slots < 1:
            raise NativeLibraryError("no pinned staging buffers")
        threads = max(1, min(self.decode_threads if self.decode_threads > 0 else 1, slots, n))
        cond = threading.Condition()
        state = {"released": 0, "abort": False}

        def job(i):
            with cond:
                while i - slots >= state["released"] and not state["abort"]:
                    cond.wait(0.05)
                if state["abort"]:
                    return None
            return load_into(i, ring[i % slots][1])

        # the workers are the parallelism: cv2's own thread pool serialises calls that arrive from several threads, so
        # each conversion (RGB -> BGR of a mapped TIFF) runs single-threaded while the pipeline is active
        cv2_threads = None
        if threads > 1:
            try:
                import cv2  # pylint: disable=import-outside-toplevel
                cv2_threads = cv2.getNumThreads()
                cv2.setNumThreads(1)
            except Exception:  # pylint: disable=broad-except
                cv2_threads = None
        pool = ThreadPoolExecutor(max_workers=threads)
        futures = [pool.submit(job, i) for i in range(n)]
        half = max(1, slots // 2)
        try:
            for i in range(n):
                info = futures[i].result()
                on_frame(i, info)
                self._push_slot(i, ring[i % slots])
                self._callbacks_after_push(i)
                if (i + 1) % half == 0 or i + 1 == n:
                    self._sync_copies()                            # those frames are on the device: free their slots
                    with cond:
                        state["released"] = i + 1
                        cond.notify_all()
        finally:
            with cond:
                state["abort"] = True
                cond.notify_all()
            for fut in futures:
                fut.cancel()
            pool.shutdown(wait=True)
            if cv2_threads is not None:
                import cv2  # pylint: disable=import-outside-toplevel
                cv2.setNumThreads(cv2_threads)

    def _push_slot(self, i, slot):
        """Asynchronous upload of the frame staged in the pinned ``slot`` = (pointer, flat uint8 view)."""
        push = self._lib.ss_push_frame_rgb if i in self._rgb_frames else self._lib.ss_push_frame
        _check(push(self._h, slot[0], self._frame_shape[1] * 3 * np.dtype(self.dtype).itemsize))

    def _sync_copies(self):
        _check(self._lib.ss_sync_copies(self._h))

    def _callbacks_after_push(self, i):
        """Hook of ``_stream`` (focus_stack reports the "processed" step of frame i here)."""

    def _finish(self, n, shape):
        lib, h = self._lib, self._h
        select_base(lib, h, n, self.num_pixel_values, self.float_type)
        out = np.empty((shape[0], shape[1], 3), dtype=self.dtype)
        _check(lib.ss_finish_collapse(h, out.ctypes.data_as(ctypes.c_void_p), out.strides[0]))
        ms = ctypes.c_float()
        _check(lib.ss_kernel_time_ms(h, ctypes.byref(ms)))
        self.last_kernel_ms = ms.value
        nl = ctypes.c_int64()
        _check(lib.ss_launch_count(h, ctypes.byref(nl)))
        self.last_launches = nl.value
        # the integer arg-max maps stay in HBM until the next stack; `depth_maps` downloads them on first access
        self._depth_maps, self._depth_maps_ready = None, bool(self.keep_depth_maps)
        return out

    @property
    def depth_maps(self):
        """dict(best=[level 0 .. L-1], best_e, best_d) of the last stack (``keep_depth_maps``), else None.  Downloaded
        from the device on first access, so stacks whose maps nobody looks at do not pay for the transfer."""
        if self._depth_maps is None and self._depth_maps_ready and self._h is not None:
            self._depth_maps = self.get_depth_maps()
        return self._depth_maps

    def get_depth_maps(self):
        """Integer arg-max maps of the last stack (np.argmax results of reference pyramid.py:51 and
        :104-105): dict(best=[level 0 .. L-1], best_e, best_d)."""
        lib, h = self._lib, self._h
        L = lib.ss_num_levels(h)
        maps = []
        for level in range(L + 2):
            hh, ww = ctypes.c_int32(), ctypes.c_int32()
            _check(lib.ss_level_size(h, min(level, L), ctypes.byref(hh), ctypes.byref(ww)))
            m = np.empty((hh.value, ww.value), dtype=np.int32)
            _check(lib.ss_get_argmax(h, level, m.ctypes.data_as(ctypes.c_void_p)))
            maps.append(m)
        return dict(best=maps[:L], best_e=maps[L], best_d=maps[L + 1], levels=L)

    def get_fused_pyramid(self):
        """Fused Laplacian levels + fused base of the last stack, H_l x W_l x 3 float32 (tests; float-32 only)."""
        lib, h = self._lib, self._h
        L = lib.ss_num_levels(h)
        out = []
        for level in range(L + 1):
            hh, ww = ctypes.c_int32(), ctypes.c_int32()
            _check(lib.ss_level_size(h, level, ctypes.byref(hh), ctypes.byref(ww)))
            a = np.empty((hh.value, ww.value, 3), dtype=np.float32)
            _check(lib.ss_get_fused_level(h, level, a.ctypes.data_as(ctypes.c_void_p)))
            out.append(a)
        return out

    def _probe(self, path):
        """(shape, dtype, ndim, channels) of what read_img would return for ``path``: from the header when that is
        conclusive, else by decoding the file (like the reference's validation pass, pyramid.py:157-158)."""
        if not os.path.isfile(path):
            raise RuntimeError("File does not exist: " + path)     # reference utils.py:12-13
        hdr = fastio.probe_header(path)
        if hdr is not None:
            h, w, ch, dtype = hdr
            return (h, w, ch), dtype
        img = read_img(path)
        if img is None:
            return None
        return img.shape, img.dtype

    def focus_stack(self, filenames):
        """reference pyramid.py:150-179: list of image paths -> fused H x W x 3 BGR array of the input dtype.

        Two passes in the reference's order: every file is validated first (steps 0 .. n-1: shape / bit depth of
        all frames against the first one, pyramid.py:155-169 -- here from the file headers, so a mismatch in
        frame k surfaces before any GPU work), then processed (steps n .. 2n-1, pyramid.py:170-177; each file is
        decoded once, ahead of the upload, on worker threads)."""
        n = len(filenames)
        metadata = None
        for i, img_path in enumerate(filenames):
            self.print_message(f": validating file {img_path.split('/')[-1]}")
            probed = self._probe(img_path)
            if probed is None:
                raise ImageLoadError(img_path)
            shape, dtype = probed
            if metadata is None:
                if len(shape) != 3 or shape[2] != 3 or dtype not in (np.uint8, np.uint16):
                    raise ImageLoadError(img_path, "expected a 3-channel 8 or 16 bit image")
                metadata = (tuple(shape[:2]), np.dtype(dtype))
                self.dtype = metadata[1]
                self.num_pixel_values = 256 if self.dtype == np.uint8 else 65536
                self.max_pixel_value = 255 if self.dtype == np.uint8 else 65535
            else:
                if tuple(shape[:2]) != metadata[0]:
                    raise ShapeError(metadata[0], tuple(shape[:2]))
                if np.dtype(dtype) != metadata[1]:
                    raise BitDepthError(metadata[1], np.dtype(dtype))
            self._callbacks(i)
        if metadata is None:
            raise ImageLoadError("", "empty file list")
        shape2, dtype = metadata
        frame_bytes = shape2[0] * shape2[1] * 3 * dtype.itemsize
        self._frame_shape = (shape2[0], shape2[1], 3)
        rgb_slots = self._rgb_frames = set()

        def load_into(i, view):
            # worker thread: uncompressed TIFFs (what the reference's write_img produces) are mapped and converted
            # straight into the pinned slot; everything else is decoded by cv2.imread like the reference and copied
            img = fastio.read_frame(filenames[i], read_img)
            if img is None:
                return None
            try:
                if img.ndim != 3 or img.shape[2] != 3 or tuple(img.shape[:2]) != shape2 or img.dtype != dtype:
                    return (img.shape, img.dtype, img.ndim)
                if isinstance(img, fastio.PlainTiff) and self.device_channel_swap:
                    img.read_rgb_into(view)          # raw R, G, B bytes: the device swaps the channels
                    rgb_slots.add(i)
                else:
                    stage_frame(img, view)
            finally:
                if isinstance(img, fastio.PlainTiff):
                    img.close()
            return (img.shape, img.dtype, img.ndim)

        def on_frame(i, info):
            self.print_message(f": processing file {filenames[i].split('/')[-1]}")
            if info is None:
                raise ImageLoadError(filenames[i])
            shape, dt, ndim = info
            if ndim != 3 or shape[2] != 3:
                raise ImageLoadError(filenames[i], "expected a 3-channel 8 or 16 bit image")
            if tuple(shape[:2]) != shape2:                     # the decoder has the last word (reference utils.py:56-63)
                raise ShapeError(shape2, tuple(shape[:2]))
            if dt != dtype:
                raise BitDepthError(dtype, dt)

        try:
            self._ensure_handle(shape2, dtype, n)
            ring = self._staging(frame_bytes, self._ring_slots())
            self._callbacks_after_push = lambda i: self._callbacks(i + n)
            self._stream(n, load_into, ring, on_frame)
            self.print_message(': pyramids fusion completed')
            return self._finish(n, shape2)
        except BaseException:
            if self._h is not None:
                self._lib.ss_reset(self._h)
            raise
        finally:
            self._callbacks_after_push = lambda i: None

    def focus_stack_arrays(self, frames):
        """In-memory entry point: iterable of H x W x 3 BGR uint8/uint16 arrays -> fused image."""
        frames = list(frames)
        if not frames:
            raise ImageLoadError("", "empty frame list")
        img0 = frames[0]
        if img0.ndim != 3 or img0.shape[2] != 3 or img0.dtype not in (np.uint8, np.uint16):
            raise ImageLoadError("<array>", "expected a 3-channel 8 or 16 bit image")
        self._set_metadata(img0)
        self._ensure_handle(img0.shape, img0.dtype, len(frames))
        ring = self._staging(img0.nbytes, self._ring_slots())
        self._frame_shape = tuple(img0.shape)
        self._rgb_frames = set()
        shape2, dtype = img0.shape[:2], img0.dtype

        def load_into(i, view):
            img = frames[i]
            if img is not None and img.shape == img0.shape and img.dtype == dtype:
                stage_frame(img, view)         # pageable -> pinned on a worker thread (numpy releases the GIL)
            return img

        def on_frame(_i, img):
            self._validate(img, shape2, dtype)

        try:
            self._stream(len(frames), load_into, ring, on_frame)
            return self._finish(len(frames), img0.shape[:2])
        except BaseException:
            if self._h is not None:
                self._lib.ss_reset(self._h)
            raise
