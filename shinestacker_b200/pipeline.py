"""Fused AlignFrames -> PyramidStack path (BASELINE.json config 5: "AlignFrames ECC + FocusStack end-to-end").

In the reference the two are separate actions of a job: ``CombinedActions([AlignFrames()])`` writes aligned
files, ``FocusStack`` reads them back (reference stack_framework.py:246-302, stack.py:14-113).  Here one call
keeps every frame in HBM between the two: upload once, fit the transform on the GPU (``ssal_ecc_estimate``),
warp on the GPU (``ssal_warp`` with the resident frame), hand the aligned frame to the stacker by device pointer
(``ss_push_frame_device``).  Frames are aligned to ONE reference frame (the reference's default
``ref_idx`` = middle frame, ``step_process`` off).
"""
import ctypes
from concurrent.futures import ThreadPoolExecutor

import numpy as np

from .align import Warper, ALIGN_RIGID, BORDER_REPLICATE_BLUR, ECC_MIN_RHO, AlignmentError
from .pyramid import PyramidStack, _check
from .exceptions import ImageLoadError

STAGE_SLOTS = 4         # pinned slots frames in pageable memory are copied into, ahead of the frame being aligned
STAGE_THREADS = 3


class _Sources:
    """Frame i as an array in page-locked memory, ready for an asynchronous upload: the caller's own array when it
    already is page-locked, else a copy made ahead of time by worker threads into a small ring of pinned slots."""

    def __init__(self, lib, frames):
        self.lib, self.frames = lib, frames
        self.pinned = [bool(f.flags.c_contiguous and lib.ss_host_is_pinned(ctypes.c_void_p(f.ctypes.data))) for f in frames]
        self.ring, self.views, self.pending, self.pool = [], [], {}, None
        if not all(self.pinned):
            nbytes = frames[0].nbytes
            for _ in range(STAGE_SLOTS):
                ptr = ctypes.c_void_p()
                _check(lib.ss_host_alloc(ctypes.byref(ptr), nbytes))
                self.ring.append(ptr)
                buf = (ctypes.c_uint8 * nbytes).from_address(ptr.value)
                self.views.append(np.frombuffer(buf, dtype=frames[0].dtype).reshape(frames[0].shape))
            self.pool = ThreadPoolExecutor(STAGE_THREADS)

    def stage_ahead(self, order, pos):
        """Submits the copies of positions pos .. pos + STAGE_SLOTS - 2 of ``order``.  Called for pos = k + 1 while
        frame k is the current one: every upload up to frame k - 1 has completed by then (the fit of k - 1 synchronised
        behind it), so the slots of positions <= k - 1, i.e. those of k + 1 .. k + STAGE_SLOTS - 1, are free."""
        if self.pool is None:
            return
        for k in range(pos, min(pos + STAGE_SLOTS - 1, len(order))):
            i = order[k]
            if not self.pinned[i] and k not in self.pending:
                self.pending[k] = self.pool.submit(np.copyto, self.views[k % STAGE_SLOTS], self.frames[i])

    def get(self, order, k):
        i = order[k]
        if self.pinned[i]:
            return self.frames[i]
        self.stage_ahead(order, k)
        self.pending.pop(k).result()
        return self.views[k % STAGE_SLOTS]

    def close(self):
        if self.pool is not None:
            self.pool.shutdown(wait=True)
        for ptr in self.ring:
            self.lib.ss_host_free(ptr)
        self.ring, self.views = [], []


def align_and_stack_arrays(frames, ref_idx=None, stack_algo=None, border_mode=BORDER_REPLICATE_BLUR,
                           border_value=(0, 0, 0, 0), border_blur=50, ecc_finest_level=None, ecc_max_iters=30,
                           ecc_eps=1e-3, device=0, warper=None, return_transforms=False):
    """Aligns every frame onto ``frames[ref_idx]`` (ECC similarity fit + warp, all on the GPU) and fuses the
    aligned stack with ``stack_algo`` (default ``PyramidStack()``); returns the fused image (and the list of
    2x3 forward transforms with ``return_transforms``).

    Pipelined per frame: while frame i is fitted and warped, frame i+1 uploads into the warper's second source
    buffer (``ssal_prefetch``) and the aligned frame i-1 is copied into the stack by the stacker's own stream (the
    warper's outputs are double-buffered, so nothing waits for that copy).

    ``ecc_finest_level``: pyramid level the fit stops at; default (None) = 1 for frames of at least 2048 pixels on the
    short side, else 0.  The reference estimates on a subsampled image by default as well (align.py:167-173,
    ``subsample``); on the 4096x3072 stack the two full-resolution iterations cost 0.73 ms of the 1.6 ms fit and
    change the error against the known transform from 0.0025 to 0.0023 px (tools/pipe_once.py)."""
    frames = list(frames)
    if not frames:
        raise ImageLoadError("", "empty frame list")
    n = len(frames)
    ref_idx = n // 2 if ref_idx is None else int(ref_idx)
    if ecc_finest_level is None:
        ecc_finest_level = 1 if min(frames[0].shape[:2]) >= 2048 else 0
    algo = stack_algo or PyramidStack(device=device)
    first = frames[0]
    algo._set_metadata(first)                                  # pylint: disable=protected-access
    algo._ensure_handle(first.shape, first.dtype, n)           # pylint: disable=protected-access
    lib, h = algo._lib, algo._h                                # pylint: disable=protected-access
    for img in frames:
        algo._validate(img, first.shape[:2], first.dtype)      # pylint: disable=protected-access
    wp = warper or Warper(first.shape[:2], first.dtype, device)
    transforms = []
    moving = [i for i in range(n) if i != ref_idx]             # the frames that go through the warper, in order
    src = _Sources(lib, frames)
    pitch = first.shape[1] * 3 * first.dtype.itemsize
    try:
        wp.ecc_set_reference(frames[ref_idx])
        if moving:
            wp.prefetch(src.get(moving, 0))
        k = 0
        for i in range(n):
            if i == ref_idx:
                m = np.array([[1.0, 0.0, 0.0], [0.0, 1.0, 0.0]])
                ring = algo._staging(first.nbytes, 2)          # pylint: disable=protected-access
                ptr, view = ring[0]
                _check(lib.ss_sync_copies(h))
                view[:first.nbytes] = np.ascontiguousarray(frames[i]).reshape(-1).view(np.uint8)
                _check(lib.ss_push_frame(h, ptr, pitch))
                _check(lib.ss_sync_copies(h))
            else:
                wp.use_prefetched()                            # frame i is the current one; its slot is free again
                if k + 1 < len(moving):
                    wp.prefetch(src.get(moving, k + 1))        # uploads behind the fit of frame i
                m, info = wp.ecc_estimate(None, ecc_finest_level, ecc_max_iters, ecc_eps)
                if not info["rho"] >= ECC_MIN_RHO:
                    raise AlignmentError(i, f"ecc correlation {info['rho']:.3f} < {ECC_MIN_RHO}")
                _check(lib.ss_sync_copies(h))                  # the copy of the frame before last left this buffer
                wp.warp(None, m, ALIGN_RIGID, border_mode, border_value, border_blur, download=False)
                d_ptr, d_pitch = wp.result_device()
                _check(lib.ss_push_frame_device(h, d_ptr, d_pitch))
                k += 1
            transforms.append(m)
        out = algo._finish(n, first.shape[:2])                 # pylint: disable=protected-access
    except BaseException:
        lib.ss_reset(h)
        for retire in (wp.use_prefetched, wp.sync):     # a frame may still be uploading out of a pinned slot that
            try:                                        # src.close() is about to free: retire it and wait
                retire()
            except Exception:  # pylint: disable=broad-except
                pass
        raise
    finally:
        src.close()
        if warper is None:
            wp.close()
    return (out, transforms) if return_transforms else out
