"""Fused AlignFrames -> PyramidStack path (BASELINE.json config 5: "AlignFrames ECC + FocusStack end-to-end").

In the reference the two are separate actions of a job: ``CombinedActions([AlignFrames()])`` writes aligned
files, ``FocusStack`` reads them back (reference stack_framework.py:246-302, stack.py:14-113).  Here one call
keeps every frame in HBM between the two: upload once, fit the transform on the GPU (``ssal_ecc_estimate``),
warp on the GPU (``ssal_warp`` with the resident frame), hand the aligned frame to the stacker by device pointer
(``ss_push_frame_device``).  Frames are aligned to ONE reference frame (the reference's default
``ref_idx`` = middle frame, ``step_process`` off).
"""
import numpy as np

from .align import Warper, ALIGN_RIGID, BORDER_REPLICATE_BLUR, ECC_MIN_RHO, AlignmentError
from .pyramid import PyramidStack, _check
from .exceptions import ImageLoadError


def align_and_stack_arrays(frames, ref_idx=None, stack_algo=None, border_mode=BORDER_REPLICATE_BLUR,
                           border_value=(0, 0, 0, 0), border_blur=50, ecc_finest_level=0, ecc_max_iters=30,
                           ecc_eps=1e-3, device=0, warper=None, return_transforms=False):
    """Aligns every frame onto ``frames[ref_idx]`` (ECC similarity fit + warp, all on the GPU) and fuses the
    aligned stack with ``stack_algo`` (default ``PyramidStack()``); returns the fused image (and the list of
    2x3 forward transforms with ``return_transforms``)."""
    frames = list(frames)
    if not frames:
        raise ImageLoadError("", "empty frame list")
    n = len(frames)
    ref_idx = n // 2 if ref_idx is None else int(ref_idx)
    algo = stack_algo or PyramidStack(device=device)
    first = frames[0]
    algo._set_metadata(first)                                  # pylint: disable=protected-access
    algo._ensure_handle(first.shape, first.dtype, n)           # pylint: disable=protected-access
    lib, h = algo._lib, algo._h                                # pylint: disable=protected-access
    wp = warper or Warper(first.shape[:2], first.dtype, device)
    transforms = []
    try:
        wp.ecc_set_reference(frames[ref_idx])
        for i, img in enumerate(frames):
            algo._validate(img, first.shape[:2], first.dtype)  # pylint: disable=protected-access
            if i == ref_idx:
                m = np.array([[1.0, 0.0, 0.0], [0.0, 1.0, 0.0]])
                ring = algo._staging(img.nbytes, 2)            # pylint: disable=protected-access
                ptr, view = ring[0]
                _check(lib.ss_sync_copies(h))
                view[:img.nbytes] = np.ascontiguousarray(img).reshape(-1).view(np.uint8)
                _check(lib.ss_push_frame(h, ptr, img.shape[1] * 3 * img.dtype.itemsize))
                _check(lib.ss_sync_copies(h))
            else:
                m, info = wp.ecc_estimate(img, ecc_finest_level, ecc_max_iters, ecc_eps)
                if not info["rho"] >= ECC_MIN_RHO:
                    raise AlignmentError(i, f"ecc correlation {info['rho']:.3f} < {ECC_MIN_RHO}")
                wp.warp(None, m, ALIGN_RIGID, border_mode, border_value, border_blur, download=False)
                d_ptr, pitch = wp.result_device()
                _check(lib.ss_push_frame_device(h, d_ptr, pitch))
                _check(lib.ss_sync_copies(h))                  # the warper's output buffer is reused next
            transforms.append(m)
        out = algo._finish(n, first.shape[:2])                 # pylint: disable=protected-access
    except BaseException:
        lib.ss_reset(h)
        raise
    finally:
        if warper is None:
            wp.close()
    return (out, transforms) if return_transforms else out
