"""Host-side mirror of the reference's ``AlignFrames`` sub-action and ``align_images`` (reference
src/shinestacker/algorithms/align.py:170-252 and :254-353) with the image work on a B200.

Division of labour:

* the estimator -- feature detection / description, matching, RANSAC fit (align.py:48-168) -- stays on the
  host and calls the very same cv2 functions as the reference (SIFT / ORB / AKAZE / BRISK, FLANN or Hamming
  matcher, ``estimateAffinePartial2D`` / ``findHomography``), so the recovered transform is the reference's;
* what follows once the transform is known (align.py:228-251) -- ``cv2.warpAffine`` / ``cv2.warpPerspective``
  of the full-resolution frame, the validity mask of ``BORDER_REPLICATE_BLUR``, the 21x21 Gaussian blur and
  the paste into the invalid border -- runs in libshinestack_b200.so (``ssal_warp``,
  include/shinestack_b200_align.h), bit-exact against cv2 for uint8 and for uint16 + rigid transforms.

``AlignFrames`` keeps the sub-action protocol ``CombinedActions`` drives (stack_framework.py:253-302):
``enabled``, ``begin(process)``, ``run_frame(idx, ref_idx, img)``, ``end()``; same keyword handling, messages
and ``AlignmentError``.  There is no CPU warp here: without the library / an sm_100 GPU the warp raises
``NativeLibraryError``.
"""
import ctypes
import threading
import logging

import numpy as np

from . import _native
from .exceptions import InvalidOptionError, FocusStackError, ShapeError, BitDepthError, AlignmentError
from .pyramid import _check

# reference src/shinestacker/config/constants.py:64-110
ALIGN_HOMOGRAPHY, ALIGN_RIGID = "ALIGN_HOMOGRAPHY", "ALIGN_RIGID"
BORDER_CONSTANT, BORDER_REPLICATE, BORDER_REPLICATE_BLUR = "BORDER_CONSTANT", "BORDER_REPLICATE", "BORDER_REPLICATE_BLUR"
DETECTOR_SIFT, DETECTOR_ORB, DETECTOR_SURF, DETECTOR_AKAZE, DETECTOR_BRISK = "SIFT", "ORB", "SURF", "AKAZE", "BRISK"
DESCRIPTOR_SIFT, DESCRIPTOR_ORB, DESCRIPTOR_AKAZE, DESCRIPTOR_BRISK = "SIFT", "ORB", "AKAZE", "BRISK"
MATCHING_KNN, MATCHING_NORM_HAMMING = "KNN", "NORM_HAMMING"
ALIGN_RANSAC, ALIGN_LMEDS = "RANSAC", "LMEDS"
NOKNN_DETECTORS = (DETECTOR_ORB, DETECTOR_SURF, DETECTOR_AKAZE, DETECTOR_BRISK)
NOKNN_DESCRIPTORS = (DESCRIPTOR_ORB, DESCRIPTOR_AKAZE, DESCRIPTOR_BRISK)

DEFAULT_FEATURE_CONFIG = {'detector': DETECTOR_SIFT, 'descriptor': DESCRIPTOR_SIFT}
DEFAULT_MATCHING_CONFIG = {'match_method': MATCHING_KNN, 'flann_idx_kdtree': 2, 'flann_trees': 5,
                           'flann_checks': 50, 'threshold': 0.75}
DEFAULT_ALIGNMENT_CONFIG = {'transform': ALIGN_RIGID, 'align_method': ALIGN_RANSAC, 'rans_threshold': 3.0,
                            'refine_iters': 100, 'align_confidence': 99.9, 'max_iters': 2000,
                            'border_mode': BORDER_REPLICATE_BLUR, 'border_value': [0] * 4, 'border_blur': 50,
                            'subsample': 2, 'fast_subsampling': False, 'min_good_matches': 100}

# extra keys of alignment_config in this mirror (not reference options): which estimator fits the transform
# ('features' = the reference's cv2 pipeline on the host, 'ecc' = gradient-based fit on the GPU) and its knobs
ESTIMATOR_FEATURES, ESTIMATOR_ECC = "features", "ecc"
DEFAULT_ALIGNMENT_CONFIG.update({'estimator': ESTIMATOR_FEATURES, 'ecc_finest_level': 0, 'ecc_max_iters': 30,
                                 'ecc_eps': 1e-3})

_BORDER_CODE = {BORDER_CONSTANT: 0, BORDER_REPLICATE: 1, BORDER_REPLICATE_BLUR: 2}
_BLUR_KSIZE = 21      # align.py:249: cv2.GaussianBlur(img_warp, (21, 21), sigmaX=border_blur)


def _merged(defaults, override):
    cfg = dict(defaults)
    cfg.update(override or {})
    return cfg


# ---- estimator: the reference's own cv2 calls, on the host ----------------------------------------------------
def _gray8(img):
    """utils.img_bw_8bit (reference utils.py:33-42): 16-bit images drop their low byte, colour -> gray."""
    import cv2  # pylint: disable=import-outside-toplevel
    if img.dtype == np.uint16:
        img = (img >> 8).astype('uint8')
    if img.ndim == 3:
        return cv2.cvtColor(img, cv2.COLOR_BGR2GRAY)
    if img.ndim == 2:
        return img
    raise ValueError(f"Unsupported image format: {img.shape}")


def _subsampled(img, factor, fast):
    """utils.img_subsample (reference utils.py:78-86)."""
    import cv2  # pylint: disable=import-outside-toplevel
    if fast:
        return img[::factor, ::factor]
    return cv2.resize(img, (0, 0), fx=1 / factor, fy=1 / factor, interpolation=cv2.INTER_AREA)


def validate_align_config(detector, descriptor, match_method):
    """Same incompatibility rules and messages as reference align.py:69-86."""
    if descriptor == DESCRIPTOR_SIFT and match_method == MATCHING_NORM_HAMMING:
        raise ValueError("Descriptor SIFT requires matching method KNN")
    if detector == DETECTOR_ORB and descriptor == DESCRIPTOR_AKAZE and match_method == MATCHING_NORM_HAMMING:
        raise ValueError("Detector ORB and descriptor AKAZE require matching method KNN")
    if detector == DETECTOR_BRISK and descriptor == DESCRIPTOR_AKAZE:
        raise ValueError("Detector BRISK is incompatible with descriptor AKAZE")
    if detector == DETECTOR_SURF and descriptor == DESCRIPTOR_AKAZE:
        raise ValueError("Detector SURF is incompatible with descriptor AKAZE")
    if detector == DETECTOR_SIFT and descriptor != DESCRIPTOR_SIFT:
        raise ValueError("Detector SIFT requires descriptor SIFT")
    if detector in NOKNN_DETECTORS and descriptor in NOKNN_DESCRIPTORS and match_method != MATCHING_NORM_HAMMING:
        raise ValueError(f"Detector {detector} and descriptor {descriptor}"
                         " require matching method Hamming distance")


def get_good_matches(des_0, des_1, matching_config=None):
    """reference align.py:48-66: Lowe-ratio filtered FLANN k-NN matches, or cross-checked Hamming matches."""
    import cv2  # pylint: disable=import-outside-toplevel
    cfg = _merged(DEFAULT_MATCHING_CONFIG, matching_config)
    method = cfg['match_method']
    if method == MATCHING_KNN:
        matcher = cv2.FlannBasedMatcher({'algorithm': cfg['flann_idx_kdtree'], 'trees': cfg['flann_trees']},
                                        {'checks': cfg['flann_checks']})
        return [best for best, second in matcher.knnMatch(des_0, des_1, k=2)
                if best.distance < cfg['threshold'] * second.distance]
    if method == MATCHING_NORM_HAMMING:
        matcher = cv2.BFMatcher(cv2.NORM_HAMMING, crossCheck=True)
        return sorted(matcher.match(des_0, des_1), key=lambda mt: mt.distance)
    raise InvalidOptionError('match_method', method,
                             f". Valid options are: {MATCHING_KNN}, {MATCHING_NORM_HAMMING}")


def detect_and_compute(img_0, img_1, feature_config=None, matching_config=None):
    """reference align.py:89-121: key points + descriptors of both images (8-bit gray) and the good matches."""
    import cv2  # pylint: disable=import-outside-toplevel
    fcfg = _merged(DEFAULT_FEATURE_CONFIG, feature_config)
    mcfg = _merged(DEFAULT_MATCHING_CONFIG, matching_config)
    det_name, des_name = fcfg['detector'], fcfg['descriptor']
    validate_align_config(det_name, des_name, mcfg['match_method'])
    factories = {DETECTOR_SIFT: cv2.SIFT_create, DETECTOR_ORB: cv2.ORB_create,
                 DETECTOR_SURF: cv2.FastFeatureDetector_create, DETECTOR_AKAZE: cv2.AKAZE_create,
                 DETECTOR_BRISK: cv2.BRISK_create}
    grays = (_gray8(img_0), _gray8(img_1))
    detector = factories[det_name]()
    if det_name == des_name and det_name in (DETECTOR_SIFT, DETECTOR_AKAZE, DETECTOR_BRISK):
        (kp_0, des_0), (kp_1, des_1) = (detector.detectAndCompute(g, None) for g in grays)
    else:
        describer = {k: factories[k] for k in (DESCRIPTOR_SIFT, DESCRIPTOR_ORB, DESCRIPTOR_AKAZE,
                                               DESCRIPTOR_BRISK)}[des_name]()
        (kp_0, des_0), (kp_1, des_1) = (describer.compute(g, detector.detect(g, None)) for g in grays)
    return kp_0, kp_1, get_good_matches(des_0, des_1, mcfg)


def find_transform(src_pts, dst_pts, transform=ALIGN_RIGID, method=ALIGN_RANSAC, rans_threshold=3.0,
                   max_iters=2000, align_confidence=99.9, refine_iters=100):
    """reference align.py:124-148."""
    import cv2  # pylint: disable=import-outside-toplevel
    try:
        cv2_method = {ALIGN_RANSAC: cv2.RANSAC, ALIGN_LMEDS: cv2.LMEDS}[method]
    except KeyError as e:
        raise InvalidOptionError('align_method', method,
                                 f". Valid options are: {ALIGN_RANSAC}, {ALIGN_LMEDS}") from e
    if transform == ALIGN_HOMOGRAPHY:
        return cv2.findHomography(src_pts, dst_pts, method=cv2_method, ransacReprojThreshold=rans_threshold,
                                  maxIters=max_iters)
    if transform == ALIGN_RIGID:
        return cv2.estimateAffinePartial2D(src_pts, dst_pts, method=cv2_method,
                                           ransacReprojThreshold=rans_threshold,
                                           confidence=align_confidence / 100.0, refineIters=refine_iters)
    raise InvalidOptionError("transform", transform)


def rescale_transform(m, transform, subsample, full_shape, sub_shape):
    """Transform estimated on subsampled images -> full resolution (reference align.py:207-222)."""
    import cv2  # pylint: disable=import-outside-toplevel
    if subsample <= 1:
        return m
    h, w = full_shape[:2]
    h_sub, w_sub = sub_shape[:2]
    if transform == ALIGN_HOMOGRAPHY:
        low = np.float32([[0, 0], [0, h_sub], [w_sub, h_sub], [w_sub, 0]])
        high = np.float32([[0, 0], [0, h], [w, h], [w, 0]])
        return cv2.getPerspectiveTransform(low, high) @ m @ cv2.getPerspectiveTransform(high, low)
    if transform == ALIGN_RIGID:
        full = np.empty((2, 3), dtype=np.float32)
        full[:2, :2] = m[:2, :2]
        full[:, 2] = m[:, 2] * subsample
        return full
    raise InvalidOptionError("transform", transform)


# ---- warp stage on the GPU -------------------------------------------------------------------------------------
def fixed_point_blur_kernel(dtype, sigma, ksize=_BLUR_KSIZE):
    """Integer kernel of cv2.GaussianBlur((ksize, ksize), sigmaX=sigma) for uint8 (8 fractional bits) / uint16
    (16), read out of the installed cv2 with a full-scale impulse: every tap but the centre comes back exactly,
    the centre is what is missing to 1.0.  Using cv2's own kernel keeps the device blur bit-identical to the
    reference whatever cv2 build is installed."""
    import cv2  # pylint: disable=import-outside-toplevel
    dtype = np.dtype(dtype)
    one = 256 if dtype == np.uint8 else 65536
    impulse = np.zeros((1, 4 * ksize + 1), dtype)
    centre = 2 * ksize
    impulse[0, centre] = np.iinfo(dtype).max
    resp = cv2.GaussianBlur(impulse, (ksize, 1), sigmaX=sigma, sigmaY=sigma)[0].astype(np.int64)
    taps = resp[centre - ksize // 2:centre + ksize // 2 + 1].copy()
    taps[ksize // 2] = one - (taps.sum() - taps[ksize // 2])
    return np.ascontiguousarray(taps, dtype=np.int32)


class Warper:
    """ctypes wrapper of one ``ssal_handle``: warps frames of one geometry on one device."""

    def __init__(self, shape, dtype, device=0):
        self._lib = _native.load()
        self.shape = (int(shape[0]), int(shape[1]))
        self.dtype = np.dtype(dtype)
        if self.dtype not in (np.uint8, np.uint16):
            raise BitDepthError(np.uint8, self.dtype)
        hp = ctypes.c_void_p()
        _check(self._lib.ssal_create(self.shape[1], self.shape[0], self.dtype.itemsize, int(device), ctypes.byref(hp)))
        self._h = hp
        self._kernels = {}
        self.last_kernel_ms = None
        self.ecc_generation = 0     # bumped by every ecc_set_reference: identifies the reference pyramid now in HBM

    def close(self):
        if getattr(self, "_h", None) is not None:
            self._lib.ssal_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:  # pylint: disable=broad-except
            pass

    def warp(self, img, m, transform=ALIGN_RIGID, border_mode=BORDER_REPLICATE_BLUR, border_value=(0, 0, 0, 0),
             border_blur=50, download=True):
        """align.py:228-251 for one frame; returns the aligned image (or None with download=False: the frame
        then stays in HBM, see ``result_device``)."""
        if img is not None:     # None: warp the frame the last ecc_estimate uploaded
            self._check_frame(img)
        try:
            mode = _BORDER_CODE[border_mode]
        except KeyError as e:
            raise InvalidOptionError("border_mode", border_mode) from e
        if transform == ALIGN_RIGID:
            code, mat = 0, np.ascontiguousarray(np.asarray(m, dtype=np.float64).reshape(2, 3))
        elif transform == ALIGN_HOMOGRAPHY:
            code, mat = 1, np.ascontiguousarray(np.asarray(m, dtype=np.float64).reshape(3, 3))
        else:
            raise InvalidOptionError("transform", transform)
        taps, ktab = 0, None
        if mode == 2:
            key = float(border_blur)
            if key not in self._kernels:
                self._kernels[key] = fixed_point_blur_kernel(self.dtype, key)
            ktab = self._kernels[key]
            taps = len(ktab)
        bval = np.zeros(4, np.float64)
        bv = np.asarray(border_value, np.float64).ravel()
        bval[:min(4, bv.size)] = bv[:4]
        if img is not None:
            img = np.ascontiguousarray(img)
        out = np.empty(self.shape + (3,), self.dtype) if download else None
        _check(self._lib.ssal_warp(
            self._h, img.ctypes.data_as(ctypes.c_void_p) if img is not None else None,
            img.strides[0] if img is not None else 0,
            mat.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), code, mode,
            bval.ctypes.data_as(ctypes.POINTER(ctypes.c_double)),
            ktab.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)) if ktab is not None else None, taps,
            out.ctypes.data_as(ctypes.c_void_p) if download else None, out.strides[0] if download else 0))
        ms = ctypes.c_float()
        _check(self._lib.ssal_kernel_time_ms(self._h, ctypes.byref(ms)))
        self.last_kernel_ms = ms.value
        return out

    # ---- gradient-based estimator (ssal_ecc_*) -----------------------------------------------------------------
    def ecc_set_reference(self, img):
        """The frame the following frames are aligned to (uploaded once; its gray pyramid stays in HBM)."""
        self._check_frame(img)
        img = np.ascontiguousarray(img)
        _check(self._lib.ssal_ecc_set_reference(self._h, img.ctypes.data_as(ctypes.c_void_p), img.strides[0]))
        self.ecc_generation += 1
        return self.ecc_generation

    def ecc_estimate(self, img, finest_level=0, max_iters=30, eps_px=1e-3, init_m=None):
        """Forward 2x3 similarity (moving -> reference frame, like cv2.estimateAffinePartial2D) of ``img`` by
        ECC iterations on the GPU; returns ``(m, info)`` with info = dict(rho, iterations, overlap, last_step).
        The frame stays on the device: ``warp(None, m, ...)`` then warps it without a second upload."""
        if img is not None:     # None: the frame made current by use_prefetched()
            self._check_frame(img)
            img = np.ascontiguousarray(img)
        m = np.zeros(6, np.float64)
        info = np.zeros(4, np.float64)
        init = None if init_m is None else np.ascontiguousarray(np.asarray(init_m, np.float64).reshape(6))
        dp = ctypes.POINTER(ctypes.c_double)
        _check(self._lib.ssal_ecc_estimate(self._h, img.ctypes.data_as(ctypes.c_void_p) if img is not None else None,
                                           img.strides[0] if img is not None else 0,
                                           int(finest_level), int(max_iters), float(eps_px),
                                           init.ctypes.data_as(dp) if init is not None else None,
                                           m.ctypes.data_as(dp), info.ctypes.data_as(dp)))
        ms = ctypes.c_float()
        _check(self._lib.ssal_kernel_time_ms(self._h, ctypes.byref(ms)))
        self.last_kernel_ms = ms.value
        return m.reshape(2, 3), dict(rho=info[0], iterations=int(info[1]), overlap=info[2], last_step=info[3])

    def prefetch(self, img):
        """Starts the upload of the NEXT frame (``img``: C-contiguous array in page-locked host memory, which the
        caller keeps alive and unchanged until ``use_prefetched``) behind the work on the current one."""
        self._check_frame(img)
        if not img.flags.c_contiguous:
            raise ValueError("prefetch needs a C-contiguous frame")
        _check(self._lib.ssal_prefetch(self._h, img.ctypes.data_as(ctypes.c_void_p), img.strides[0]))

    def use_prefetched(self):
        """Makes the prefetched frame the current one: ``ecc_estimate(None)`` / ``warp(None, ...)`` then work on it."""
        _check(self._lib.ssal_use_prefetched(self._h))

    def sync(self):
        """Waits for everything queued on this warper, uploads of prefetched frames included."""
        _check(self._lib.ssal_sync(self._h))

    def _check_frame(self, img):
        if img.shape[:2] != self.shape or img.ndim != 3 or img.shape[2] != 3:
            raise ShapeError(self.shape, img.shape[:2])
        if img.dtype != self.dtype:
            raise BitDepthError(self.dtype, img.dtype)

    def result_device(self):
        """(device pointer, row pitch) of the last aligned frame, for ss_push_frame_device."""
        ptr, pitch = ctypes.c_void_p(), ctypes.c_size_t()
        _check(self._lib.ssal_result_device(self._h, ctypes.byref(ptr), ctypes.byref(pitch)))
        return ptr, pitch.value


_WARPERS = {}
_WARPERS_LOCK = threading.Lock()


_SELF_CHECKED = False


def _self_check(device=0):
    """The warp stage promises cv2's results bit for bit, which ties it to the fixed-point kernels of the installed
    cv2 build (validated: 4.13).  Once per process a small random tile goes through cv2.warpAffine and through the
    device; a mismatch does not stop the run but is reported."""
    global _SELF_CHECKED  # pylint: disable=global-statement
    if _SELF_CHECKED:
        return
    _SELF_CHECKED = True
    try:
        import cv2  # pylint: disable=import-outside-toplevel
        rng = np.random.default_rng(7)
        tile = rng.integers(0, 256, (48, 80, 3)).astype(np.uint8)
        m = np.array([[0.9993, -0.012, 1.37], [0.012, 0.9993, -0.81]])
        wp = Warper(tile.shape[:2], tile.dtype, device)
        try:
            got = wp.warp(tile, m, ALIGN_RIGID, BORDER_REPLICATE)
        finally:
            wp.close()
        expect = cv2.warpAffine(tile, m, (80, 48), borderMode=cv2.BORDER_REPLICATE)
        if not np.array_equal(got, expect):
            import warnings  # pylint: disable=import-outside-toplevel
            warnings.warn(f"shinestacker_b200.align: the GPU warp differs from cv2 {cv2.__version__}.warpAffine on a "
                          f"test tile ({int((got != expect).sum())} samples); the bit-exact contract was validated "
                          "against cv2 4.13 -- results may differ from the reference by a few LSB", RuntimeWarning)
    except Exception:  # pylint: disable=broad-except
        pass


def _warper_for(img, device=0):
    """The process-wide warper of this geometry (shared by every caller: whoever relies on state inside it -- the ECC
    reference pyramid -- must check ``ecc_generation``, see AlignFrames.run_frame)."""
    _self_check(device)
    key = (img.shape[0], img.shape[1], img.dtype.str, device)
    with _WARPERS_LOCK:
        if key not in _WARPERS:
            if len(_WARPERS) >= 4:      # a handful of geometries at most stay resident
                _WARPERS.pop(next(iter(_WARPERS))).close()
            _WARPERS[key] = Warper(img.shape[:2], img.dtype, device)
        return _WARPERS[key]


def align_images(img_1, img_0, feature_config=None, matching_config=None, alignment_config=None,
                 plot_path=None, callbacks=None, device=0):
    """reference align.py:170-252: align ``img_0`` onto the reference frame ``img_1``.

    Returns ``(n_good_matches, m, img_warp)`` like the reference (``m``/``img_warp`` are None when there are too
    few matches).  ``plot_path`` (a matches plot, matplotlib) is not produced by this mirror.
    ``alignment_config['estimator'] = 'ecc'`` (not a reference option) replaces the feature pipeline by the
    gradient-based estimator on the GPU (ALIGN_RIGID only); the first return value is then the number of
    overlapping pixels."""
    fcfg = _merged(DEFAULT_FEATURE_CONFIG, feature_config)
    mcfg = _merged(DEFAULT_MATCHING_CONFIG, matching_config)
    acfg = _merged(DEFAULT_ALIGNMENT_CONFIG, alignment_config)
    callbacks = callbacks or {}
    if acfg['border_mode'] not in _BORDER_CODE:
        raise InvalidOptionError("border_mode", acfg['border_mode'])
    transform = acfg['transform']
    min_matches = 4 if transform == ALIGN_HOMOGRAPHY else 3
    if img_0 is None:
        raise RuntimeError("Image is None")
    if img_0.shape[:2] != img_1.shape[:2]:
        raise ShapeError(img_1.shape[:2], img_0.shape[:2])
    if img_0.dtype != img_1.dtype:
        raise BitDepthError(img_1.dtype, img_0.dtype)
    if 'message' in callbacks:
        callbacks['message']()
    if acfg.get('estimator', ESTIMATOR_FEATURES) == ESTIMATOR_ECC:
        return _align_images_ecc(img_1, img_0, acfg, callbacks, device)
    if acfg.get('estimator', ESTIMATOR_FEATURES) != ESTIMATOR_FEATURES:
        raise InvalidOptionError("estimator", acfg['estimator'],
                                 f". Valid options are: {ESTIMATOR_FEATURES}, {ESTIMATOR_ECC}")
    subsample = acfg['subsample']
    while True:
        if subsample > 1:
            sub_0 = _subsampled(img_0, subsample, acfg['fast_subsampling'])
            sub_1 = _subsampled(img_1, subsample, acfg['fast_subsampling'])
        else:
            sub_0, sub_1 = img_0, img_1
        kp_0, kp_1, good = detect_and_compute(sub_0, sub_1, fcfg, mcfg)
        n_good = len(good)
        if n_good > acfg['min_good_matches'] or subsample == 1:
            break
        subsample = 1
        if 'warning' in callbacks:
            callbacks['warning'](f"only {n_good} < {acfg['min_good_matches']} matches found, "
                                 "retrying without subsampling")
    if 'matches_message' in callbacks:
        callbacks['matches_message'](n_good)
    if n_good < min_matches:
        return n_good, None, None
    src_pts = np.float32([kp_0[g.queryIdx].pt for g in good]).reshape(-1, 1, 2)
    dst_pts = np.float32([kp_1[g.trainIdx].pt for g in good]).reshape(-1, 1, 2)
    m, _inliers = find_transform(src_pts, dst_pts, transform, acfg['align_method'], acfg['rans_threshold'],
                                 acfg['max_iters'], acfg['align_confidence'], acfg['refine_iters'])
    m = rescale_transform(m, transform, subsample, img_0.shape, sub_0.shape)
    if 'align_message' in callbacks:
        callbacks['align_message']()
    if acfg['border_mode'] == BORDER_REPLICATE_BLUR and 'blur_message' in callbacks:
        callbacks['blur_message']()
    img_warp = _warper_for(img_0, device).warp(img_0, m, transform, acfg['border_mode'], acfg['border_value'],
                                               acfg['border_blur'])
    return n_good, m, img_warp


ECC_MIN_RHO = 0.2     # below this correlation the ECC fit is reported as a failed alignment


def _align_images_ecc(img_1, img_0, acfg, callbacks, device, reference_is_set=False):
    """align_images with the gradient-based estimator: both the fit and the warp run on the GPU, the moving
    frame is uploaded once.  Returns (overlapping pixels, m, img_warp) or (0, None, None) when the fit fails."""
    if acfg['transform'] != ALIGN_RIGID:
        raise InvalidOptionError("transform", acfg['transform'],
                                 details=" the ecc estimator fits the ALIGN_RIGID similarity only")
    wp = _warper_for(img_0, device)
    if not reference_is_set:
        wp.ecc_set_reference(img_1)
    try:
        m, info = wp.ecc_estimate(img_0, finest_level=acfg.get('ecc_finest_level', 0),
                                  max_iters=acfg.get('ecc_max_iters', 30), eps_px=acfg.get('ecc_eps', 1e-3))
    except FocusStackError as e:
        if 'warning' in callbacks:
            callbacks['warning'](f"ecc estimator failed: {e}")
        return 0, None, None
    n_px = int(info['overlap'] * img_0.shape[0] * img_0.shape[1])
    if 'matches_message' in callbacks:
        callbacks['matches_message'](n_px)
    if not info['rho'] >= ECC_MIN_RHO:
        return 0, None, None
    if 'align_message' in callbacks:
        callbacks['align_message']()
    if acfg['border_mode'] == BORDER_REPLICATE_BLUR and 'blur_message' in callbacks:
        callbacks['blur_message']()
    out = wp.warp(None, m, ALIGN_RIGID, acfg['border_mode'], acfg['border_value'], acfg['border_blur'])
    return n_px, m, out


class AlignFrames:
    """Sub-action with the interface of the reference ``AlignFrames`` (align.py:254-353; ``SubAction``,
    stack_framework.py:235-243).  ``plot_summary`` / ``plot_matches`` are accepted and ignored."""

    def __init__(self, enabled=True, feature_config=None, matching_config=None, alignment_config=None,
                 device=0, **kwargs):
        self.enabled = enabled
        self.process = None
        self.n_matches = None
        self.device = int(device)
        self.feature_config = _merged(DEFAULT_FEATURE_CONFIG, feature_config)
        self.matching_config = _merged(DEFAULT_MATCHING_CONFIG, matching_config)
        self.alignment_config = _merged(DEFAULT_ALIGNMENT_CONFIG, alignment_config)
        self.plot_summary = kwargs.get('plot_summary', False)
        self.plot_matches = kwargs.get('plot_matches', False)
        for cfg in (self.feature_config, self.matching_config, self.alignment_config):
            for key in cfg:
                if key in kwargs:
                    cfg[key] = kwargs[key]
        self.min_matches = 4 if self.alignment_config['transform'] == ALIGN_HOMOGRAPHY else 3
        self._ecc_ref = None

    def begin(self, process):
        self.process = process
        self.n_matches = np.zeros(process.counts)
        self._ecc_ref = None

    def end(self):
        pass

    def sub_msg(self, msg):
        self.process.sub_message_r(msg)

    def run_frame(self, idx, ref_idx, img_0):
        if idx == self.process.ref_idx:
            return img_0
        if self.alignment_config['estimator'] == ESTIMATOR_ECC and self._ecc_ref is not None:
            # the reference pyramid of ref_idx may still be in HBM -- but the warper is shared process-wide: only if
            # it is the same object and nobody has replaced its reference since we set it
            wp = _warper_for(img_0, self.device)
            if self._ecc_ref == (ref_idx, img_0.shape, img_0.dtype, id(wp), wp.ecc_generation):
                return self.align_images(idx, None, img_0)
        return self.align_images(idx, self.process.img_ref(ref_idx), img_0, ref_idx)

    def align_images(self, idx, img_1, img_0, ref_idx=None):
        callbacks = {
            'message': lambda: self.sub_msg(': find matches'),
            'matches_message': lambda n: self.sub_msg(f": good matches: {n}"),
            'align_message': lambda: self.sub_msg(': align images'),
            'blur_message': lambda: self.sub_msg(': blur borders'),
            'warning': lambda msg: self.sub_msg(f': {msg}'),
        }
        if self.alignment_config['estimator'] == ESTIMATOR_ECC:
            callbacks['message']()
            n_good, _m, img = _align_images_ecc(img_1, img_0, self.alignment_config, callbacks, self.device,
                                                reference_is_set=img_1 is None)
            if img_1 is not None and ref_idx is not None:
                wp = _warper_for(img_0, self.device)
                self._ecc_ref = (ref_idx, img_0.shape, img_0.dtype, id(wp), wp.ecc_generation)
        else:
            n_good, _m, img = align_images(img_1, img_0, feature_config=self.feature_config,
                                           matching_config=self.matching_config,
                                           alignment_config=self.alignment_config, callbacks=callbacks,
                                           device=self.device)
        self.n_matches[idx] = n_good
        if n_good < self.min_matches:
            self.process.sub_message(f": image not aligned, too few matches found: {n_good}",
                                     level=logging.CRITICAL)
            raise AlignmentError(idx, f"too few matches found: {n_good} < {self.min_matches}")
        return img
