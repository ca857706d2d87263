"""``Vignetting`` sub-action with the radial divide on the GPU.

Drop-in for the reference's ``Vignetting`` (src/shinestacker/algorithms/vignetting.py:100-210) inside
``CombinedActions``: same constructor keywords (``r_steps``, ``black_threshold``, ``max_correction``, ``subsample``,
``fast_subsampling``, ``percentiles``, ``plot_correction``, ``plot_summary``) and protocol.

The FIT stays on the host, as in the reference: radial mean intensity of the sub-sampled 8-bit gray frame
(vignetting.py:23-41) and scipy's ``curve_fit`` of the sigmoid model (:44-53, :60-69).  The CORRECTION of the full
frame -- every pixel divided by the model at its radius, in float64, clipped and truncated (:71-97) -- runs as an
elementwise CUDA kernel on the uploaded frame (``sssub_apply_vignette``).  The kernel evaluates ``exp`` with CUDA's
float64 routine, numpy with the host's: results agree to 1 LSB (a last-ulp difference of the quotient can move a
sample across an integer boundary; measured: < 1e-5 of the samples), the tolerance written in the tests.
Plots are not produced.  No CPU fallback for the correction.
"""
import logging
import traceback

import numpy as np

from . import subactions
from .balance import subsampled

DEFAULT_R_STEPS, DEFAULT_BLACK_THRESHOLD, DEFAULT_MAX_CORRECTION, DEFAULT_VIGN_SUBSAMPLE = 100, 1.0, 1, 8
_CLIP = 10


def sigmoid_model(r, i0, k, r0):
    """vignetting.py:16-20."""
    return i0 / (1.0 + np.exp(np.minimum(_CLIP, np.exp(np.clip(k * (r - r0), -_CLIP, _CLIP)))))


def gray_subsampled(image, subsample, fast):
    """vignetting.py:56-58: 8-bit gray version of the frame, sub-sampled."""
    import cv2  # pylint: disable=import-outside-toplevel
    img8 = (image >> 8).astype('uint8') if image.dtype == np.uint16 else image
    return subsampled(cv2.cvtColor(img8, cv2.COLOR_BGR2GRAY), subsample, fast)


def radial_mean_intensity(gray, r_steps):
    """vignetting.py:23-41 -- mean intensity in ``r_steps`` rings around the image centre.  One pass: every pixel is
    assigned its ring (radii[i] <= d < radii[i+1]); sums of 8-bit samples are exact in float64 in any order, so the
    ring means equal ``np.mean(image[mask])``."""
    if gray.ndim > 2:
        raise ValueError("The image must be grayscale")
    h, w = gray.shape
    r_max = np.sqrt((w / 2) ** 2 + (h / 2) ** 2)
    radii = np.linspace(0, r_max, r_steps + 1)
    y, x = np.ogrid[:h, :w]
    dist = np.sqrt((x - w / 2) ** 2 + (y - h / 2) ** 2)
    ring = np.searchsorted(radii, dist.ravel(), side='right') - 1
    inside = (ring >= 0) & (ring < r_steps)
    counts = np.bincount(ring[inside], minlength=r_steps)
    sums = np.bincount(ring[inside], weights=gray.ravel()[inside].astype(np.float64), minlength=r_steps)
    with np.errstate(invalid='ignore', divide='ignore'):
        means = np.where(counts > 0, sums / counts, np.nan)
    return (radii[1:] + radii[:-1]) / 2, means


def fit_sigmoid(radii, intensities):
    """vignetting.py:44-53."""
    from scipy.optimize import curve_fit  # pylint: disable=import-outside-toplevel
    ok = ~np.isnan(intensities)
    i_ok, r_ok = intensities[ok], radii[ok]
    r_max = radii.max()
    return curve_fit(sigmoid_model, r_ok, i_ok, p0=[2 * np.max(i_ok), 10 / r_max, 0.8 * r_max],
                     bounds=([0, 0, 0], [np.inf, np.inf, np.inf]))[0]


def fit_parameters(image, r_steps=DEFAULT_R_STEPS, subsample=DEFAULT_VIGN_SUBSAMPLE, fast_subsampling=False):
    """vignetting.py:60-69: (i0, k, r0) in full-resolution pixel units."""
    radii, means = radial_mean_intensity(gray_subsampled(image, subsample, fast_subsampling), r_steps)
    params = fit_sigmoid(radii, means)
    params[1] /= subsample
    params[2] *= subsample
    return params


def correct_vignetting(image, max_correction=DEFAULT_MAX_CORRECTION, black_threshold=DEFAULT_BLACK_THRESHOLD,
                       r_steps=DEFAULT_R_STEPS, params=None, v0=None, subsample=DEFAULT_VIGN_SUBSAMPLE,
                       fast_subsampling=False, device=0):
    """vignetting.py:71-97 with the divide on the GPU."""
    if params is None:
        if r_steps is None:
            raise RuntimeError("Either r_steps or pars must not be None")
        params = fit_parameters(image, r_steps, subsample, fast_subsampling)
    if v0 is None:
        v0 = sigmoid_model(0, *params)
    frame = subactions.frame_for(image, device)
    return frame.upload(image).apply_vignette(params, v0, max_correction, black_threshold).download()


class Vignetting:
    """Sub-action (protocol of reference stack_framework.py:235-243)."""

    def __init__(self, enabled=True, percentiles=(0.05, 0.1, 0.25, 0.5, 0.75, 0.9, 0.95), **kwargs):
        self.enabled = enabled
        self.r_steps = kwargs.get('r_steps', DEFAULT_R_STEPS)
        self.black_threshold = kwargs.get('black_threshold', DEFAULT_BLACK_THRESHOLD)
        self.max_correction = kwargs.get('max_correction', DEFAULT_MAX_CORRECTION)
        self.subsample = kwargs.get('subsample', DEFAULT_VIGN_SUBSAMPLE)
        self.fast_subsampling = kwargs.get('fast_subsampling', False)
        self.plot_correction = kwargs.get('plot_correction', False)   # accepted, not produced
        self.plot_summary = kwargs.get('plot_summary', False)
        self.percentiles = np.sort(percentiles)
        self.device = int(kwargs.get('device', 0))
        self.process = None
        self.corrections = None
        self.v0 = self.r_max = self.w_2 = self.h_2 = None
        self.last_params = None

    def begin(self, process):
        self.process = process
        self.corrections = [np.full(process.counts, None, dtype=float) for _ in self.percentiles]

    def run_frame(self, idx, _ref_idx, img_0):
        from scipy.optimize import fsolve  # pylint: disable=import-outside-toplevel
        if self.process is not None:
            self.process.sub_message_r(": compute vignetting")
        h, w = img_0.shape[:2]
        self.w_2, self.h_2 = w / 2, h / 2
        self.r_max = np.sqrt((w / 2) ** 2 + (h / 2) ** 2)
        try:
            radii, means = radial_mean_intensity(gray_subsampled(img_0, self.subsample, self.fast_subsampling),
                                                 self.r_steps)
            params = fit_sigmoid(radii, means)
            params[1] /= self.subsample
            params[2] *= self.subsample
        except Exception as e:  # pylint: disable=broad-except
            traceback.print_tb(e.__traceback__)
            if self.process is not None:
                self.process.sub_message(": could not find vignetting model", level=logging.WARNING)
            return img_0
        self.last_params = params
        self.v0 = sigmoid_model(0, *params)
        if self.corrections is not None:
            for i, p in enumerate(self.percentiles):
                self.corrections[i][idx] = fsolve(lambda x, p=p: sigmoid_model(x, *params) / self.v0 - p, params[2])[0]
        if self.process is not None:
            self.process.sub_message_r(": correct vignetting")
        return correct_vignetting(img_0, self.max_correction, self.black_threshold, None, params, self.v0,
                                  self.subsample, self.fast_subsampling, self.device)

    def end(self):
        pass
