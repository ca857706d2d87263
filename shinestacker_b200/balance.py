"""``BalanceFrames`` sub-action with the table look-up on the GPU.

Drop-in for the reference's ``BalanceFrames`` (src/shinestacker/algorithms/balance.py:366-416) inside
``CombinedActions`` (stack_framework.py:246-302): same constructor keywords (``corr_map``, ``channel``, ``subsample``,
``fast_subsampling``, ``mask_size``, ``intensity_interval``, ``plot_histograms``, ``plot_summary``), same
``begin(process)`` / ``run_frame(idx, ref_idx, image)`` / ``end()`` protocol, same ``InvalidOptionError``.

Split of the work, as SURVEY.md prescribes: the FIT -- histograms of the (sub-sampled, optionally masked) frame, the
scalar solve for the linear factor / gamma / histogram match and the construction of the 256- or 65536-entry
table (balance.py:54-121) -- is a few kilobytes of scalar host work and is computed here with numpy / scipy exactly
as the reference does; the APPLICATION of the table to every sample of the frame (balance.py:27-48:
``cv2.LUT`` / ``np.take``) runs as an elementwise CUDA kernel on the uploaded frame (``sssub_apply_lut``).
The matplotlib plots of the reference are not produced (``plot_histograms`` / ``plot_summary`` are accepted and
ignored).  No CPU fallback for the table look-up.
"""
import numpy as np

from .exceptions import InvalidOptionError
from .pyramid import read_img
from . import subactions

# reference src/shinestacker/config/constants.py:112-127
BALANCE_LINEAR, BALANCE_GAMMA, BALANCE_MATCH_HIST = "LINEAR", "GAMMA", "MATCH_HIST"
BALANCE_LUMI, BALANCE_RGB, BALANCE_HSV, BALANCE_HLS = "LUMI", "RGB", "HSV", "HLS"
DEFAULT_BALANCE_SUBSAMPLE = 8


def subsampled(img, factor, fast):
    """reference utils.py img_subsample: plain striding, or cv2's area resize."""
    if factor == 1:
        return img
    if fast:
        return img[::factor, ::factor]
    import cv2  # pylint: disable=import-outside-toplevel
    return cv2.resize(img, (0, 0), fx=1 / factor, fy=1 / factor, interpolation=cv2.INTER_AREA)


def channel_histogram(plane, num_values, factor, fast, mask_size):
    """balance.py:160-181: histogram of one (sub-sampled, optionally disc-masked) plane, one bin per value."""
    sub = subsampled(plane, factor, fast)
    if mask_size != 0:
        h, w = sub.shape[:2]
        xv, yv = np.meshgrid(np.linspace(0, w - 1, w), np.linspace(0, h - 1, h))
        radius = min(w, h) * mask_size / 2
        sub = sub[(xv - w / 2) ** 2 + (yv - h / 2) ** 2 <= radius ** 2]
    # integer samples: counting them equals np.histogram over the half-integer bin edges the reference uses
    return np.bincount(np.asarray(sub).ravel(), minlength=num_values)[:num_values].astype(np.int64)


class _Tables:
    """The three correction maps of the reference (balance.py:14-121), reduced to what the sub-action needs:
    ``solve(hist)`` -> correction per channel, ``table(correction, c)`` -> look-up table, ``size(correction)``."""

    def __init__(self, kind, dtype, ref_hist, interval):
        self.kind, self.dtype = kind, np.dtype(dtype)
        self.nv = 256 if self.dtype == np.uint8 else 65536
        self.maxv = self.nv - 1
        self.values = np.arange(self.nv)
        interval = {"min": 0, "max": -1, **(interval or {})}
        self.lo = interval["min"]
        self.hi = interval["max"] + 1 if interval["max"] >= 0 else self.nv
        if kind == BALANCE_MATCH_HIST:
            self.ref = [self._cdf(h) for h in ref_hist]
            self.ref_mean = [r.mean() for r in self.ref]
        else:
            self.ref = [self._centre(self.values, h) for h in ref_hist]

    def _centre(self, table, hist):
        # balance.py:90-91: histogram-weighted mean of the mapped values inside the intensity interval
        return np.average(table[self.lo:self.hi], weights=np.asarray(hist).flatten()[self.lo:self.hi])

    def _cdf(self, hist):
        return np.cumsum(hist) / hist.sum() * self.maxv                     # balance.py:61-62

    def _gamma_table(self, gamma):
        return (((self.values / self.maxv) ** (1.0 / gamma)) * self.maxv).astype(self.dtype)   # balance.py:102-106

    def solve(self, hist):
        if self.kind == BALANCE_LINEAR:                                      # balance.py:120-121
            return [r / self._centre(self.values, h) for h, r in zip(hist, self.ref)]
        if self.kind == BALANCE_GAMMA:                                       # balance.py:98-100
            from scipy.optimize import bisect  # pylint: disable=import-outside-toplevel
            return [bisect(lambda g, h=h, r=r: self._centre(self._gamma_table(g), h) - r, 0.1, 5)
                    for h, r in zip(hist, self.ref)]
        return [self._cdf(h) for h in hist]                                  # balance.py:81-82

    def table(self, correction, c):
        if self.kind == BALANCE_LINEAR:                                      # balance.py:116-118
            return np.clip(self.values * correction, 0, self.maxv).astype(self.dtype)
        if self.kind == BALANCE_GAMMA:
            return self._gamma_table(correction)
        return self._match_table(correction, self.ref[c])

    def _match_table(self, cdf, ref_cdf):
        """balance.py:64-79: map every value through the inverse of the reference's cumulative histogram, then spread
        the two saturated ends linearly."""
        from scipy.interpolate import interp1d  # pylint: disable=import-outside-toplevel
        inverse = interp1d(ref_cdf, self.values)
        t = np.array([inverse(v) for v in np.clip(cdf, ref_cdf.min(), ref_cdf.max())])
        first, last = t[0], t[-1]
        inner = t[(t != first) & (t != last)]
        if inner.size > 0:
            lo, hi = inner.min(), inner.max()
            at_first, at_last = self.values[t == first], self.values[t == last]
            top = at_first.max()
            t[t == first] = (at_first / top * lo) if top > 0 else 0
            t[t == last] = at_last + (at_last - self.maxv) * (self.maxv - hi) / float(at_last.size) \
                if at_last.size > 0 else self.maxv
        return t.astype(self.dtype)

    def size(self, correction):
        if self.kind == BALANCE_MATCH_HIST:                                  # balance.py:84-85
            return [c.mean() / m for c, m in zip(correction, self.ref_mean)]
        return correction


class BalanceFrames:
    """Sub-action (reference stack_framework.py:235-243 protocol: ``enabled``, ``begin``, ``run_frame``, ``end``)."""

    def __init__(self, enabled=True, **kwargs):
        self.enabled = enabled
        self.process = None
        self.shape = None
        self.corr_map = kwargs.get("corr_map", BALANCE_LINEAR)
        if self.corr_map not in (BALANCE_LINEAR, BALANCE_GAMMA, BALANCE_MATCH_HIST):
            raise InvalidOptionError("corr_map", self.corr_map)
        self.channel = kwargs.get("channel", BALANCE_LUMI)
        if self.channel not in (BALANCE_LUMI, BALANCE_RGB, BALANCE_HSV, BALANCE_HLS):
            raise InvalidOptionError("channel", self.channel)
        sub = kwargs.get("subsample", DEFAULT_BALANCE_SUBSAMPLE)
        # balance.py:376-378: histogram matching looks at every pixel unless told otherwise
        self.subsample = (1 if self.corr_map == BALANCE_MATCH_HIST else DEFAULT_BALANCE_SUBSAMPLE) if sub == -1 else sub
        self.fast_subsampling = kwargs.get("fast_subsampling", False)
        self.mask_size = kwargs.get("mask_size", 0)
        self.intensity_interval = kwargs.get("intensity_interval", None)
        self.plot_histograms = kwargs.get("plot_histograms", False)   # accepted, not produced
        self.plot_summary = kwargs.get("plot_summary", False)
        self.device = int(kwargs.get("device", 0))
        self.n_channels = {BALANCE_LUMI: 1, BALANCE_RGB: 3}.get(self.channel, 2)
        self.tables = None
        self.corrections = None
        self.last_kernel_ms = None

    # ---- colour handling of the four channel modes (balance.py:232-363) ---------------------------------
    def _pre(self, image):
        import cv2  # pylint: disable=import-outside-toplevel
        if self.channel == BALANCE_HSV:
            return cv2.cvtColor(image, cv2.COLOR_BGR2HSV)
        if self.channel == BALANCE_HLS:
            return cv2.cvtColor(image, cv2.COLOR_BGR2HLS)
        return image

    def _post(self, image):
        import cv2  # pylint: disable=import-outside-toplevel
        if self.channel == BALANCE_HSV:
            return cv2.cvtColor(image, cv2.COLOR_HSV2BGR)
        if self.channel == BALANCE_HLS:
            return cv2.cvtColor(image, cv2.COLOR_HLS2BGR)
        return image

    def _histograms(self, image):
        import cv2  # pylint: disable=import-outside-toplevel
        nv = 256 if image.dtype == np.uint8 else 65536
        args = (nv, self.subsample, self.fast_subsampling, self.mask_size)
        if self.channel == BALANCE_LUMI:
            return [channel_histogram(cv2.cvtColor(image, cv2.COLOR_BGR2GRAY), *args)]
        planes = [channel_histogram(image[..., c], *args) for c in range(3)]
        return planes if self.channel == BALANCE_RGB else planes[1:]

    # ---- sub-action protocol -----------------------------------------------------------------------------
    def begin_with_reference(self, ref_image, size):
        """``Correction.begin`` (balance.py:142-158) with the reference frame given as an array."""
        self.shape = ref_image.shape
        self.tables = _Tables(self.corr_map, ref_image.dtype, self._histograms(self._pre(ref_image)),
                              self.intensity_interval)
        self.corrections = np.ones((size, self.n_channels))

    def begin(self, process):
        self.process = process
        img = read_img(process.input_full_path + "/" + process.filenames[process.ref_idx])
        self.begin_with_reference(img, process.counts)

    def look_up_tables(self, image, idx=None):
        """The (1 | 3, num_values) tables for ``image`` (already in the working colour space)."""
        correction = self.tables.solve(self._histograms(image))
        if idx is not None:
            self.corrections[idx] = self.tables.size(correction)
        if self.n_channels == 1:
            return np.stack([self.tables.table(correction[0], 0)])
        if self.n_channels == 3:
            return np.stack([self.tables.table(correction[c], c) for c in range(3)])
        identity = self.tables.values.astype(self.tables.dtype)              # the first plane (hue) is left alone
        return np.stack([identity] + [self.tables.table(correction[c], c) for c in range(2)])

    def apply_correction(self, idx, image):
        work = self._pre(image)
        luts = self.look_up_tables(work, idx)
        frame = subactions.frame_for(work, self.device)
        out = frame.upload(work).apply_lut(luts).download()
        self.last_kernel_ms = frame.kernel_ms
        return self._post(out)

    def run_frame(self, idx, _ref_idx, image):
        if self.process is None or idx != self.process.ref_idx:
            if self.process is not None:
                self.process.sub_message_r(': balance image')
            image = self.apply_correction(idx, image)
        return image

    def end(self):
        if self.process is not None:
            self.process.print_message(' ' * 60)
