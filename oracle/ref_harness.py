"""Live-reference harness (TEST INFRASTRUCTURE ONLY -- never imported by the product path).

Imports the *unmodified* reference (lucalista/shinestacker) from ``/root/reference/src`` (or
``$SHINESTACKER_REF``) with the non-arithmetic third-party modules it needs stubbed out
(matplotlib, tifffile, imagecodecs, psdtags and the setuptools-scm generated ``_version``), and
wraps ``PyramidStack`` (src/shinestacker/algorithms/pyramid.py:114-179) in a recording subclass
so the per-level arg-max maps (``best`` at pyramid.py:51, ``best_e``/``best_d`` at
pyramid.py:104-105) can be captured.  The reference tree exists only in the build container; the
GPU box never sees it, so this module is used only by ``tests/golden/make_golden.py`` and by the
CPU tests that pin ``oracle/`` against the live reference when it is present.
"""
import os
import sys
import types
from unittest.mock import MagicMock

import numpy as np

_REF_CANDIDATES = [os.environ.get("SHINESTACKER_REF", ""), "/root/reference/src"]


def reference_available():
    return any(p and os.path.isdir(os.path.join(p, "shinestacker")) for p in _REF_CANDIDATES)


def import_reference():
    """Return the reference's ``shinestacker.algorithms.pyramid`` module."""
    if "shinestacker.algorithms.pyramid" in sys.modules:
        return sys.modules["shinestacker.algorithms.pyramid"]
    src = next((p for p in _REF_CANDIDATES
                if p and os.path.isdir(os.path.join(p, "shinestacker"))), None)
    if src is None:
        raise RuntimeError("reference tree not found (set SHINESTACKER_REF)")
    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.backends",
                 "matplotlib.backends.backend_pdf", "tifffile", "imagecodecs", "psdtags"):
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:  # pylint: disable=broad-except
                sys.modules[name] = MagicMock()
    if "shinestacker._version" not in sys.modules:
        ver = types.ModuleType("shinestacker._version")
        ver.__version__ = "0.0.0-ref"
        ver.version = "0.0.0-ref"
        sys.modules["shinestacker._version"] = ver
    if src not in sys.path:
        sys.path.insert(0, src)
    import shinestacker.algorithms.pyramid as pyr  # pylint: disable=import-outside-toplevel
    return pyr


def make_recording_stack(**kwargs):
    """A reference ``PyramidStack`` that records every arg-max map it computes."""
    pyr = import_reference()

    class RecordingPyramidStack(pyr.PyramidStack):
        def __init__(self, **kw):
            super().__init__(**kw)
            self.rec_best = []      # level L-1 ... 0 in call order (pyramid.py:144-147)
            self.rec_best_e = None
            self.rec_best_d = None
            self.rec_fused = None

        def fuse_laplacian(self, laplacians):
            gray_laps = [pyr.cv2.cvtColor(lap.astype(np.float32), pyr.cv2.COLOR_BGR2GRAY)
                         for lap in laplacians]
            energies = [self.convolve(np.square(g)) for g in gray_laps]
            self.rec_best.append(np.argmax(energies, axis=0))
            return super().fuse_laplacian(laplacians)

        def get_fused_base(self, images):
            layers = images.shape[0]
            gray = np.array([pyr.cv2.cvtColor(images[i].astype(np.float32),
                                              pyr.cv2.COLOR_BGR2GRAY).astype(self.dtype)
                             for i in range(layers)])
            ent = np.array([self.entropy(g) for g in gray])
            dev = np.array([self.deviation(g) for g in gray])
            self.rec_best_e = np.argmax(ent, axis=0)
            self.rec_best_d = np.argmax(dev, axis=0)
            self.rec_entropy = ent
            self.rec_deviation = dev
            return super().get_fused_base(images)

        def fuse_pyramids(self, all_laplacians):
            self.rec_fused = super().fuse_pyramids(all_laplacians)
            return self.rec_fused

    algo = RecordingPyramidStack(**kwargs)
    algo.process = MagicMock()
    algo.process.callback.return_value = True
    return algo


def reference_stack_arrays(frames, **kwargs):
    """Run the reference on in-memory frames (list of HxWx3 u8/u16 BGR arrays).

    Follows PyramidStack.focus_stack (pyramid.py:150-179) minus the file reads; returns
    ``(fused_image, dict(best=[level0..L-1], best_e, best_d, levels))``.
    """
    pyr = import_reference()
    algo = make_recording_stack(**kwargs)
    img0 = frames[0]
    algo.dtype = img0.dtype
    const = pyr.constants
    algo.num_pixel_values = const.NUM_UINT8 if img0.dtype == np.uint8 else const.NUM_UINT16
    algo.max_pixel_value = const.MAX_UINT8 if img0.dtype == np.uint8 else const.MAX_UINT16
    levels = int(np.log2(min(img0.shape[:2]) / algo.min_size))
    pyrs = [algo.process_single_image(f, levels) for f in frames]
    fused = algo.collapse(algo.fuse_pyramids(pyrs))
    out = fused.astype(algo.dtype)
    info = dict(best=algo.rec_best[::-1], best_e=algo.rec_best_e, best_d=algo.rec_best_d,
                levels=len(pyrs[0]) - 1, fused_pyramid=algo.rec_fused,
                entropy=algo.rec_entropy, deviation=algo.rec_deviation)
    return out, info


def import_reference_depthmap():
    """Return the reference's ``shinestacker.algorithms.depth_map`` module."""
    import_reference()
    import shinestacker.algorithms.depth_map as dm  # pylint: disable=import-outside-toplevel
    return dm


def reference_depthmap_arrays(frames, **kwargs):
    """Run the reference ``DepthMapStack`` (depth_map.py:10-123) on in-memory BGR frames.

    ``focus_stack`` reads every file twice through ``read_img`` (depth_map.py:70,96); the harness feeds the
    arrays through those two hooks instead of files, everything else is the unmodified reference code.
    Returns ``(fused_image, dict(energy=..., smoothed=..., weights=...))`` (stages recorded by overriding
    ``smooth_energy`` / ``get_focus_map``).
    """
    dm = import_reference_depthmap()
    frames = list(frames)
    names = [f"/virtual/{i:04d}.tif" for i in range(len(frames))]
    lookup = dict(zip(names, frames))
    rec = {}

    class RecordingDepthMapStack(dm.DepthMapStack):
        def read_image_and_update_metadata(self, img_path, metadata):
            img = lookup[img_path]
            if metadata is None:
                return img, (img.shape[:2], img.dtype), True
            return img, metadata, False

        def smooth_energy(self, energy_map):
            rec["normalized"] = np.array(energy_map)
            out = super().smooth_energy(energy_map)
            rec["smoothed"] = np.array(out)
            return out

        def get_focus_map(self, energies):
            if "smoothed" not in rec:
                rec["smoothed"] = np.array(energies)
            out = super().get_focus_map(energies)
            rec["weights"] = np.array(out)
            return out

    algo = RecordingDepthMapStack(**kwargs)
    algo.process = MagicMock()
    algo.process.callback.return_value = True
    algo.print_message = MagicMock()
    saved = dm.read_img
    dm.read_img = lambda p: lookup[p]
    try:
        out = algo.focus_stack(names)
    finally:
        dm.read_img = saved
    return out, rec
