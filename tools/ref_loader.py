"""Loader of the UNMODIFIED reference (lucalista/shinestacker) for bench.py's reference arm and for the test that lets
the reference's own host objects drive the GPU stacker.  NOT used by the product path.

The reference is installed (pip --target, no sources copied into the history) under ``baseline/_ref`` -- git-ignored,
but shipped to the GPU box by gpurun; in the build container ``/root/reference/src`` works too.  Its non-arithmetic
third-party modules that this image lacks (matplotlib, tifffile, imagecodecs, psdtags) are stubbed, and the
setuptools-scm generated ``shinestacker._version`` is supplied; nothing on the arithmetic path is touched.
"""
import os
import sys
import types
from unittest.mock import MagicMock

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CANDIDATES = [os.environ.get("SHINESTACKER_REF", ""), os.path.join(ROOT, "baseline", "_ref"), "/root/reference/src"]


def reference_root():
    if os.environ.get("SS_NO_REFERENCE"):      # tests: behave as on a box without the reference
        return None
    for p in CANDIDATES:
        if p and os.path.isdir(os.path.join(p, "shinestacker", "algorithms")):
            return p
    return None


def load():
    """Imports the reference package; returns its root directory or None when it is not available."""
    root = reference_root()
    if root is None:
        return None
    if "shinestacker.algorithms.pyramid" in sys.modules:
        return root
    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.backends", "matplotlib.backends.backend_pdf",
                 "tifffile", "imagecodecs", "psdtags"):
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:  # pylint: disable=broad-except
                sys.modules[name] = MagicMock()
    if "shinestacker._version" not in sys.modules and \
            not os.path.isfile(os.path.join(root, "shinestacker", "_version.py")):
        ver = types.ModuleType("shinestacker._version")
        ver.__version__ = ver.version = "0.0.0-ref"
        sys.modules["shinestacker._version"] = ver
    if root not in sys.path:
        sys.path.insert(0, root)
    import shinestacker.algorithms.pyramid  # noqa: F401  pylint: disable=import-outside-toplevel,unused-import
    return root


def versions():
    import cv2  # pylint: disable=import-outside-toplevel
    import numpy  # pylint: disable=import-outside-toplevel
    ipp = "?"
    try:
        ipp = "on" if cv2.ipp.useIPP() else "off"
    except Exception:  # pylint: disable=broad-except
        pass
    return f"cv2 {cv2.__version__} (IPP {ipp}, {cv2.getNumThreads()} threads), numpy {numpy.__version__}"


def run_stack_job(work_dir, input_subdir, stack_algo, output_subdir="out", name="ref"):
    """The reference's own smoke path (tests/test_0060_stack.py:7-14): StackJob -> FocusStack(stack_algo) -> job.run();
    returns the path of the written image."""
    from shinestacker.algorithms.stack_framework import StackJob  # pylint: disable=import-outside-toplevel
    from shinestacker.algorithms.stack import FocusStack  # pylint: disable=import-outside-toplevel
    import logging  # pylint: disable=import-outside-toplevel
    # the reference's print helpers write through handlers[0] of the root and "tqdm" loggers (core/framework.py:116);
    # give them quiet ones instead of calling its setup_logging(), which creates log files inside the package tree
    for lname in (None, "tqdm"):
        lg = logging.getLogger(lname)
        if not lg.handlers:
            lg.addHandler(logging.NullHandler())
    job = StackJob(name, work_dir, input_path=input_subdir, logger_name="b200_ref_arm")
    job.add_action(FocusStack("stack", stack_algo, output_path=output_subdir, plot_stack=False))
    job.run()
    out_dir = os.path.join(work_dir, output_subdir)
    files = sorted(f for f in os.listdir(out_dir) if not f.startswith("."))
    return os.path.join(out_dir, files[0]) if files else None
