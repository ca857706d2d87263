"""Exception types raised by the B200 stack algorithms.

When the reference package is importable its own classes are re-exported, so ``except`` clauses of a
shinestacker application keep working.  Otherwise stand-ins are built here from a table: each has the name,
the constructor arguments (kept as attributes) and the message text of its counterpart in the reference
(src/shinestacker/core/exceptions.py:2-52), which is what callers and log files see.
"""


class _Fallback(Exception):
    """Root of the stand-in hierarchy (reference: FocusStackError)."""


def _with_details(prefix):
    def text(subject, details=""):
        return f"{prefix} {subject}" + (f": {details}" if details != "" else "")
    return text


def _shape_text(shape_ref, shape):
    return (f"\nImage has shape ({shape[1]}x{shape[0]}), while it was expected "
            f"({shape_ref[1]}x{shape_ref[0]}).\n")


def _stand_in(name, arg_names, text, base):
    """An exception class whose positional arguments are ``arg_names`` (stored under those names) and whose
    message is ``text(*args)``."""
    def __init__(self, *args, **kwargs):
        bound = dict(zip(arg_names, args))
        bound.update(kwargs)
        for key, value in bound.items():
            setattr(self, key, value)
        base.__init__(self, text(**bound))
    return type(name, (base,), {"__init__": __init__, "__doc__": f"stand-in for shinestacker.core.exceptions.{name}"})


def _invalid_option_text(option, value, details=""):
    return _with_details("Invalid option")(f"{option} = {value}", details)


def _load_text(path, details=""):
    return _with_details("Failed to load")(path, details)


def _alignment_text(index, details):
    return f"Alignment failed for image {index}: {details}"


def _bit_depth_text(dtype_ref, dtype):
    return f"Image has type {dtype}, expected {dtype_ref}."


def _run_stop_text(name):
    return "Job " + (f"{name} " if name != "" else "") + "stopped"


_TABLE = {
    "InvalidOptionError": (("option", "value", "details"), _invalid_option_text),
    "ImageLoadError": (("path", "details"), _load_text),
    "AlignmentError": (("index", "details"), _alignment_text),
    "BitDepthError": (("dtype_ref", "dtype"), _bit_depth_text),
    "ShapeError": (("shape_ref", "shape"), _shape_text),
    "RunStopException": (("name",), _run_stop_text),
}

try:  # pragma: no cover - depends on the environment
    import shinestacker.core.exceptions as _ref  # type: ignore
    FocusStackError = _ref.FocusStackError
    _made = {n: getattr(_ref, n) for n in _TABLE}
except Exception:  # pylint: disable=broad-except
    FocusStackError = _Fallback
    FocusStackError.__name__ = "FocusStackError"
    _made = {n: _stand_in(n, args, text, FocusStackError) for n, (args, text) in _TABLE.items()}

InvalidOptionError = _made["InvalidOptionError"]
ImageLoadError = _made["ImageLoadError"]
AlignmentError = _made["AlignmentError"]
BitDepthError = _made["BitDepthError"]
ShapeError = _made["ShapeError"]
RunStopException = _made["RunStopException"]


class NativeLibraryError(FocusStackError):
    """libshinestack_b200.so is missing, cannot be loaded, or reported a CUDA failure.  There is no
    CPU fallback: the product path fails loudly instead."""
