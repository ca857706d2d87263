"""Exception types raised by the B200 stack algorithm.

When the reference package is importable its own classes are used (so existing ``except`` clauses in
a shinestacker application keep working); otherwise equivalents with the same names, constructor
arguments and messages (reference src/shinestacker/core/exceptions.py:2-52) are defined here.
"""
try:  # pragma: no cover - depends on the environment
    from shinestacker.core.exceptions import (  # type: ignore # noqa: F401
        FocusStackError, InvalidOptionError, ImageLoadError, BitDepthError, ShapeError,
        RunStopException)
except Exception:  # pylint: disable=broad-except
    class FocusStackError(Exception):
        pass

    class InvalidOptionError(FocusStackError):
        def __init__(self, option, value, details=""):
            self.option = option
            self.value = value
            self.details = details
            super().__init__(f"Invalid option {option} = {value}" +
                             ("" if details == "" else f": {details}"))

    class ImageLoadError(FocusStackError):
        def __init__(self, path, details=""):
            self.path = path
            self.details = details
            super().__init__(f"Failed to load {path}" + ("" if details == "" else f": {details}"))

    class BitDepthError(FocusStackError):
        def __init__(self, dtype_ref, dtype):
            super().__init__(f"Image has type {dtype}, expected {dtype_ref}.")

    class ShapeError(FocusStackError):
        def __init__(self, shape_ref, shape):
            super().__init__(f'''
Image has shape ({shape[1]}x{shape[0]}), while it was expected ({shape_ref[1]}x{shape_ref[0]}).
''')

    class RunStopException(FocusStackError):
        def __init__(self, name):
            if name != "":
                name = f"{name} "
            super().__init__(f"Job {name}stopped")


class NativeLibraryError(FocusStackError):
    """libshinestack_b200.so is missing, cannot be loaded, or reported a CUDA failure.  There is no
    CPU fallback: the product path fails loudly instead."""
