"""shinestacker_b200 -- B200-native (sm_100a CUDA) focus-stacking core.

Drop-in ``stack_algo`` for lucalista/shinestacker's ``FocusStack`` / ``FocusStackBunch``
(reference src/shinestacker/algorithms/stack.py:14-113):

    from shinestacker import StackJob, FocusStack
    from shinestacker_b200 import PyramidStack          # instead of shinestacker's PyramidStack
    job.add_action(FocusStack("stack", PyramidStack(), ...))

``DepthMapStack`` (reference src/shinestacker/algorithms/depth_map.py) is available behind the same boundary;
``AlignFrames`` / ``align_images`` (reference src/shinestacker/algorithms/align.py) warp the frames on the GPU;
``BalanceFrames`` / ``Vignetting`` (reference balance.py, vignetting.py) apply their per-frame correction on the GPU.
"""
from .pyramid import PyramidStack, PyramidBase  # noqa: F401
from .depth_map import DepthMapStack  # noqa: F401
from .align import AlignFrames, align_images, AlignmentError  # noqa: F401
from .balance import BalanceFrames  # noqa: F401
from .vignetting import Vignetting  # noqa: F401
from .exceptions import (FocusStackError, InvalidOptionError, ImageLoadError, BitDepthError,  # noqa: F401
                         ShapeError, RunStopException, NativeLibraryError)

__all__ = ["PyramidStack", "PyramidBase", "DepthMapStack", "AlignFrames", "align_images", "AlignmentError", "BalanceFrames", "Vignetting", "FocusStackError", "InvalidOptionError", "ImageLoadError",
           "BitDepthError", "ShapeError", "RunStopException", "NativeLibraryError"]
