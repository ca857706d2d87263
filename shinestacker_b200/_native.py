"""ctypes binding of libshinestack_b200.so (C ABI: include/shinestack_b200.h).

The library is built in-tree by ``__graft_entry__.build()`` / ``python -m shinestacker_b200.build``.
If it is missing or cannot be loaded the import of this module raises: there is no CPU fallback.
"""
import ctypes
import os

from .exceptions import NativeLibraryError

_HERE = os.path.dirname(os.path.abspath(__file__))
# SHINESTACK_B200_LIB selects an alternative build of the same ABI (kernel A/B experiments)
LIB_PATH = os.environ.get("SHINESTACK_B200_LIB") or os.path.join(_HERE, "libshinestack_b200.so")


class SsConfig(ctypes.Structure):
    """struct ss_config (include/shinestack_b200.h)."""
    _fields_ = [
        ("width", ctypes.c_int32),
        ("height", ctypes.c_int32),
        ("bytes_per_sample", ctypes.c_int32),
        ("min_size", ctypes.c_int32),
        ("kernel_size", ctypes.c_int32),
        ("gen_kernel", ctypes.c_double),
        ("float_bits", ctypes.c_int32),
        ("device", ctypes.c_int32),
        ("max_frames", ctypes.c_int32),
        ("batch_frames", ctypes.c_int32),
        ("frame_index_base", ctypes.c_int32),
        ("resident_frames", ctypes.c_int32),
        ("reserved", ctypes.c_int32 * 4),
    ]


class SsdmConfig(ctypes.Structure):
    """struct ssdm_config (include/shinestack_b200_depthmap.h)."""
    _fields_ = [
        ("width", ctypes.c_int32),
        ("height", ctypes.c_int32),
        ("bytes_per_sample", ctypes.c_int32),
        ("map_type", ctypes.c_int32),
        ("energy", ctypes.c_int32),
        ("kernel_size", ctypes.c_int32),
        ("blur_size", ctypes.c_int32),
        ("smooth_size", ctypes.c_int32),
        ("temperature", ctypes.c_double),
        ("levels", ctypes.c_int32),
        ("float_bits", ctypes.c_int32),
        ("device", ctypes.c_int32),
        ("max_frames", ctypes.c_int32),
        ("reserved", ctypes.c_int32 * 4),
    ]


SS_OK = 0
SS_ERR_INVALID_OPTION = -1
SS_ERR_SHAPE = -2
SS_ERR_BIT_DEPTH = -3
SS_ERR_CUDA = -4
SS_ERR_NO_DEVICE = -5
SS_ERR_STATE = -6
SS_ERR_NOMEM = -7
SS_ERR_COMM = -8

_vp = ctypes.c_void_p
_sz = ctypes.c_size_t
_i = ctypes.c_int

# name -> (restype, argtypes); every symbol declared in include/shinestack_b200.h
SIGNATURES = {
    "ss_version": (ctypes.c_char_p, []),
    "ss_last_error": (ctypes.c_char_p, []),
    "ss_device_count": (_i, []),
    "ss_create": (_i, [ctypes.POINTER(SsConfig), ctypes.POINTER(_vp)]),
    "ss_destroy": (None, [_vp]),
    "ss_reset": (_i, [_vp]),
    "ss_num_levels": (_i, [_vp]),
    "ss_level_size": (_i, [_vp, _i, ctypes.POINTER(ctypes.c_int32), ctypes.POINTER(ctypes.c_int32)]),
    "ss_host_alloc": (_i, [ctypes.POINTER(_vp), _sz]),
    "ss_host_free": (_i, [_vp]),
    "ss_host_is_pinned": (_i, [_vp]),
    "ss_push_frame": (_i, [_vp, _vp, _sz]),
    "ss_push_frame_device": (_i, [_vp, _vp, _sz]),
    "ss_push_frame_rgb": (_i, [_vp, _vp, _sz]),
    "ss_sync": (_i, [_vp]),
    "ss_sync_copies": (_i, [_vp]),
    "ss_finish_begin": (_i, [_vp, _vp]),
    "ss_finish_end": (_i, [_vp, _vp, _vp, _sz]),
    "ss_finish": (_i, [_vp, _vp, _sz]),
    "ss_get_argmax": (_i, [_vp, _i, _vp]),
    "ss_get_fused_level": (_i, [_vp, _i, _vp]),
    "ss_kernel_time_ms": (_i, [_vp, ctypes.POINTER(ctypes.c_float)]),
    "ss_set_profiling": (_i, [_vp, _i]),
    "ss_level_kernel_time_ms": (_i, [_vp, _i, ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_int32)]),
    "ss_launch_count": (_i, [_vp, ctypes.POINTER(ctypes.c_int64)]),
    "ss_rerun_resident": (_i, [_vp]),
    "ss_finish_select_base": (_i, [_vp, _vp]),
    "ss_get_base_gray": (_i, [_vp, _vp]),
    "ss_finish_select_base_f64": (_i, [_vp, _vp, _i]),
    "ss_finish_select_base_px": (_i, [_vp, _vp]),
    "ss_finish_collapse": (_i, [_vp, _vp, _sz]),
    "ss_segment_time_ms": (_i, [_vp, _i, ctypes.POINTER(ctypes.c_float)]),
    "ss_peer_export": (_i, [_vp, _vp]),
    "ss_peer_connect": (_i, [_vp, _vp, _i, _i, _i]),
    "ss_peer_ptrs": (_i, [_vp, _vp]),
    "ss_peer_connect_ptrs": (_i, [_vp, _vp, _i, _i, _i]),
    "ss_peer_finish": (_i, [_vp, _vp, _sz]),
    "ss_peer_finish_async": (_i, [_vp]),
    "ss_peer_wait": (_i, [_vp, _vp, _sz]),
    "ss_peer_phase": (_i, [_vp, _i]),
    "ss_test_peer_bands": (_i, [_i, _i, _i, _i, _i, _vp, _vp]),
    "ss_set_log_lut": (_i, [_vp, _vp, _i]),
    "ss_finish_base_device": (_i, [_vp]),
    "ss_download_output": (_i, [_vp, _vp, _sz]),
    "ss_get_best_energy": (_i, [_vp, _i, _vp]),
    "ss_get_gauss_level": (_i, [_vp, _i, _i, _vp]),
    "ss_test_collapse_step": (_i, [_vp, _i, _i, _vp, _i, _i, ctypes.c_double, _vp]),
}

# every symbol declared in include/shinestack_b200_depthmap.h
SIGNATURES_DEPTHMAP = {
    "ssdm_create": (_i, [ctypes.POINTER(SsdmConfig), ctypes.POINTER(_vp)]),
    "ssdm_destroy": (None, [_vp]),
    "ssdm_reset": (_i, [_vp]),
    "ssdm_push_frame": (_i, [_vp, _vp, _sz]),
    "ssdm_push_frame_device": (_i, [_vp, _vp, _sz]),
    "ssdm_sync_copies": (_i, [_vp]),
    "ssdm_finish_begin": (_i, [_vp]),
    "ssdm_blend_frame": (_i, [_vp, _i]),
    "ssdm_finish_end": (_i, [_vp, _vp, _sz]),
    "ssdm_finish": (_i, [_vp, _vp, _sz]),
    "ssdm_rerun_resident": (_i, [_vp]),
    "ssdm_get_plane": (_i, [_vp, _i, _i, _vp]),
    "ssdm_sync": (_i, [_vp]),
    "ssdm_shard_local_max": (_i, [_vp, ctypes.POINTER(ctypes.c_float)]),
    "ssdm_shard_finish_local": (_i, [_vp, ctypes.c_float]),
    "ssdm_shard_buffer_count": (_i, [_vp]),
    "ssdm_shard_export": (_i, [_vp, _vp]),
    "ssdm_shard_connect": (_i, [_vp, _vp, _i, _i, _i]),
    "ssdm_shard_ptrs": (_i, [_vp, ctypes.POINTER(_vp)]),
    "ssdm_shard_connect_ptrs": (_i, [_vp, ctypes.POINTER(_vp), _i, _i, _i]),
    "ssdm_shard_softmax_rel": (_i, [_vp]),
    "ssdm_shard_finish_weights": (_i, [_vp]),
    "ssdm_shard_reduce_band": (_i, [_vp]),
    "ssdm_shard_finish_end": (_i, [_vp, _vp, _sz]),
    "ssdm_kernel_time_ms": (_i, [_vp, ctypes.POINTER(ctypes.c_float)]),
    "ssdm_launch_count": (_i, [_vp, ctypes.POINTER(ctypes.c_int64)]),
}

# every symbol declared in include/shinestack_b200_align.h
_dp = ctypes.POINTER(ctypes.c_double)
_ip = ctypes.POINTER(ctypes.c_int32)
SIGNATURES_ALIGN = {
    "ssal_create": (_i, [ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, ctypes.POINTER(_vp)]),
    "ssal_destroy": (None, [_vp]),
    "ssal_warp": (_i, [_vp, _vp, _sz, _dp, _i, _i, _dp, _ip, _i, _vp, _sz]),
    "ssal_warp_device": (_i, [_vp, _vp, _sz, _dp, _i, _i, _dp, _ip, _i, _vp, _sz]),
    "ssal_ecc_set_reference": (_i, [_vp, _vp, _sz]),
    "ssal_ecc_estimate": (_i, [_vp, _vp, _sz, _i, _i, ctypes.c_double, _dp, _dp, _dp]),
    "ssal_prefetch": (_i, [_vp, _vp, _sz]),
    "ssal_use_prefetched": (_i, [_vp]),
    "ssal_sync": (_i, [_vp]),
    "ssal_result_device": (_i, [_vp, ctypes.POINTER(_vp), ctypes.POINTER(_sz)]),
    "ssal_kernel_time_ms": (_i, [_vp, ctypes.POINTER(ctypes.c_float)]),
    "ssal_launch_count": (_i, [_vp, ctypes.POINTER(ctypes.c_int64)]),
}

# every symbol declared in include/shinestack_b200_subactions.h
SIGNATURES_SUBACTIONS = {
    "sssub_create": (_i, [ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, ctypes.POINTER(_vp)]),
    "sssub_destroy": (None, [_vp]),
    "sssub_upload": (_i, [_vp, _vp, _sz]),
    "sssub_apply_lut": (_i, [_vp, _vp, _i]),
    "sssub_apply_vignette": (_i, [_vp, _dp, ctypes.c_double, ctypes.c_double, ctypes.c_double]),
    "sssub_download": (_i, [_vp, _vp, _sz]),
    "sssub_sync": (_i, [_vp]),
    "sssub_device_frame": (_i, [_vp, ctypes.POINTER(_vp), ctypes.POINTER(_sz)]),
    "sssub_kernel_time_ms": (_i, [_vp, ctypes.POINTER(ctypes.c_float)]),
    "sssub_launch_count": (_i, [_vp, ctypes.POINTER(ctypes.c_int64)]),
}

_LIB = None


def load():
    """Load the shared library (once) and declare every prototype.  Raises NativeLibraryError."""
    global _LIB  # pylint: disable=global-statement
    if _LIB is not None:
        return _LIB
    if not os.path.exists(LIB_PATH):
        raise NativeLibraryError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; "
            "g.build()'` (nvcc, sm_100a).  shinestacker_b200 has no CPU fallback.")
    try:
        lib = ctypes.CDLL(LIB_PATH)
    except OSError as e:
        raise NativeLibraryError(f"cannot load {LIB_PATH}: {e}") from e
    for name, (res, args) in list(SIGNATURES.items()) + list(SIGNATURES_DEPTHMAP.items()) + \
            list(SIGNATURES_ALIGN.items()) + list(SIGNATURES_SUBACTIONS.items()):
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _LIB = lib
    return lib


def last_error():
    return load().ss_last_error().decode("utf-8", "replace")
