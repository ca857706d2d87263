"""CPU-only checks of the drop-in boundary: the C-ABI library loads and exports every symbol that
include/shinestack_b200.h declares, the host mirror keeps the reference's constructor / error
contract, and the product path fails loudly (never falls back) without a GPU."""
import ctypes
import os
import re

import numpy as np
import pytest

from shinestacker_b200 import _native, pyramid
from shinestacker_b200 import PyramidStack, InvalidOptionError, NativeLibraryError

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols(header="shinestack_b200.h", prefix="ss_"):
    text = open(os.path.join(ROOT, "include", header)).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(" + prefix + r"[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = _native.load()
    names = _declared_symbols()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f"{n} declared in the header but not exported"
    # and the ctypes prototypes cover the header
    assert set(names) == set(_native.SIGNATURES), set(names) ^ set(_native.SIGNATURES)


def test_library_exports_every_depthmap_symbol():
    lib = _native.load()
    names = _declared_symbols("shinestack_b200_depthmap.h", "ssdm_")
    assert len(names) >= 14
    for n in names:
        assert hasattr(lib, n), f"{n} declared in the header but not exported"
    assert set(names) == set(_native.SIGNATURES_DEPTHMAP), set(names) ^ set(_native.SIGNATURES_DEPTHMAP)
    assert ctypes.sizeof(_native.SsdmConfig) == 72
    # option errors come back as status codes + message, never as a crash
    cfg = _native.SsdmConfig(width=64, height=64, bytes_per_sample=1, kernel_size=5, blur_size=5, smooth_size=15,
                             temperature=0.1, levels=3, float_bits=64, device=0, max_frames=2)
    h = ctypes.c_void_p()
    assert lib.ssdm_create(ctypes.byref(cfg), ctypes.byref(h)) == _native.SS_ERR_INVALID_OPTION
    assert b"float" in lib.ss_last_error()


def test_library_exports_every_subaction_symbol():
    lib = _native.load()
    names = _declared_symbols("shinestack_b200_subactions.h", "sssub_")
    assert len(names) >= 9
    for n in names:
        assert hasattr(lib, n), f"{n} declared in the header but not exported"
    assert set(names) == set(_native.SIGNATURES_SUBACTIONS), set(names) ^ set(_native.SIGNATURES_SUBACTIONS)
    h = ctypes.c_void_p()
    if lib.ss_device_count() == 0:      # no silent CPU fallback for the sub-action passes either
        assert lib.sssub_create(64, 64, 1, 0, ctypes.byref(h)) == _native.SS_ERR_NO_DEVICE
    assert lib.sssub_create(64, 64, 3, 0, ctypes.byref(h)) == _native.SS_ERR_BIT_DEPTH


def test_library_exports_every_align_symbol():
    lib = _native.load()
    names = _declared_symbols("shinestack_b200_align.h", "ssal_")
    assert len(names) >= 7
    for n in names:
        assert hasattr(lib, n), f"{n} declared in the header but not exported"
    assert set(names) == set(_native.SIGNATURES_ALIGN), set(names) ^ set(_native.SIGNATURES_ALIGN)
    h = ctypes.c_void_p()
    if lib.ss_device_count() == 0:      # no silent CPU fallback for the warp either
        assert lib.ssal_create(64, 64, 1, 0, ctypes.byref(h)) == _native.SS_ERR_NO_DEVICE


def test_version_and_config_layout():
    lib = _native.load()
    assert b"sm_100a" in lib.ss_version()
    assert ctypes.sizeof(_native.SsConfig) == 72


def test_float_type_contract():
    # reference base_stack_algo.py:13-22
    with pytest.raises(InvalidOptionError):
        PyramidStack(float_type="float-16")
    assert PyramidStack(float_type="float-64").float_type is np.float64
    a = PyramidStack()
    assert a.name() == "pyramid" and a.steps_per_frame() == 2
    assert a.float_type is np.float32 and a.do_step_callback is False
    assert a.gen_kernel.shape == (5, 5) and a.gen_kernel.dtype == np.float64
    assert a.pad_amount == 2 and list(a.offset) == [-2, -1, 0, 1, 2]


def test_invalid_options_reported_by_native_layer():
    lib = _native.load()
    cfg = _native.SsConfig(width=64, height=64, bytes_per_sample=1, min_size=32, kernel_size=5,
                           gen_kernel=0.4, float_bits=16, device=0, max_frames=2)
    h = ctypes.c_void_p()
    assert lib.ss_create(ctypes.byref(cfg), ctypes.byref(h)) == _native.SS_ERR_INVALID_OPTION
    assert b"float" in lib.ss_last_error()
    cfg.float_bits, cfg.bytes_per_sample = 32, 3
    assert lib.ss_create(ctypes.byref(cfg), ctypes.byref(h)) == _native.SS_ERR_BIT_DEPTH


def test_no_silent_cpu_fallback():
    lib = _native.load()
    if lib.ss_device_count() > 0:
        pytest.skip("a B200 is present")
    with pytest.raises(NativeLibraryError):
        PyramidStack().focus_stack_arrays([np.zeros((64, 64, 3), np.uint8)])


def test_log_prob_table_matches_reference_expression(oracle):
    rng = np.random.default_rng(3)
    gray = rng.integers(0, 256, size=(3, 48, 64)).astype(np.int32)
    counts = np.stack([np.bincount(g.ravel(), minlength=256) for g in gray])
    tab = pyramid.log_prob_table(counts)
    for i in range(3):
        assert np.array_equal(tab[i], oracle.log_prob_table(gray[i], 256), equal_nan=True)
        # literal reference expression (pyramid.py:67-69)
        levels, cnt = np.unique(gray[i].astype(np.uint8), return_counts=True)
        prob = np.zeros(256, dtype=np.float32)
        prob[levels] = cnt.astype(np.float32) / cnt.sum()
        with np.errstate(divide="ignore"):
            assert np.array_equal(tab[i], np.log(prob))


def test_log_prob_per_pixel_matches_table(oracle):
    rng = np.random.default_rng(4)
    gray = rng.integers(0, 65536, size=(2, 20, 30)).astype(np.int32)
    gray[0, :5, :5] = 777
    px = pyramid.log_prob_per_pixel(gray)
    for f in range(2):
        tab = oracle.log_prob_table(gray[f], 65536)
        assert np.array_equal(px[f], tab[gray[f]])


def test_read_img_contract(tmp_path):
    import cv2
    with pytest.raises(RuntimeError):
        pyramid.read_img(str(tmp_path / "missing.tif"))
    img = (np.arange(40 * 50 * 3) % 65536).astype(np.uint16).reshape(40, 50, 3)
    p = str(tmp_path / "a.tif")
    cv2.imwrite(p, img)
    assert np.array_equal(pyramid.read_img(p), img)
    q = str(tmp_path / "a.bmp")
    cv2.imwrite(q, (img >> 8).astype(np.uint8))
    assert pyramid.read_img(q) is None      # unsupported extension -> None, like the reference


def test_stand_in_exceptions_carry_the_reference_messages():
    # reference src/shinestacker/core/exceptions.py:2-52: same names, constructor arguments, attributes and text
    import importlib.util
    from shinestacker_b200 import exceptions as ex
    path = "/root/reference/src/shinestacker/core/exceptions.py"
    if not os.path.isfile(path):
        pytest.skip("reference tree not present on this box")
    spec = importlib.util.spec_from_file_location("ref_exceptions", path)
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    cases = [("InvalidOptionError", ("float_type", "x"), {}), ("InvalidOptionError", ("a", 3), {"details": " valid"}),
             ("ImageLoadError", ("/p/a.tif",), {}), ("ImageLoadError", ("/p", "bad"), {}),
             ("AlignmentError", (3, "too few"), {}), ("BitDepthError", (np.uint8, np.uint16), {}),
             ("ShapeError", ((10, 20), (11, 20)), {}), ("RunStopException", ("job",), {}), ("RunStopException", ("",), {})]
    for name, args, kwargs in cases:
        mine, theirs = getattr(ex, name)(*args, **kwargs), getattr(ref, name)(*args, **kwargs)
        assert str(mine) == str(theirs) and isinstance(mine, ex.FocusStackError)
    e = ex.InvalidOptionError("o", "v", details="d")
    assert (e.option, e.value, e.details) == ("o", "v", "d") and ex.ImageLoadError("p").path == "p"
