"""Live differential test of the oracle against the unmodified reference (and against cv2 itself):
runs only where /root/reference (or $SHINESTACKER_REF) and cv2 are present, i.e. in the build
container.  This is the check that pins the oracle; the golden vectors freeze its outcome."""
import numpy as np
import pytest

from oracle import ref_harness
from shinestacker_b200 import synthetic

pytestmark = pytest.mark.skipif(not ref_harness.reference_available(),
                                reason="reference tree not present on this box")


@pytest.fixture(scope="module")
def ref_algo():
    return ref_harness.make_recording_stack()


@pytest.mark.parametrize("h,w", [(1, 1), (3, 5), (4, 4), (7, 9), (16, 8), (17, 23), (33, 40), (50, 127),
                                 (64, 128), (65, 129), (20, 375), (31, 188), (40, 1500)])
def test_filter_primitives_match_cv2(oracle, ref_algo, h, w):
    """so_filter5x5 / so_reduce / so_expand vs PyramidBase.convolve / reduce_layer / expand_layer
    (reference pyramid.py:24-46), incl. widths with a scalar tail (w % 8 != 0)."""
    rng = np.random.default_rng(h * 1000 + w)
    k = oracle.make_kernel(0.4)
    img = (rng.random((h, w), dtype=np.float32) * 65535).astype(np.float32)
    assert np.array_equal(ref_algo.convolve(img), oracle.filter5x5(img, k))
    if h >= 2 and w >= 2:
        assert np.array_equal(ref_algo.reduce_layer(img), oracle.reduce_plane(img, k))
    exp = ref_algo.expand_layer(img)
    assert np.array_equal(exp, oracle.expand_plane(img, k))
    assert np.array_equal(exp, oracle.expand_plane(img, k, literal=True))


def test_numerics_self_probe_gray():
    """cv2.cvtColor(float32 BGR2GRAY) order on this box."""
    import cv2
    rng = np.random.default_rng(5)
    img = ((rng.random((37, 53, 3), dtype=np.float32) - 0.5) * 500).astype(np.float32)
    b, g, r = (img[..., i].astype(np.float64) for i in range(3))
    # fma(R,.299f, fma(B,.114f, G*.587f)) emulated in float64 with explicit float32 roundings
    t = (g.astype(np.float32) * np.float32(0.587)).astype(np.float32)
    t = (b * np.float64(np.float32(0.114)) + t.astype(np.float64)).astype(np.float32)
    t = (r * np.float64(np.float32(0.299)) + t.astype(np.float64)).astype(np.float32)
    assert np.array_equal(cv2.cvtColor(img, cv2.COLOR_BGR2GRAY), t)


CASES = [
    ("stack", 5, 300, 407, np.uint8, 21, {}),
    ("stack", 4, 260, 376, np.uint16, 22, {}),
    ("noise", 3, 131, 200, np.uint8, 23, {}),
    ("noise", 3, 140, 133, np.uint16, 24, {"gen_kernel": 0.3}),
    ("stack", 3, 100, 90, np.uint8, 25, {"min_size": 8}),
    ("stack", 4, 96, 150, np.uint8, 26, {"kernel_size": 3}),
    ("stack", 2, 40, 50, np.uint8, 27, {}),
    ("noise", 2, 70, 66, np.uint8, 28, {"kernel_size": 4}),
    ("noise", 3, 131, 200, np.uint8, 29, {"float_type": "float-64"}),
    ("stack", 4, 260, 376, np.uint16, 30, {"float_type": "float-64"}),
    ("noise", 3, 140, 133, np.uint16, 31, {"float_type": "float-64", "gen_kernel": 0.3, "kernel_size": 3}),
    # base windows above 128 values: numpy's pairwise sum halves the run (13 -> 169, 20 -> 361, 32 -> 961 values)
    ("noise", 3, 150, 200, np.uint8, 32, {"kernel_size": 13}),
    ("stack", 3, 300, 260, np.uint16, 33, {"kernel_size": 20}),
    ("noise", 2, 280, 300, np.uint8, 34, {"kernel_size": 32}),
    ("noise", 2, 150, 200, np.uint8, 35, {"kernel_size": 17, "float_type": "float-64"}),
]


@pytest.mark.parametrize("gen,n,h,w,dtype,seed,kwargs", CASES)
def test_oracle_matches_live_reference(oracle, gen, n, h, w, dtype, seed, kwargs):
    f = synthetic.make_stack if gen == "stack" else synthetic.make_noise_stack
    frames = f(n, h, w, dtype, seed)
    ref, ri = ref_harness.reference_stack_arrays(frames, **kwargs)
    out, oi = oracle.focus_stack(frames, **kwargs)
    assert ri["levels"] == oi["levels"]
    for a, b in zip(ri["best"], oi["best"]):
        assert np.array_equal(a, b)
    assert np.array_equal(ri["best_e"], oi["best_e"])
    assert np.array_equal(ri["best_d"], oi["best_d"])
    if "fused_pyramid" in oi:
        for a, b in zip(ri["fused_pyramid"], oi["fused_pyramid"]):
            assert np.array_equal(a, b)
    assert np.array_equal(ref, out)
