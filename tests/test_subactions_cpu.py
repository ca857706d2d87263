"""BalanceFrames / Vignetting without a GPU: the host-side fits of
shinestacker_b200/balance.py and vignetting.py (tables, sigmoid parameters) and the numpy oracle of the two elementwise
passes (oracle/subaction_oracle.py) against golden vectors made by the unmodified reference, and against the live
reference where its tree is present."""
import numpy as np
import pytest

from oracle import subaction_oracle as so
from shinestacker_b200 import balance as bal, vignetting as vig, InvalidOptionError
from tools import ref_loader

import subaction_util as su

cv2 = pytest.importorskip("cv2")


def _host_balance(frames, kwargs):
    """Our fit + the oracle's table look-up (the GPU pass is checked against the same oracle on the GPU box)."""
    act = bal.BalanceFrames(**kwargs)
    act.begin_with_reference(frames[0], len(frames))
    outs = []
    for i, f in enumerate(frames[1:], 1):
        work = act._pre(f)
        outs.append(act._post(so.apply_lut(work, act.look_up_tables(work, i))))
    return np.array(outs), act.corrections


@pytest.mark.parametrize("name", su.cases("bal_"))
def test_balance_tables_reproduce_the_reference(name):
    d, kwargs, frames = su.load(name)
    outs, corr = _host_balance(frames, kwargs)
    assert np.array_equal(outs, d["out"])
    assert np.allclose(corr, d["corrections"], rtol=1e-12, atol=0)


@pytest.mark.parametrize("name", su.cases("vig_"))
def test_vignetting_fit_and_oracle_reproduce_the_reference(name):
    d, kwargs, frames = su.load(name)
    act = vig.Vignetting(**kwargs)
    for i, f in enumerate(frames):
        p = vig.fit_parameters(f, act.r_steps, act.subsample, act.fast_subsampling)
        assert np.allclose(p, d["params"][i], rtol=1e-9, atol=0)
    p0 = d["params"][0]
    out = so.correct_vignette(frames[0], p0, None, act.max_correction, act.black_threshold)
    assert np.array_equal(out, d["out"][0])


def test_radial_means_equal_the_ring_by_ring_definition():
    rng = np.random.default_rng(3)
    g = rng.integers(0, 256, (61, 87)).astype(np.uint8)
    radii, means = vig.radial_mean_intensity(g, 40)
    h, w = g.shape
    r_max = np.sqrt((w / 2) ** 2 + (h / 2) ** 2)
    edges = np.linspace(0, r_max, 41)
    y, x = np.ogrid[:h, :w]
    dist = np.sqrt((x - w / 2) ** 2 + (y - h / 2) ** 2)
    for i in range(40):                                   # reference vignetting.py:34-40
        m = (dist >= edges[i]) & (dist < edges[i + 1])
        expect = np.mean(g[m]) if m.any() else np.nan
        assert (np.isnan(expect) and np.isnan(means[i])) or expect == means[i]
    assert np.array_equal(radii, (edges[1:] + edges[:-1]) / 2)


def test_option_errors_match_the_reference():
    with pytest.raises(InvalidOptionError):
        bal.BalanceFrames(corr_map="CUBIC")
    with pytest.raises(InvalidOptionError):
        bal.BalanceFrames(channel="LAB")
    assert bal.BalanceFrames(corr_map="MATCH_HIST", subsample=-1).subsample == 1      # balance.py:376-378
    assert bal.BalanceFrames(subsample=-1).subsample == 8


@pytest.mark.skipif(ref_loader.reference_root() is None, reason="reference tree not present on this box")
@pytest.mark.parametrize("kwargs,dtype", [
    (dict(corr_map="LINEAR", channel="RGB", subsample=3), np.uint8),
    (dict(corr_map="GAMMA", channel="LUMI", mask_size=0.6), np.uint16),
    (dict(corr_map="MATCH_HIST", channel="RGB"), np.uint8),
])
def test_balance_against_the_live_reference(tmp_path, kwargs, dtype):
    ref_loader.load()
    from shinestacker.algorithms.balance import BalanceFrames as RefBalance
    frames = su.make_frames(3, 72, 110, dtype, 77)
    names = []
    for i, f in enumerate(frames):
        names.append(f"f{i}.tif")
        cv2.imwrite(str(tmp_path / names[-1]), f, [int(cv2.IMWRITE_TIFF_COMPRESSION), 1])
    proc = su.Proc(str(tmp_path), names)
    ref = RefBalance(**kwargs)
    ref.begin(proc)
    expect = np.array([ref.run_frame(i, 0, f.copy()) for i, f in enumerate(frames)][1:])
    outs, corr = _host_balance(frames, kwargs)
    assert np.array_equal(outs, expect)
    assert np.allclose(corr, ref.correction.corrections, rtol=1e-12, atol=0)


@pytest.mark.skipif(ref_loader.reference_root() is None, reason="reference tree not present on this box")
def test_vignette_oracle_against_the_live_reference():
    ref_loader.load()
    from shinestacker.algorithms import vignetting as rv
    for dtype, kw in ((np.uint8, {}), (np.uint16, dict(max_correction=0.5, black_threshold=3.0))):
        f = su.make_frames(1, 150, 201, dtype, 78)[0]
        p = rv.compute_fit_parameters(f, 100)
        assert np.allclose(vig.fit_parameters(f), p, rtol=1e-9, atol=0)
        expect = rv.correct_vignetting(f, params=p, **kw)
        assert np.array_equal(so.correct_vignette(f, p, None, kw.get("max_correction", 1), kw.get("black_threshold", 1.0)),
                              expect)
