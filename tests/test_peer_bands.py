"""Row bands of the fused peer exchange (ss_peer_connect, csrc/ss_lib.cu ss_band_rows), checked on the CPU: the owned
bands partition every level, and collapsing a pyramid band by band -- every rank seeing only the rows it merges --
reproduces the full collapse (reference pyramid.py:57-64) on the rows it owns."""
import ctypes

import numpy as np
import pytest

from shinestacker_b200 import _native


def bands(h, w, world, rank, min_size=32):
    lib = _native.load()
    out = (ctypes.c_int32 * (4 * 16))()
    lh = (ctypes.c_int32 * 16)()
    L = lib.ss_test_peer_bands(h, w, min_size, world, rank, out, lh)
    assert L >= 0
    a = np.array(out[:4 * (L + 2)]).reshape(4, L + 2)
    return L, a[0], a[1], a[2], a[3], list(lh[:L + 1])


@pytest.mark.parametrize("h,w", [(3072, 4096), (4000, 6000), (1300, 2000), (515, 750), (300, 407), (70, 66)])
@pytest.mark.parametrize("world", [1, 2, 3, 8])
def test_owned_bands_partition_every_level(h, w, world):
    per_rank = [bands(h, w, world, r) for r in range(world)]
    L, lh = per_rank[0][0], per_rank[0][5]
    for s in range(L + 2):
        hs = lh[min(s, L)]
        edges = [(b[3][s], b[4][s]) for b in per_rank]
        assert edges[0][0] == 0 and edges[-1][1] == hs
        for (a0, a1), (b0, b1) in zip(edges, edges[1:]):
            assert a1 == b0 and a0 <= a1
        for b in per_rank:                                   # a rank merges at least what it owns
            if b[4][s] > b[3][s]:
                assert b[1][s] <= b[3][s] and b[2][s] >= b[4][s]
            assert 0 <= b[1][s] <= b[2][s] <= hs


@pytest.mark.parametrize("h,w,world", [(515, 750, 8), (300, 407, 3), (1300, 400, 5), (260, 390, 2)])
def test_banded_collapse_equals_full_collapse(oracle, h, w, world):
    rng = np.random.default_rng(h + world)
    L, _, _, _, _, lh = bands(h, w, world, 0)
    k = oracle.make_kernel(0.4)
    sizes = [(h, w)]
    for _ in range(L):
        sizes.append(((sizes[-1][0] + 1) // 2, (sizes[-1][1] + 1) // 2))
    assert [s[0] for s in sizes] == lh
    pyr = [rng.standard_normal(s).astype(np.float32) * 20 for s in sizes]      # one channel is enough

    def collapse(levels):
        img = levels[-1]
        for l in range(L - 1, -1, -1):
            hh, ww = sizes[l]
            img = oracle.expand_plane(img, k, hh, ww) + levels[l]
        return img

    full = collapse(pyr)
    for r in range(world):
        _, row0, row1, own0, own1, _ = bands(h, w, world, r)
        seen = []
        for l in range(L):                                   # the rank knows only the rows it merged
            m = np.full(sizes[l], np.nan, np.float32)
            m[row0[l]:row1[l]] = pyr[l][row0[l]:row1[l]]
            seen.append(m)
        seen.append(pyr[L])                                  # base planes are merged whole
        # emulate the banded chain: at level l only rows [row0, row1) are produced from the coarser image
        img = seen[L]
        for l in range(L - 1, -1, -1):
            hh, ww = sizes[l]
            nxt = np.full((hh, ww), np.nan, np.float32)
            if row1[l] > row0[l]:
                e = oracle.expand_plane(img, k, hh, ww)
                nxt[row0[l]:row1[l]] = e[row0[l]:row1[l]] + seen[l][row0[l]:row1[l]]
            img = nxt
        got = img[own0[0]:own1[0]]
        assert not np.isnan(got).any()
        assert np.array_equal(got, full[own0[0]:own1[0]])
