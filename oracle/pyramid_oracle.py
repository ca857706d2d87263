"""CPU oracle for the Laplacian-pyramid focus stacker -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import this module; nothing under ``shinestacker_b200/`` does.

It restates ``PyramidStack.focus_stack`` of lucalista/shinestacker
(``src/shinestacker/algorithms/pyramid.py:150-179``) as a *streaming* computation (one frame resident
at a time, online strict-``>`` arg-max -- identical to ``np.argmax`` first-maximum semantics,
pyramid.py:51) on top of the C primitives in ``pyramid_oracle.c``.  The arithmetic order is the one
pinned in SURVEY.md §9 for cv2 4.13.0 / numpy 2.3.5; it is re-checked against the live reference by
``tests/test_oracle_vs_reference.py`` (when ``/root/reference`` is present) and against the golden
vectors in ``tests/golden/`` (produced by the live reference with ``tests/golden/make_golden.py``).

Parity status: pinned by live differential runs against the reference in the build container and by
the committed golden vectors; the reference's own tests hold no golden vectors for this path
(tests/test_0060_stack.py only asserts "no exception").
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

_f32p = ctypes.POINTER(ctypes.c_float)
_i32p = ctypes.POINTER(ctypes.c_int32)


def build(force=False):
    """Compile pyramid_oracle.c with the committed Makefile (gcc only)."""
    so = os.path.join(_HERE, "libpyramid_oracle.so")
    src = os.path.join(_HERE, "pyramid_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-B", "libpyramid_oracle.so"], check=True,
                       stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB  # pylint: disable=global-statement
    if _LIB is None:
        so = os.path.join(_HERE, "libpyramid_oracle.so")
        if not os.path.exists(so):
            build()
        _LIB = ctypes.CDLL(so)
        _LIB.so_set_threads.restype = ctypes.c_int
    return _LIB


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def set_threads(n):
    return lib().so_set_threads(int(n))


def make_kernel(gen_kernel=0.4):
    """float32 5x5 Burt-Adelson kernel as cv2.filter2D sees it (pyramid.py:19-22)."""
    k = np.zeros(25, dtype=np.float32)
    lib().so_make_kernel(ctypes.c_double(gen_kernel), _p(k))
    return k


def filter5x5(img, k25):
    """PyramidBase.convolve (pyramid.py:24-25) on one float32 plane."""
    img = np.ascontiguousarray(img, dtype=np.float32)
    out = np.empty_like(img)
    lib().so_filter5x5(_p(img), img.shape[0], img.shape[1], _p(k25), _p(out))
    return out


def reduce_plane(img, k25):
    """PyramidBase.reduce_layer (pyramid.py:27-32) on one plane."""
    img = np.ascontiguousarray(img, dtype=np.float32)
    h, w = img.shape
    out = np.empty(((h + 1) // 2, (w + 1) // 2), dtype=np.float32)
    lib().so_reduce(_p(img), h, w, _p(k25), _p(out))
    return out


def expand_plane(img, k25, out_h=None, out_w=None, literal=False):
    """PyramidBase.expand_layer (pyramid.py:34-41) on one plane, optionally cropped."""
    img = np.ascontiguousarray(img, dtype=np.float32)
    h, w = img.shape
    if literal:
        out = np.empty((2 * h, 2 * w), dtype=np.float32)
        lib().so_expand_literal(_p(img), h, w, _p(k25), _p(out))
        if out_h is not None:
            out = np.ascontiguousarray(out[:out_h, :out_w])
        return out
    out_h = 2 * h if out_h is None else out_h
    out_w = 2 * w if out_w is None else out_w
    out = np.empty((out_h, out_w), dtype=np.float32)
    lib().so_expand(_p(img), h, w, _p(k25), _p(out), out_h, out_w)
    return out


def split_planes(img):
    """img.astype(float32) (pyramid.py:126) as three B, G, R planes."""
    img = np.ascontiguousarray(img)
    h, w, c = img.shape
    assert c == 3 and img.dtype in (np.uint8, np.uint16)
    planes = np.empty((3, h, w), dtype=np.float32)
    lib().so_split_planes(_p(img), ctypes.c_size_t(h * w), img.dtype.itemsize,
                          _p(planes[0]), _p(planes[1]), _p(planes[2]))
    return planes


def num_levels(h, w, min_size=32):
    """pyramid.py:165 -- levels = int(log2(min(h, w) / min_size))."""
    return int(np.log2(min(h, w) / min_size))


def gaussian_pyramid(planes, levels, k25):
    """First loop of process_single_image (pyramid.py:126-131): list of (3,h,w) float32 arrays."""
    pyr = [planes]
    for _ in range(levels):
        cur = pyr[-1]
        nxt = np.stack([reduce_plane(cur[c], k25) for c in range(3)])
        if min(nxt.shape[1:]) < 4:
            break
        pyr.append(nxt)
    return pyr


def laplacian_pyramid(img, levels, k25):
    """process_single_image (pyramid.py:125-139): [lap_0 .. lap_{L-1}, g_L], planes-first."""
    gp = gaussian_pyramid(split_planes(img), levels, k25)
    out = []
    for l in range(len(gp) - 1):
        g, gn = gp[l], gp[l + 1]
        lap = np.empty_like(g)
        for c in range(3):
            lib().so_laplacian(_p(g[c]), g.shape[1], g.shape[2], _p(gn[c]), gn.shape[1],
                               gn.shape[2], _p(k25), _p(lap[c]))
        out.append(lap)
    out.append(gp[-1])
    return out


def energy(lap, k25):
    """convolve(square(cvtColor(lap, BGR2GRAY))) (pyramid.py:49-50) for a (3,h,w) Laplacian."""
    _, h, w = lap.shape
    e = np.empty((h, w), dtype=np.float32)
    lib().so_energy(_p(lap[0]), _p(lap[1]), _p(lap[2]), h, w, _p(k25), _p(e))
    return e


def base_gray(base):
    """cvtColor(base, BGR2GRAY).astype(dtype) (pyramid.py:99-101) as int32."""
    _, h, w = base.shape
    g = np.empty((h, w), dtype=np.int32)
    lib().so_base_gray(_p(base[0]), _p(base[1]), _p(base[2]), ctypes.c_size_t(h * w), _p(g))
    return g


def log_prob_table(gray, num_values):
    """np.log(probabilities) of PyramidBase.entropy (pyramid.py:67-69), float32, numpy's own log.

    probabilities[levels] = counts.astype(float32) / counts.sum() is evaluated by numpy in float64
    (float32 array / int64 scalar) and rounded to float32 on assignment.  Levels that do not occur
    keep probability 0 (log -> -inf) but are never indexed by a window of the same image.
    """
    counts = np.bincount(gray.ravel(), minlength=num_values).astype(np.int64)
    prob = np.zeros(num_values, dtype=np.float32)
    nz = counts > 0
    prob[nz] = counts[nz].astype(np.float32) / counts.sum()
    with np.errstate(divide="ignore"):
        return np.log(prob)


def entropy_map(gray, log_p, ksize=5):
    h, w = gray.shape
    out = np.empty((h, w), dtype=np.float32)
    lib().so_entropy_map(_p(gray), h, w, int(ksize), _p(log_p), _p(out))
    return out


def deviation_map(gray, ksize=5):
    h, w = gray.shape
    out = np.empty((h, w), dtype=np.float32)
    lib().so_deviation_map(_p(gray), h, w, int(ksize), _p(out))
    return out


class StreamingStack64:
    """float_type='float-64' variant (reference base_stack_algo.py:15-16): Gaussian / Laplacian pyramids,
    base blend and collapse in float64 with cv2's float64 filter2D order (sequential multiply + add); gray,
    energy and the per-level arg-max stay float32 (pyramid.py:49); entropy / deviation are float64."""

    def __init__(self, min_size=32, kernel_size=5, gen_kernel=0.4):
        self.min_size, self.kernel_size = min_size, kernel_size
        self.k32 = make_kernel(gen_kernel)
        self.k64 = np.zeros(25, dtype=np.float64)
        lib().so_make_kernel_f64(ctypes.c_double(gen_kernel), _p(self.k64))
        self.n = 0
        self.state = None

    def _pyramid(self, img):
        L = lib()
        h, w, _ = img.shape
        planes = np.empty((3, h, w), np.float64)
        L.so_split_planes_f64(_p(np.ascontiguousarray(img)), ctypes.c_size_t(h * w), img.dtype.itemsize,
                              _p(planes[0]), _p(planes[1]), _p(planes[2]))
        gp = [planes]
        for _ in range(self.levels):
            cur = gp[-1]
            hh, ww = cur.shape[1:]
            nxt = np.empty((3, (hh + 1) // 2, (ww + 1) // 2), np.float64)
            for c in range(3):
                L.so_reduce_f64(_p(cur[c]), hh, ww, _p(self.k64), _p(nxt[c]))
            if min(nxt.shape[1:]) < 4:
                break
            gp.append(nxt)
        out = []
        for l in range(len(gp) - 1):
            g, gn = gp[l], gp[l + 1]
            lap = np.empty_like(g)
            for c in range(3):
                L.so_laplacian_f64(_p(g[c]), g.shape[1], g.shape[2], _p(gn[c]), gn.shape[1], gn.shape[2],
                                   _p(self.k64), _p(lap[c]))
            out.append(lap)
        out.append(gp[-1])
        return out

    def push(self, img):
        if self.n == 0:
            self.shape, self.dtype = img.shape, img.dtype
            self.levels = num_levels(img.shape[0], img.shape[1], self.min_size)
        pyr = self._pyramid(img)
        first = self.n == 0
        if first:
            self.state = [dict(e=np.empty(l.shape[1:], np.float32), idx=np.empty(l.shape[1:], np.int32),
                               fused=np.empty(l.shape, np.float64)) for l in pyr[:-1]]
            sh = pyr[-1].shape
            self.base = dict(ve=np.empty(sh[1:], np.float64), ie=np.empty(sh[1:], np.int32),
                             vd=np.empty(sh[1:], np.float64), id=np.empty(sh[1:], np.int32),
                             pe=np.empty(sh, np.float64), pd=np.empty(sh, np.float64))
        for lap, st in zip(pyr[:-1], self.state):
            e = energy(np.ascontiguousarray(lap.astype(np.float32)), self.k32)
            upd = np.ones(e.shape, bool) if first else e > st["e"]
            st["e"][upd] = e[upd]
            st["idx"][upd] = self.n
            st["fused"][:, upd] = 0.0 + lap[:, upd]
        base = pyr[-1]
        gray = base_gray(np.ascontiguousarray(base.astype(np.float32)))
        nvals = 256 if self.dtype == np.uint8 else 65536
        counts = np.bincount(gray.ravel(), minlength=nvals).astype(np.int64)
        prob = np.zeros(nvals, dtype=np.float64)
        nz = counts > 0
        prob[nz] = counts[nz].astype(np.float64) / counts.sum()
        with np.errstate(divide="ignore"):
            logp = np.log(prob)
        h, w = gray.shape
        ent = np.empty((h, w), np.float64)
        dev = np.empty((h, w), np.float64)
        lib().so_entropy_map_f64(_p(gray), h, w, int(self.kernel_size), _p(logp), _p(ent))
        lib().so_deviation_map_f64(_p(gray), h, w, int(self.kernel_size), _p(dev))
        b = self.base
        for val, bv, bi, bp in ((ent, b["ve"], b["ie"], b["pe"]), (dev, b["vd"], b["id"], b["pd"])):
            upd = np.ones(val.shape, bool) if first else val > bv
            bv[upd] = val[upd]
            bi[upd] = self.n
            bp[:, upd] = base[:, upd]
        self.n += 1

    def finish(self):
        L = lib()
        b = self.base
        img = ((0.0 + b["pe"]) + b["pd"]) / 2
        for st in self.state[::-1]:
            lay = st["fused"]
            nxt = np.empty_like(lay)
            for c in range(3):
                L.so_collapse_step_f64(_p(np.ascontiguousarray(img[c])), img.shape[1], img.shape[2], _p(lay[c]),
                                       lay.shape[1], lay.shape[2], _p(self.k64), _p(nxt[c]))
            img = nxt
        _, H, W = img.shape
        out = np.empty((H, W, 3), dtype=self.dtype)
        for c in range(3):
            L.so_finalize_plane_f64(_p(np.ascontiguousarray(img[c])), ctypes.c_size_t(H * W), self.dtype.itemsize, c, _p(out))
        info = dict(best=[st["idx"] for st in self.state], best_e=b["ie"], best_d=b["id"], levels=len(self.state))
        return out, info


class StreamingStack:
    """Streaming restatement of PyramidStack.focus_stack (pyramid.py:150-179).

    push(frame) must be called in frame order; finish() returns (fused_image, info) where info holds
    the integer depth maps: ``best`` (list, level 0..L-1), ``best_e``, ``best_d``.
    """

    def __init__(self, min_size=32, kernel_size=5, gen_kernel=0.4):
        self.min_size = min_size
        self.kernel_size = kernel_size
        self.k25 = make_kernel(gen_kernel)
        self.n = 0
        self.shape = None
        self.dtype = None
        self.levels = None
        self.state = None

    def push(self, img):
        L = lib()
        if self.n == 0:
            self.shape, self.dtype = img.shape, img.dtype
            self.levels = num_levels(img.shape[0], img.shape[1], self.min_size)
        assert img.shape == self.shape and img.dtype == self.dtype
        pyr = laplacian_pyramid(img, self.levels, self.k25)
        first = int(self.n == 0)
        if first:
            self.state = []
            for lap in pyr[:-1]:
                _, h, w = lap.shape
                self.state.append(dict(e=np.empty((h, w), np.float32),
                                       idx=np.empty((h, w), np.int32),
                                       fused=np.empty((3, h, w), np.float32)))
            _, h, w = pyr[-1].shape
            self.base = dict(ve=np.empty((h, w), np.float32), ie=np.empty((h, w), np.int32),
                             vd=np.empty((h, w), np.float32), id=np.empty((h, w), np.int32),
                             pe=np.empty((3, h, w), np.float32), pd=np.empty((3, h, w), np.float32))
        for lap, st in zip(pyr[:-1], self.state):
            e = energy(lap, self.k25)
            f = st["fused"]
            L.so_select_update(_p(e), _p(lap[0]), _p(lap[1]), _p(lap[2]), ctypes.c_size_t(e.size),
                               self.n, first, _p(st["e"]), _p(st["idx"]), _p(f[0]), _p(f[1]),
                               _p(f[2]))
        base = pyr[-1]
        gray = base_gray(base)
        nvals = 256 if self.dtype == np.uint8 else 65536
        ent = entropy_map(gray, log_prob_table(gray, nvals), self.kernel_size)
        dev = deviation_map(gray, self.kernel_size)
        b = self.base
        for val, bv, bi, bp in ((ent, b["ve"], b["ie"], b["pe"]), (dev, b["vd"], b["id"], b["pd"])):
            upd = np.ones(val.shape, bool) if first else val > bv
            bv[upd] = val[upd]
            bi[upd] = self.n
            bp[:, upd] = base[:, upd]
        self.n += 1

    def finish(self):
        L = lib()
        b = self.base
        _, h, w = b["pe"].shape
        img = np.empty((3, h, w), np.float32)
        for c in range(3):
            L.so_fuse_base(_p(b["pe"][c]), _p(b["pd"][c]), ctypes.c_size_t(h * w), _p(img[c]))
        fused_pyr = [st["fused"] for st in self.state] + [img.copy()]
        for st in self.state[::-1]:
            lay = st["fused"]
            _, lh, lw = lay.shape
            nxt = np.empty((3, lh, lw), np.float32)
            for c in range(3):
                L.so_collapse_step(_p(img[c]), img.shape[1], img.shape[2], _p(lay[c]), lh, lw,
                                   _p(self.k25), _p(nxt[c]))
            img = nxt
        _, H, W = img.shape
        out = np.empty((H, W, 3), dtype=self.dtype)
        for c in range(3):
            L.so_finalize_plane(_p(img[c]), ctypes.c_size_t(H * W), self.dtype.itemsize, c, _p(out))
        info = dict(best=[st["idx"] for st in self.state], best_e=b["ie"], best_d=b["id"],
                    levels=len(self.state),
                    fused_pyramid=[np.ascontiguousarray(p.transpose(1, 2, 0)) for p in fused_pyr])
        return out, info


def focus_stack(frames, min_size=32, kernel_size=5, gen_kernel=0.4, float_type='float-32'):
    """Oracle entry point: list/iterable of HxWx3 u8/u16 BGR frames -> (fused image, info)."""
    s = (StreamingStack64 if float_type == 'float-64' else StreamingStack)(min_size, kernel_size, gen_kernel)
    for f in frames:
        s.push(f)
    return s.finish()
