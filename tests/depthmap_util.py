"""Shared helpers of the depth-map tests: golden loader (tests/golden/depthmap, see
tests/golden/make_golden_depthmap.py), the ctypes wrapper of the CPU emulation of the kernels
(tests/emu/emu_depthmap.cpp) and the 1-LSB comparison that is the parity contract of DepthMapStack
(soft weights -> tolerance parity, <= 1 LSB)."""
import ast
import ctypes
import glob
import os
import subprocess

import numpy as np

from shinestacker_b200 import synthetic

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden", "depthmap")


def case_names():
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "*.npz")))


def load_case(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    kwargs = ast.literal_eval(str(z["kwargs"]))
    if "frames" in z.files:
        frames = [np.ascontiguousarray(f) for f in z["frames"]]
    else:
        gen, n, h, w, dtype, seed = ast.literal_eval(str(z["recipe"]))
        f = synthetic.make_stack if gen == "stack" else synthetic.make_noise_stack
        frames = f(n, h, w, np.dtype(dtype), seed)
    return frames, kwargs, dict(fused=z["fused"], weights0=z["weights0"].astype(np.float32))


def assert_within_1lsb(out, ref, what, max_frac=None):
    """<= 1 LSB everywhere; and the 1-LSB differences (truncation of values that sit on an integer
    boundary) must stay rare: float32 carries 24 bits, so for 16-bit samples a last-ulp difference in the
    blend flips the truncated integer for a few percent of the samples, for 8-bit samples for ~1e-4."""
    assert out.dtype == ref.dtype and out.shape == ref.shape, what
    if max_frac is None:
        max_frac = 0.002 if out.dtype == np.uint8 else 0.08
    diff = np.abs(out.astype(np.int64) - ref.astype(np.int64))
    assert diff.max() <= 1, f"{what}: max difference {diff.max()} LSB"
    frac = float((diff > 0).mean())
    assert frac <= max_frac, f"{what}: {frac:.4f} of the samples differ by 1 LSB"
    return frac


class Emu:
    """The product's pixel functors + launch sequence compiled for the host (test infrastructure)."""
    _lib = None

    @classmethod
    def lib(cls):
        if cls._lib is None:
            src = os.path.join(HERE, "emu", "emu_depthmap.cpp")
            so = os.path.join(HERE, "emu", "libemu_depthmap.so")
            csrc = os.path.join(HERE, "..", "shinestacker_b200", "csrc")
            deps = [src] + [os.path.join(csrc, f) for f in ("ss_dm_kernels.cuh", "ss_dm_engine.h", "ss_kernels.cuh")] + [os.path.join(HERE, "emu", "emu_backend.h")]
            if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(d) for d in deps):
                subprocess.run(["g++", "-std=c++17", "-O2", "-mfma", "-ffp-contract=off", "-shared", "-fPIC",
                                "-o", so, src], check=True)
            lib = ctypes.CDLL(so)
            vp, i, sz = ctypes.c_void_p, ctypes.c_int, ctypes.c_size_t
            lib.emu_dm_create.restype = vp
            lib.emu_dm_create.argtypes = [i] * 8 + [ctypes.c_double, i, i]
            lib.emu_dm_error.restype = ctypes.c_char_p
            lib.emu_dm_launches.restype = ctypes.c_longlong
            for f in ("emu_dm_destroy", "emu_dm_reset", "emu_dm_finish_begin", "emu_dm_rerun", "emu_dm_error",
                      "emu_dm_launches"):
                getattr(lib, f).argtypes = [vp]
            lib.emu_dm_destroy.restype = None
            lib.emu_dm_push.argtypes = [vp, vp, sz]
            lib.emu_dm_blend.argtypes = [vp, i]
            lib.emu_dm_finish_end.argtypes = [vp, vp, sz]
            lib.emu_dm_get_plane.argtypes = [vp, i, i, vp]
            lib.emu_dm_shard_local_max.argtypes = [vp, ctypes.POINTER(ctypes.c_float)]
            lib.emu_dm_shard_finish_local.argtypes = [vp, ctypes.c_float]
            lib.emu_dm_shard_count.argtypes = [vp]
            lib.emu_dm_shard_ptrs.argtypes = [vp, ctypes.POINTER(vp)]
            lib.emu_dm_shard_connect.argtypes = [vp, ctypes.POINTER(vp), i, i, i]
            lib.emu_dm_shard_softmax_rel.argtypes = [vp]
            lib.emu_dm_shard_finish_weights.argtypes = [vp]
            lib.emu_dm_shard_reduce_band.argtypes = [vp]
            lib.emu_dm_shard_finish_end.argtypes = [vp, vp, sz]
            cls._lib = lib
        return cls._lib

    def __init__(self, h, w, dtype, n, map_type="average", energy="laplacian", kernel_size=5, blur_size=5,
                 smooth_size=15, temperature=0.1, levels=3):
        self.lib_ = self.lib()
        self.h, self.w, self.dtype, self.n = h, w, np.dtype(dtype), 0
        self.p = self.lib_.emu_dm_create(w, h, self.dtype.itemsize, int(map_type == "max"), int(energy == "sobel"),
                                         kernel_size, blur_size, smooth_size, temperature, levels, n)
        if not self.p:
            raise ValueError("emu_dm_create rejected the options")

    def close(self):
        if self.p:
            self.lib_.emu_dm_destroy(self.p)
            self.p = None

    def __del__(self):
        self.close()

    def error(self):
        return self.lib_.emu_dm_error(self.p).decode()

    def push(self, img):
        img = np.ascontiguousarray(img)
        rc = self.lib_.emu_dm_push(self.p, img.ctypes.data, img.strides[0])
        self.n += rc == 0
        return rc

    def plane(self, what, frame):
        a = np.empty((self.h, self.w), np.float32)
        rc = self.lib_.emu_dm_get_plane(self.p, {"normalized": 0, "weights": 1, "energy": 2}[what], frame, a.ctypes.data)
        assert rc == 0, self.error()
        return a

    def finish(self, stages=None):
        lib = self.lib_
        if stages is not None:
            stages["energy"] = np.array([self.plane("energy", i) for i in range(self.n)])
        assert lib.emu_dm_finish_begin(self.p) == 0, self.error()
        if stages is not None:
            stages["weights"] = np.array([self.plane("weights", i) for i in range(self.n)])
        for i in range(self.n):
            assert lib.emu_dm_blend(self.p, i) == 0, self.error()
        out = np.empty((self.h, self.w, 3), self.dtype)
        assert lib.emu_dm_finish_end(self.p, out.ctypes.data, out.strides[0]) == 0, self.error()
        return out


def emu_focus_stack(frames, stages=None, **kwargs):
    h, w = frames[0].shape[:2]
    e = Emu(h, w, frames[0].dtype, len(frames), **kwargs)
    try:
        for f in frames:
            assert e.push(f) == 0, e.error()
        return e.finish(stages)
    finally:
        e.close()


def emu_sharded_focus_stack(frames, world, root=0, stages=None, **kwargs):
    """The sharded call sequence of include/shinestack_b200_depthmap.h with `world` handles of the host build in one
    process (contiguous runs of the frames, plain pointers instead of IPC handles); returns the root's image."""
    from shinestacker_b200.sharded import shard_range
    h, w = frames[0].shape[:2]
    lib = Emu.lib()
    ranks = []
    for r in range(world):
        a, b = shard_range(len(frames), r, world)
        e = Emu(h, w, frames[0].dtype, max(b - a, 1), **kwargs)
        for f in frames[a:b]:
            assert e.push(f) == 0, e.error()
        ranks.append((e, a, b))
    try:
        maxima = []
        for e, _, _ in ranks:
            m = ctypes.c_float()
            assert lib.emu_dm_shard_local_max(e.p, ctypes.byref(m)) == 0, e.error()
            maxima.append(m.value)
        gmax = max(maxima)
        count = lib.emu_dm_shard_count(ranks[0][0].p)
        table = (ctypes.c_void_p * (world * count))()
        for r, (e, _, _) in enumerate(ranks):
            own = (ctypes.c_void_p * count)()
            assert lib.emu_dm_shard_ptrs(e.p, own) == 0, e.error()
            table[r * count:(r + 1) * count] = list(own)
        for r, (e, _, _) in enumerate(ranks):
            assert lib.emu_dm_shard_connect(e.p, table, world, r, root) == 0, e.error()
            assert lib.emu_dm_shard_finish_local(e.p, ctypes.c_float(gmax)) == 0, e.error()
        for e, _, _ in ranks:                                   # "barrier": every partial sum / maximum exists
            assert lib.emu_dm_shard_softmax_rel(e.p) == 0, e.error()        # (no-op for map_type average)
        for e, _, _ in ranks:                                   # "barrier"
            assert lib.emu_dm_shard_finish_weights(e.p) == 0, e.error()
        if stages is not None:
            stages["weights"] = np.array([e.plane("weights", i) for e, a, b in ranks for i in range(b - a)])
        for e, a, b in ranks:
            for i in range(b - a):
                assert lib.emu_dm_blend(e.p, i) == 0, e.error()
        for e, _, _ in ranks:                                   # "barrier", then every rank sums its band of rows
            assert lib.emu_dm_shard_reduce_band(e.p) == 0, e.error()
        out = np.empty((h, w, 3), frames[0].dtype)
        for r, (e, _, _) in enumerate(ranks):
            if r != root:
                assert lib.emu_dm_shard_finish_end(e.p, None, 0) == 0, e.error()
        e = ranks[root][0]
        assert lib.emu_dm_shard_finish_end(e.p, out.ctypes.data, out.strides[0]) == 0, e.error()
        return out
    finally:
        for e, _, _ in ranks:
            e.close()
