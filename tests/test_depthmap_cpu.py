"""CPU tests of the depth-map path: the numpy oracle against the golden
vectors of the live reference and -- when cv2 / the reference tree are present -- against cv2's primitives
and the unmodified reference itself; the product's kernels + launch sequence, compiled for the host
(tests/emu/emu_depthmap.cpp), against the oracle and the golden vectors; the host mirror's constructor /
option contract.  Nothing here needs a GPU."""
import numpy as np
import pytest

import depthmap_util as du
from oracle import depthmap_oracle as dmo
from oracle import ref_harness
from shinestacker_b200 import DepthMapStack, InvalidOptionError, NativeLibraryError, synthetic, _native

needs_reference = pytest.mark.skipif(not ref_harness.reference_available(),
                                     reason="reference tree not present on this box")


@pytest.mark.parametrize("name", du.case_names())
def test_oracle_matches_golden(name):
    frames, kwargs, expect = du.load_case(name)
    out, st = dmo.focus_stack(frames, return_stages=True, **kwargs)
    du.assert_within_1lsb(out, expect["fused"], name)
    assert np.abs(st["weights"][0] - expect["weights0"]).max() < 2e-3   # float16 fixture


@pytest.mark.parametrize("name", du.case_names())
def test_kernels_on_host_match_golden_and_oracle(name):
    frames, kwargs, expect = du.load_case(name)
    st = {}
    out = du.emu_focus_stack(frames, stages=st, **kwargs)
    du.assert_within_1lsb(out, expect["fused"], name + " vs reference")
    ref, ost = dmo.focus_stack(frames, return_stages=True, **kwargs)
    du.assert_within_1lsb(out, ref, name + " vs oracle")
    # the energy stage follows the oracle's operation order exactly; the focus maps agree to float32 rounding
    assert np.array_equal(st["energy"], ost["energy"])
    assert np.abs(st["weights"] - ost["weights"]).max() < 2e-6


def test_cv2_primitives_match_the_restatement():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(5)
    for dt, mx in ((np.uint8, 256), (np.uint16, 65536)):
        img = rng.integers(0, mx, (41, 67, 3)).astype(dt)
        assert np.array_equal(dmo.bgr2gray(img), cv2.cvtColor(img, cv2.COLOR_BGR2GRAY))
    g = rng.integers(0, 256, (53, 70)).astype(np.float32)
    for ks in (3, 5, 7, 9):
        assert np.array_equal(dmo.gaussian_blur(g, ks), cv2.GaussianBlur(g, (ks, ks), 0))
    assert np.allclose(dmo.gaussian_blur(g, 13), cv2.GaussianBlur(g, (13, 13), 0), rtol=1e-6, atol=1e-4)
    bl = cv2.GaussianBlur(g, (5, 5), 0)
    for ks in (5, 7, 9):
        kd, ksm = cv2.getDerivKernels(2, 0, ks)
        a, b = dmo.deriv_kernels(ks)
        assert np.array_equal(kd.ravel(), a) and np.array_equal(ksm.ravel(), b)
        assert np.array_equal(dmo.laplacian(bl, ks), cv2.Laplacian(bl, cv2.CV_64F, ksize=ks))  # exact for u8 data
    sob = np.abs(cv2.Sobel(g, cv2.CV_64F, 1, 0, ksize=3)) + np.abs(cv2.Sobel(g, cv2.CV_64F, 0, 1, ksize=3))
    assert np.array_equal(dmo.sobel_abs_sum(g), sob)
    e = (rng.random((60, 83)) ** 4).astype(np.float32)
    for d in (15, 5, 8):
        assert np.abs(dmo.bilateral(e, d) - cv2.bilateralFilter(e, d, 25, 25)).max() < 3e-7
    flat = np.full((20, 30), 0.25, np.float32)
    assert np.array_equal(dmo.bilateral(flat, 15), cv2.bilateralFilter(flat, 15, 25, 25))
    a = (rng.random((37, 53)) * 65535).astype(np.float32)
    assert np.allclose(dmo.pyr_down(a), cv2.pyrDown(a), rtol=4e-7, atol=0)
    small = dmo.pyr_down(a)
    for dh, dw in ((37, 53), (38, 54), (38, 53)):
        assert np.allclose(dmo.pyr_up(small, dh, dw), cv2.pyrUp(small, dstsize=(dw, dh)), rtol=4e-7, atol=0)


@needs_reference
@pytest.mark.parametrize("n,h,w,dtype,seed,kwargs", [
    (3, 96, 130, np.uint8, 51, {}),
    (2, 75, 101, np.uint16, 52, {"map_type": "max", "levels": 2}),
    (4, 64, 80, np.uint8, 53, {"energy": "sobel", "smooth_size": 7}),
    (3, 70, 70, np.uint16, 54, {"smooth_size": 0, "kernel_size": 7, "blur_size": 11}),
])
def test_oracle_matches_live_reference(n, h, w, dtype, seed, kwargs):
    frames = synthetic.make_noise_stack(n, h, w, np.dtype(dtype), seed)
    ref, rec = ref_harness.reference_depthmap_arrays(frames, **kwargs)
    out, st = dmo.focus_stack(frames, return_stages=True, **kwargs)
    du.assert_within_1lsb(out, ref, "oracle vs live reference")
    # 16-bit stacks: the float32 Laplacian itself carries rounding noise of ~1e-7 of the LARGEST energy, which
    # shows up in the focus map of low-energy pixels
    assert np.abs(rec["weights"] - st["weights"]).max() < (2e-6 if dtype == np.uint8 else 1e-4)
    if "normalized" in rec:
        # u16 stacks: cv2's SIMD Laplacian rounds in a different order (absolute, not relative, agreement)
        assert np.abs(rec["normalized"] - st["energy"] / st["energy"].max()).max() < 2e-6


def test_reference_rejects_small_laplacian_kernels_and_so_do_we():
    # cv2.Laplacian(float32 -> CV_64F, ksize <= 3) raises inside the reference; the oracle and the engine refuse it too
    with pytest.raises(ValueError):
        dmo.laplacian(np.zeros((8, 8), np.float32), 3)
    with pytest.raises(ValueError):
        du.Emu(32, 32, np.uint8, 2, kernel_size=3)
    for bad in (dict(blur_size=4), dict(levels=0), dict(kernel_size=6), dict(smooth_size=200)):
        with pytest.raises(ValueError):
            du.Emu(32, 32, np.uint8, 2, **bad)


def test_engine_state_machine_and_reuse():
    frames = synthetic.make_noise_stack(3, 40, 56, np.uint8, 61)
    e = du.Emu(40, 56, np.uint8, 3)
    lib = e.lib_
    assert lib.emu_dm_finish_begin(e.p) != 0 and "no frames" in e.error()
    for f in frames:
        assert e.push(f) == 0
    assert e.push(frames[0]) != 0 and "max_frames" in e.error()
    assert lib.emu_dm_blend(e.p, 0) != 0                      # before finish_begin
    assert lib.emu_dm_finish_begin(e.p) == 0
    assert e.push(frames[0]) != 0                             # after finish
    assert lib.emu_dm_blend(e.p, 1) != 0                      # out of order
    out = np.empty((40, 56, 3), np.uint8)
    assert lib.emu_dm_finish_end(e.p, out.ctypes.data, out.strides[0]) != 0   # not every frame blended
    for i in range(3):
        assert lib.emu_dm_blend(e.p, i) == 0
    assert lib.emu_dm_finish_end(e.p, out.ctypes.data, out.strides[0]) == 0
    first = out.copy()
    # whole chain again over the resident frames (bench path), then a different, shorter stack on the same handle
    assert lib.emu_dm_rerun(e.p) == 0
    assert lib.emu_dm_launches(e.p) == 3 * 1 + 2 + 5 + 3     # energy: one tiled kernel per frame; bilateral (normalising) + focus map; blend: one batch
    out2 = np.empty_like(out)
    assert lib.emu_dm_finish_end(e.p, out2.ctypes.data, out2.strides[0]) == 0  # second call: download only
    assert np.array_equal(out2, first)
    e.close()
    e = du.Emu(40, 56, np.uint8, 3)
    for f in frames[::-1][:2]:
        assert e.push(f) == 0
    other = e.finish()
    assert lib.emu_dm_reset(e.p) == 0
    e.n = 0
    for f in frames:
        assert e.push(f) == 0
    again = e.finish()
    assert np.array_equal(again, first) and not np.array_equal(other, first)
    e.close()


def test_identical_frames_reproduce_the_input():
    # weights sum to one and the Laplacian pyramid telescopes: a stack of identical frames fuses to that frame
    img = synthetic.make_noise_stack(1, 90, 123, np.uint16, 62)[0]
    out = du.emu_focus_stack([img] * 4)
    du.assert_within_1lsb(out, img, "identical frames", max_frac=0.5)


def test_host_mirror_contract():
    a = DepthMapStack()
    assert a.name() == "depth map" and a.steps_per_frame() == 2 and a.float_type is np.float32
    assert (a.map_type, a.energy, a.kernel_size, a.blur_size, a.smooth_size, a.temperature, a.levels) == \
        ("average", "laplacian", 5, 5, 15, 0.1, 3)
    with pytest.raises(InvalidOptionError):
        DepthMapStack(float_type="float-16")
    frames = [np.zeros((32, 32, 3), np.uint8)] * 2
    # like the reference, bad energy / map_type surface when the stack runs (depth_map.py:83-87, :61-62)
    with pytest.raises(InvalidOptionError):
        DepthMapStack(energy="tenengrad").focus_stack_arrays(frames)
    with pytest.raises(InvalidOptionError):
        DepthMapStack(map_type="median").focus_stack_arrays(frames)
    if _native.load().ss_device_count() == 0:
        with pytest.raises(NativeLibraryError):     # no silent CPU fallback
            DepthMapStack().focus_stack_arrays(frames)


@pytest.mark.parametrize("n,h,w,dtype,kwargs", [
    (1, 40, 40, np.uint8, {}),                                   # a one-frame stack: weights are 1 where defined
    (2, 9, 11, np.uint8, {}),                                    # image smaller than the bilateral window
    (2, 16, 300, np.uint16, {"levels": 4}),
    (3, 31, 17, np.uint8, {"smooth_size": 31, "blur_size": 15, "kernel_size": 9}),
    (2, 20, 20, np.uint8, {"levels": 6}),                        # pyramid runs down to 1 x 1
    (2, 5, 5, np.uint8, {"smooth_size": 3, "levels": 2}),
])
def test_kernels_on_host_edge_geometries(n, h, w, dtype, kwargs):
    rng = np.random.default_rng(h * w)
    frames = [rng.integers(0, np.iinfo(dtype).max + 1, (h, w, 3)).astype(dtype) for _ in range(n)]
    du.assert_within_1lsb(du.emu_focus_stack(frames, **kwargs), dmo.focus_stack(frames, **kwargs), "edge geometry",
                          max_frac=0.2)


def test_flat_stack_documented_deviation():
    # zero energy in every frame: the reference's np.divide(where=) leaves the weights uninitialised; here they are 0
    frames = [np.full((30, 40, 3), 77, np.uint8)] * 3
    assert not du.emu_focus_stack(frames).any() and not dmo.focus_stack(frames).any()


def test_randomised_options_host_build_vs_oracle():
    rng = np.random.default_rng(77)
    for t in range(14):
        dtype = [np.uint8, np.uint16][t % 2]
        n, h, w = int(rng.integers(1, 6)), int(rng.integers(12, 80)), int(rng.integers(12, 100))
        kwargs = dict(map_type=str(rng.choice(["average", "max"])), energy=str(rng.choice(["laplacian", "sobel"])),
                      kernel_size=int(rng.choice([5, 7, 9])), blur_size=int(rng.choice([1, 3, 5, 7, 11])),
                      smooth_size=int(rng.choice([0, 3, 8, 15])), temperature=float(rng.choice([0.05, 0.1, 1.0])),
                      levels=int(rng.integers(1, 5)))
        frames = synthetic.make_noise_stack(n, h, w, np.dtype(dtype), 100 + t)
        st, ost = {}, None
        out = du.emu_focus_stack(frames, stages=st, **kwargs)
        ref, ost = dmo.focus_stack(frames, return_stages=True, **kwargs)
        du.assert_within_1lsb(out, ref, f"case {t} {kwargs}", max_frac=0.2)
        assert np.array_equal(st["energy"], ost["energy"]), (t, kwargs)
        assert np.abs(st["weights"] - ost["weights"]).max() < 5e-6, (t, kwargs)


@pytest.mark.parametrize("h,w,dtype,kwargs", [
    (70, 150, np.uint8, {}),                                           # several 64 x 32 tiles, ragged last ones
    (33, 65, np.uint16, {}),                                           # one pixel into the second tile row / column
    (32, 64, np.uint8, {"energy": "sobel"}),                           # exactly one tile
    (97, 131, np.uint8, {"kernel_size": 9, "blur_size": 11}),          # generic radii (halo 9 > tile edge reflections)
    (5, 7, np.uint8, {"kernel_size": 7, "blur_size": 7}),              # image smaller than the halo: multiple reflections
    (40, 200, np.uint16, {"energy": "sobel"}),
    (66, 34, np.uint8, {"blur_size": 1}),                              # laplacian without blur
    (100, 200, np.uint8, {}),                                          # interior tiles: the four-pixel path, uint8
    (104, 204, np.uint16, {}),                                         # ... uint16
    (100, 200, np.uint8, {"energy": "sobel"}),                         # ... without blur stages
    (100, 200, np.uint16, {"blur_size": 3}),                           # radii (1, 2): the generic instantiation
    (100, 202, np.uint8, {}),                                          # width not a multiple of 4: scalar path only
])
def test_tiled_energy_kernel_equals_the_four_streaming_passes(h, w, dtype, kwargs, monkeypatch):
    frames = synthetic.make_noise_stack(2, h, w, np.dtype(dtype), h + w)
    planes = []
    for fused in ("1", "0"):
        monkeypatch.setenv("SSDM_ENERGY_FUSED", fused)
        e = du.Emu(h, w, dtype, 2, **kwargs)
        for f in frames:
            assert e.push(f) == 0, e.error()
        planes.append(np.array([e.plane("energy", i) for i in range(2)]))
        launches = e.lib_.emu_dm_launches(e.p)
        assert launches == (2 if fused == "1" else 2 * (3 if kwargs.get("energy") == "sobel" else 4))
        e.close()
    assert np.isfinite(planes[0]).all()
    assert np.array_equal(planes[0], planes[1])
    ref = dmo.focus_stack(frames, return_stages=True, **kwargs)[1]["energy"]
    assert np.array_equal(planes[0], ref)


@pytest.mark.parametrize("dtype,levels", [(np.uint8, 3), (np.uint16, 4), (np.uint8, 1)])
def test_blend_batches_equal_frame_by_frame_blending(dtype, levels, monkeypatch):
    frames = synthetic.make_stack(5, 45, 70, np.dtype(dtype), seed=9)
    outs = []
    for batch in ("1", "2", "8"):
        monkeypatch.setenv("SSDM_BLEND_BATCH", batch)
        outs.append(du.emu_focus_stack(frames, levels=levels))
    assert np.array_equal(outs[0], outs[1]) and np.array_equal(outs[0], outs[2])
    du.assert_within_1lsb(outs[0], dmo.focus_stack(frames, levels=levels), "batched blend", max_frac=0.2)


@pytest.mark.parametrize("dtype,h,w,levels", [(np.uint16, 75, 140, 3), (np.uint8, 33, 65, 4), (np.uint16, 64, 128, 2),
                                              (np.uint8, 7, 9, 3)])
def test_tiled_pyrdown_equals_the_per_pixel_kernel(dtype, h, w, levels, monkeypatch):
    frames = synthetic.make_noise_stack(3, h, w, np.dtype(dtype), h * w)
    outs = []
    for tiled in ("1", "0"):
        monkeypatch.setenv("SSDM_PYRDOWN_TILED", tiled)
        outs.append(du.emu_focus_stack(frames, levels=levels))
    assert np.array_equal(outs[0], outs[1])


def test_bilateral_forms_agree_with_the_oracle(monkeypatch):
    # default: tiled, three moments, normalisation folded into the load; SSDM_BILATERAL_MOMENTS=5: full polynomial;
    # SSDM_BILATERAL_PAIR=1: the direct 149-tap kernel after a separate normalisation pass
    frames = synthetic.make_stack(4, 90, 140, np.dtype(np.uint16), seed=5)
    frames[1] = synthetic.make_noise_stack(1, 90, 140, np.dtype(np.uint16), 3)[0]     # large neighbour differences
    ref, ost = dmo.focus_stack(frames, return_stages=True)
    for env in ({}, {"SSDM_BILATERAL_MOMENTS": "5"}, {"SSDM_BILATERAL_PAIR": "1"}):
        for k in ("SSDM_BILATERAL_MOMENTS", "SSDM_BILATERAL_PAIR"):
            monkeypatch.delenv(k, raising=False)
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        st = {}
        out = du.emu_focus_stack(frames, stages=st)
        assert np.abs(st["weights"] - ost["weights"]).max() < 1e-6, env
        du.assert_within_1lsb(out, ref, str(env), max_frac=0.2)


@pytest.mark.parametrize("world,root,n,dtype,kwargs", [
    (2, 0, 5, np.uint8, {}),
    (3, 2, 7, np.uint16, {"levels": 4}),
    (4, 1, 4, np.uint8, {"energy": "sobel", "smooth_size": 0}),       # one frame per rank, no smoothing
    (1, 0, 3, np.uint8, {}),                                          # a "box" of one GPU goes the same way
    (3, 0, 7, np.uint8, {"map_type": "max"}),                         # softmax: maximum and sum cross the ranks
    (2, 1, 5, np.uint16, {"map_type": "max", "temperature": 0.05, "levels": 2}),
])
def test_frames_sharded_over_ranks_give_the_single_handle_result_within_rounding(world, root, n, dtype, kwargs):
    frames = synthetic.make_stack(n, 60, 90, np.dtype(dtype), seed=world + n)
    st, sst = {}, {}
    single = du.emu_focus_stack(frames, stages=st, **kwargs)
    sharded = du.emu_sharded_focus_stack(frames, world, root, stages=sst, **kwargs)
    # the sums over the stack are associated rank by rank: float32 rounding only
    assert np.abs(st["weights"] - sst["weights"]).max() < 1e-6
    du.assert_within_1lsb(sharded, single, "sharded", max_frac=0.02)
    du.assert_within_1lsb(sharded, dmo.focus_stack(frames, **kwargs), "sharded vs oracle", max_frac=0.2)
    if world == 1:
        assert np.array_equal(sharded, single)


class _EmuNative:
    """The ssdm_* entry points ShardedDepthMapStack.finish uses, served by the host build of the kernels: the real host
    orchestration (rendezvous, barriers, phase order) runs on a CPU-only box.  'IPC handles' are plain pointers, valid
    because the ranks are threads of one process."""

    def __init__(self):
        self.emu = du.Emu.lib()

    def ssdm_shard_local_max(self, h, ref):
        return self.emu.emu_dm_shard_local_max(h, ref)

    def ssdm_shard_finish_local(self, h, gmax):
        return self.emu.emu_dm_shard_finish_local(h, gmax)

    def ssdm_shard_buffer_count(self, h):
        return self.emu.emu_dm_shard_count(h)

    def ssdm_shard_export(self, h, blob):
        import ctypes
        count = self.emu.emu_dm_shard_count(h)
        ptrs = (ctypes.c_void_p * count)()
        rc = self.emu.emu_dm_shard_ptrs(h, ptrs)
        for k in range(count):
            blob[64 * k:64 * k + 8] = list((ptrs[k] or 0).to_bytes(8, "little"))
        return rc

    def ssdm_shard_connect(self, h, blobs, world, rank, root):
        import ctypes
        count = self.emu.emu_dm_shard_count(h)
        table = (ctypes.c_void_p * (world * count))()
        for i in range(world * count):
            table[i] = int.from_bytes(blobs[64 * i:64 * i + 8], "little")
        return self.emu.emu_dm_shard_connect(h, table, world, rank, root)

    def ssdm_sync(self, h):
        return 0

    def ssdm_shard_softmax_rel(self, h):
        return self.emu.emu_dm_shard_softmax_rel(h)

    def ssdm_shard_finish_weights(self, h):
        return self.emu.emu_dm_shard_finish_weights(h)

    def ssdm_blend_frame(self, h, i):
        return self.emu.emu_dm_blend(h, i)

    def ssdm_shard_reduce_band(self, h):
        return self.emu.emu_dm_shard_reduce_band(h)

    def ssdm_shard_finish_end(self, h, out, pitch):
        return self.emu.emu_dm_shard_finish_end(h, out, pitch)

    def ssdm_kernel_time_ms(self, h, ref):
        return 0


@pytest.mark.parametrize("map_type", ["average", "max"])
def test_sharded_host_orchestration_with_two_ranks_over_tcp(map_type):
    import ctypes
    import socket
    import threading
    from shinestacker_b200 import rendezvous
    from shinestacker_b200.sharded import ShardedDepthMapStack, shard_range
    frames = synthetic.make_stack(5, 50, 70, np.dtype(np.uint8), seed=12)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    results, errors = {}, []

    def rank_main(rank):
        try:
            rdv = rendezvous.TcpRendezvous(rank, 2, "127.0.0.1", port, timeout=30)
            a, b = shard_range(len(frames), rank, 2)
            emu = du.Emu(50, 70, np.uint8, b - a, map_type=map_type)
            for f in frames[a:b]:
                assert emu.push(f) == 0
            st = ShardedDepthMapStack(rank, 2, rendezvous=rdv, root=1, map_type=map_type)
            st.algo._lib, st.algo._h = _EmuNative(), ctypes.c_void_p(emu.p)
            results[rank] = st.finish(b - a, (50, 70), np.dtype(np.uint8))
            st.algo._h = None
            rdv.close()
            emu.close()
        except Exception as e:  # pylint: disable=broad-except
            errors.append(repr(e))

    threads = [threading.Thread(target=rank_main, args=(r,)) for r in range(2)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=60)
    assert not errors, errors
    assert results[0] is None                                 # rank 1 is the root
    assert np.array_equal(results[1], du.emu_sharded_focus_stack(frames, 2, root=1, map_type=map_type))


def test_randomised_geometries_of_the_tiled_kernels_against_the_oracle():
    # sizes with interior tiles (four-pixel path of the energy kernel, unreflected patches of the bilateral and of the
    # tiled pyrDown) next to ragged borders, both sample types, both energies
    rng = np.random.default_rng(2024)
    for t in range(8):
        dtype = [np.uint8, np.uint16][t % 2]
        h = int(rng.integers(64, 200))
        w = int(rng.integers(34, 75)) * 4 if t % 4 else int(rng.integers(136, 300))
        n = int(rng.integers(2, 5))
        kwargs = dict(energy=["laplacian", "sobel"][(t // 2) % 2], levels=int(rng.integers(2, 5)))
        frames = synthetic.make_noise_stack(n, h, w, np.dtype(dtype), 500 + t)
        st = {}
        out = du.emu_focus_stack(frames, stages=st, **kwargs)
        ref, ost = dmo.focus_stack(frames, return_stages=True, **kwargs)
        assert np.array_equal(st["energy"], ost["energy"]), (t, h, w, kwargs)
        assert np.abs(st["weights"] - ost["weights"]).max() < 5e-6, (t, h, w, kwargs)
        du.assert_within_1lsb(out, ref, f"case {t} {h}x{w} {kwargs}", max_frac=0.2)
