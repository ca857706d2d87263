#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 focus-stacking core.

Metric (BASELINE.json): input megapixels per second fused, synthetic 4096x3072 RGB8, 64-frame stack,
6-level Laplacian pyramid (configs[1]).  One "step" = one complete stack (pyramid build + energy +
arg-max select over the frames + base-level fusion + collapse).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

N > 1 is launched under torchrun (one rank per GPU); the frames of the ONE stack are sharded over the
ranks (strong scaling) and merged by the library's fused peer exchange over NVLink (every rank merges and collapses
one band of image rows straight out of the peers' memory; shinestacker_b200/sharded.py, include/shinestack_b200.h).
torch.distributed (NCCL) is used by this script only for the timing plumbing the bench contract prescribes (barrier,
max over ranks); the product path has no PyTorch and no collective library.

Printed JSON (rank 0, one line): value = throughput with the frames already resident in HBM (fused image left in
HBM; it is downloaded once after the timed loop for the parity check);
e2e = the same stack through the public host-buffer API (H2D of every frame from pinned host memory
and D2H of the fused image inside the timed region); roofline = level-0 fused kernel, algorithmic
bytes / CUDA-event time against MEASURED_PEAKS.json; cpu_baseline = the UNMODIFIED reference (baseline/_ref, its own
StackJob -> FocusStack -> PyramidStack from TIFF files) on a bounded sample of the same workload, with the CPU oracle
(a multi-threaded C port of the reference's arithmetic, oracle/) on the whole stack beside it; parity = the GPU result
of the timed run compared with the oracle's result on the same full stack (fused image and level-0 arg-max).
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "input megapixels/sec fused (4096x3072 RGB8, 64-frame stack, 6-level Laplacian pyramid)"
UNIT = "MP/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--frames", type=int, default=64)
    ap.add_argument("--height", type=int, default=3072)
    ap.add_argument("--width", type=int, default=4096)
    ap.add_argument("--dtype", default="uint8", choices=["uint8", "uint16"])
    ap.add_argument("--batch", type=int, default=8, help="frames per kernel-chain launch on the e2e path")
    ap.add_argument("--cpu-sample-frames", type=int, default=0,
                    help="frames of the stack the CPU oracle port fuses (0 = the whole stack, which also enables the "
                         "parity check; ~10 s for config 2 on 16 cores)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--ref-sample-frames", type=int, default=0,
                    help="frames per stack the unmodified reference fuses (0 = bounded automatically: ~10-30 s of CPU)")
    ap.add_argument("--check-parity", default="auto", choices=["auto", "on", "off"],
                    help="compare the GPU result with the oracle on the full stack (auto: whenever the CPU baseline "
                         "runs on all frames at N=1; 'on' also at N>1, rank 0 runs the streaming oracle)")
    ap.add_argument("--no-next-rows", action="store_true",
                    help="skip the short DepthMapStack / AlignFrames measurements appended as 'next_rows' (N=1 only)")
    return ap.parse_args()


def workload_name(a):
    bits = 8 if a.dtype == "uint8" else 16
    return f"synthetic {a.width}x{a.height} RGB{bits}, {a.frames} frames"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, power, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for nm, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        load = [s for s, p in zip(sm, power) if p >= 0.5 * max(power)] or sm
        return {"sm_mhz": float(np.median(load)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm), "power_w_max": float(max(power))}


def cpu_port(a, n_sample, threads=None, frames=None):
    """The oracle (CPU port of the reference arithmetic, C + OpenMP) on the first n_sample frames of the same
    synthetic stack.  Returns (entry, fused image, info) -- the result is what the parity check compares with."""
    from oracle import pyramid_oracle as po
    from shinestacker_b200 import synthetic
    po.build()
    cores = po.set_threads(threads or os.cpu_count() or 1)
    if frames is None:
        frames = synthetic.make_stack(a.frames, a.height, a.width, np.dtype(a.dtype), seed=2)
    frames = frames[:n_sample]
    t0 = time.perf_counter()
    out, info = po.focus_stack(frames)
    dt = time.perf_counter() - t0
    mp = n_sample * a.height * a.width / 1e6
    return {"value": mp / dt, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"first {n_sample} of the {a.frames} frames ({workload_name(a)}), oracle/pyramid_oracle "
                      f"(C + OpenMP, {cores} threads), {dt:.2f} s"}, out, info


def reference_stack_time(a, frames, steps=1, warmup=0):
    """The UNMODIFIED reference (baseline/_ref): its own StackJob -> FocusStack -> PyramidStack (the path of the
    reference's tests/test_0060_stack.py:7-14) over `frames` written as uncompressed TIFFs to a RAM disk (the format
    its own write_img produces, utils.py:23-30).  Returns (seconds per stack, fused image) or None if it is absent."""
    from tools import ref_loader
    if ref_loader.load() is None:
        return None
    import shutil
    import tempfile
    import cv2
    from shinestacker.algorithms.pyramid import PyramidStack as RefPyramidStack
    base = "/dev/shm" if os.path.isdir("/dev/shm") else None
    work = tempfile.mkdtemp(prefix="ss_ref_", dir=base)
    try:
        os.makedirs(os.path.join(work, "in"))
        for i, f in enumerate(frames):
            cv2.imwrite(os.path.join(work, "in", f"f{i:04d}.tif"), f, [int(cv2.IMWRITE_TIFF_COMPRESSION), 1])
        out_path = None
        for _ in range(warmup):
            ref_loader.run_stack_job(work, "in", RefPyramidStack(), output_subdir="out")
        t0 = time.perf_counter()
        for _ in range(steps):
            out_path = ref_loader.run_stack_job(work, "in", RefPyramidStack(), output_subdir="out")
        dt = (time.perf_counter() - t0) / steps
        img = cv2.imread(out_path, cv2.IMREAD_UNCHANGED) if out_path else None
        return dt, img
    finally:
        shutil.rmtree(work, ignore_errors=True)


def cpu_baseline(a, frames):
    """cpu_baseline of our arm: the unmodified reference on a bounded sample (kind "reference"); without
    baseline/_ref the oracle port stands in (kind "port")."""
    from tools import ref_loader
    n = a.ref_sample_frames if a.ref_sample_frames > 0 else (4 if a.height * a.width > 4e6 else 8)
    n = max(2, min(n, len(frames)))
    res = reference_stack_time(a, frames[:n])
    if res is None:
        return None
    dt, _ = res
    mp = n * a.height * a.width / 1e6
    return {"value": mp / dt, "unit": UNIT, "cores": os.cpu_count() or 1, "kind": "reference",
            "sample": f"first {n} of the {a.frames} frames ({workload_name(a)}) through the unmodified reference "
                      f"(StackJob -> FocusStack -> PyramidStack, uncompressed TIFFs on a RAM disk -> written image), "
                      f"1 process, {os.cpu_count()} cores visible, {ref_loader.versions()}, {dt:.2f} s; the path is "
                      f"O(frames) and effectively single-threaded"}


def next_rows(a, host):
    """SURVEY.md section 8f rows built beyond the headline path, measured on the same frames (not part of
    `value`): DepthMapStack (device-resident chain + end to end from host frames) and the AlignFrames warp
    stage (one frame: warp + mask + 21x21 blur fill) with cv2 on the host cores beside it."""
    out = {}
    from tools import time_depthmap, time_align, time_ecc, time_pipeline, time_files
    jobs = [("depth_map", lambda: time_depthmap.measure(a.frames, a.height, a.width, a.dtype, steps=3, warmup=1,
                                                        host=list(host), pinned=True)),
            ("align_warp", lambda: time_align.measure(a.height, a.width, a.dtype, steps=5)),
            ("files_e2e", lambda: time_files.measure(list(host), steps=2))]
    if a.dtype == "uint8":
        jobs += [("align_ecc_estimator", lambda: time_ecc.measure(a.height, a.width, steps=3, with_features=True)),
                 ("align_ecc_plus_stack_config5", lambda: time_pipeline.measure(list(host), steps=2))]
    for name, job in jobs:      # each entry on its own: a failure must not take the others (or the headline) with it
        try:
            out[name] = job()
        except Exception as e:  # pylint: disable=broad-except
            out[name] = {"error": f"{type(e).__name__}: {e}"}
    return out


def run_reference(a):
    """--impl reference: the reference's own CPU implementation of the path on this box's host cores -- the UNMODIFIED
    reference from baseline/_ref through its public module API (StackJob -> FocusStack -> PyramidStack, TIFF files in,
    image file out).  Each step fuses a bounded sample of the workload's frames, sized so that the K + W steps end
    within a few minutes (the reference needs ~2.5 s and ~0.6 GB per 12.6 MP frame and is O(frames)).  If the
    reference is not installed the oracle port (oracle/, all host threads) stands in and says so."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from shinestacker_b200 import synthetic
    from tools import ref_loader
    mpx = a.height * a.width / 1e6
    if ref_loader.load() is not None:
        budget_frames = max(2, int(240.0 / (0.21 * mpx)))          # ~0.21 s per megapixel-frame measured here
        n_sample = a.ref_sample_frames if a.ref_sample_frames > 0 else budget_frames // max(1, a.steps + a.warmup)
        n_sample = max(2, min(n_sample, a.frames, 16))
        frames = synthetic.make_stack(a.frames, a.height, a.width, np.dtype(a.dtype), seed=2)[:n_sample]
        dt, _ = reference_stack_time(a, frames, steps=a.steps, warmup=a.warmup)
        kind, cores = "reference", os.cpu_count() or 1
        sample = (f"each step = first {n_sample} of the {a.frames} frames ({workload_name(a)}) through the unmodified "
                  f"reference StackJob -> FocusStack -> PyramidStack (baseline/_ref; uncompressed TIFFs on a RAM disk in, "
                  f"image file out), 1 process, {cores} cores visible, {ref_loader.versions()}")
    else:
        from oracle import pyramid_oracle as po
        po.build()
        cores = po.set_threads(os.cpu_count() or 1)
        n_sample = max(1, min(a.ref_sample_frames or 16, a.frames, 16))
        frames = synthetic.make_stack(a.frames, a.height, a.width, np.dtype(a.dtype), seed=2)[:n_sample]
        for _ in range(a.warmup):
            po.focus_stack(frames[:1])
        t0 = time.perf_counter()
        for _ in range(a.steps):
            po.focus_stack(frames)
        dt = (time.perf_counter() - t0) / a.steps
        kind = "port"
        sample = (f"baseline/_ref absent: each step = first {n_sample} of the {a.frames} frames ({workload_name(a)}) "
                  f"through oracle/pyramid_oracle (C + OpenMP port of reference pyramid.py, {cores} threads)")
    val = n_sample * mpx / dt
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": a.gpus,
        "steps": a.steps, "warmup": a.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(a), "levels": 6, "sample_frames": n_sample},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def main():
    a = parse()
    if a.impl == "reference":
        run_reference(a)
        return
    import torch
    import torch.distributed as dist
    from shinestacker_b200 import synthetic, _native
    from shinestacker_b200.sharded import ShardedPyramidStack, shard_range, bind_to_gpu_numa_node
    from shinestacker_b200.pyramid import _check

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    numa_node = None
    if world > 1:
        numa_node = bind_to_gpu_numa_node(local)   # before any pinned allocation
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    lib = _native.load()
    if lib.ss_device_count() < 1:
        raise SystemExit("bench.py: no sm_100 GPU visible (there is no CPU fallback path)")
    from shinestacker_b200 import rendezvous
    rdv = rendezvous.TcpRendezvous.from_env()      # bootstrap of the peer exchange (one all-gather per native handle)

    def barrier():
        torch.cuda.synchronize(local)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(local)

    dtype = np.dtype(a.dtype)
    H, W, N = a.height, a.width, a.frames
    f0, f1 = shard_range(N, rank, world)
    n_local = f1 - f0
    frame_bytes = H * W * 3 * dtype.itemsize
    # synthetic stack: every rank generates the scene, keeps its own frames in PINNED host memory
    sharp, blurred = synthetic.make_scene(H, W, dtype, seed=2)
    maxv = 255 if dtype == np.uint8 else 65535
    sharp = np.clip(sharp, 0, maxv).astype(dtype)
    blurred = np.clip(blurred, 0, maxv).astype(dtype)
    pin = ctypes.c_void_p()
    _check(lib.ss_host_alloc(ctypes.byref(pin), n_local * frame_bytes))
    host = np.frombuffer((ctypes.c_uint8 * (n_local * frame_bytes)).from_address(pin.value),
                         dtype=dtype).reshape(n_local, H, W, 3)
    for i in range(n_local):
        host[i] = synthetic.frame_from_scene(sharp, blurred, f0 + i, N, dtype)
    del sharp, blurred
    out_pin = ctypes.c_void_p()
    _check(lib.ss_host_alloc(ctypes.byref(out_pin), frame_bytes))
    out = np.frombuffer((ctypes.c_uint8 * frame_bytes).from_address(out_pin.value), dtype=dtype).reshape(H, W, 3)
    pitch = W * 3 * dtype.itemsize

    # ---- leg 1: frames resident in HBM, one launch per level over all local frames ------------------
    res = ShardedPyramidStack(rank, world, device=local, rendezvous=rdv, batch_frames=n_local)
    algo = res.begin(host[0], n_local, N, resident=True)
    for i in range(n_local):
        _check(lib.ss_push_frame(algo._h, host[i].ctypes.data_as(ctypes.c_void_p), pitch))
    _check(lib.ss_sync(algo._h))
    _check(lib.ss_set_profiling(algo._h, 1))

    def step_resident():
        _check(lib.ss_rerun_resident(algo._h))
        return res.finish(n_local, (H, W), download=False)     # device-resident in, device-resident out

    # nvidia-smi needs ~100 ms to deliver its first sample: start it before the warm-up steps so that it is
    # sampling during the timed region; only samples taken under load (power) enter the median
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for _ in range(a.warmup):
        step_resident()
    lvl0_ms, lvl0_n, chain_ms, launches = 0.0, 0, 0.0, 0
    barrier()
    t0 = time.perf_counter()
    for _ in range(a.steps):
        step_resident()
        # per-step device-side accounting (event queries after the step's own sync: no extra stall)
        ms, nl = ctypes.c_float(), ctypes.c_int32()
        _check(lib.ss_level_kernel_time_ms(algo._h, 0, ctypes.byref(ms), ctypes.byref(nl)))
        lvl0_ms += ms.value; lvl0_n += nl.value
        _check(lib.ss_kernel_time_ms(algo._h, ctypes.byref(ms)))
        chain_ms += ms.value
        c = ctypes.c_int64()
        _check(lib.ss_launch_count(algo._h, ctypes.byref(c)))
        launches += c.value
    barrier()
    dt = time.perf_counter() - t0
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([dt], dtype=torch.float64, device=f"cuda:{local}")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dt_res = float(t.item())
    fused_resident = res.download((H, W), out).copy() if rank == 0 else None
    argmax_l0 = None
    if rank == 0:          # level-0 arg-max of the timed run (after the exchange the root holds the global map)
        argmax_l0 = np.empty((H, W), dtype=np.int32)
        _check(lib.ss_get_argmax(algo._h, 0, argmax_l0.ctypes.data_as(ctypes.c_void_p)))
    res.algo.close()

    # ---- leg 2: end to end through the public host-buffer API ---------------------------------------
    e2e = ShardedPyramidStack(rank, world, device=local, rendezvous=rdv, batch_frames=min(a.batch, n_local))

    def step_e2e():
        al = e2e.begin(host[0], n_local, N)
        for i in range(n_local):
            _check(lib.ss_push_frame(al._h, host[i].ctypes.data_as(ctypes.c_void_p), pitch))
        return e2e.finish(n_local, (H, W), out=out)

    for _ in range(a.warmup):
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(a.steps):
        step_e2e()
    barrier()
    dt = time.perf_counter() - t0
    t = torch.tensor([dt], dtype=torch.float64, device=f"cuda:{local}")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dt_e2e = float(t.item())
    same = bool(np.array_equal(out, fused_resident)) if rank == 0 else True
    e2e.algo.close()

    if rank == 0:
        mp = N * H * W / 1e6
        b = dtype.itemsize
        s = 1 if N <= 256 else 2
        alg_bytes = N * H * W * 3 * b + H * W * 3 * b + H * W * s
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except OSError:
            pass
        traffic, traffic_src = None, None
        try:   # dram__bytes_read.sum + dram__bytes_write.sum of one level-0 launch (ncu --set full, profiles/)
            prof = os.path.join("profiles", "r2_level0_summary.txt")
            txt = open(os.path.join(ROOT, prof)).read().split()
            def _val(key):
                i = txt.index(key)
                return float(txt[i + 1]) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}[txt[i + 2]]
            if H == 3072 and W == 4096 and dtype == np.uint8:
                # the capture is the 64-frame launch of config 2; the traffic of a launch is proportional to its frames
                # (inputs read once, g_1 planes written once per frame; the state is a few % of it)
                traffic = (_val("dram__bytes_read.sum") + _val("dram__bytes_write.sum")) * n_local / 64.0
                traffic_src = f"{prof} (ncu --set full, one 64-frame launch" + \
                              (")" if n_local == 64 else f"; scaled to the {n_local} frames of this launch)")
        except (OSError, ValueError, KeyError):
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "MEASURED_PEAKS.json hbm_gbs (measured)" if peaks else "fallback 6650 GB/s"
        # the dominant kernel: level-0 fused kernel; one launch covers this rank's n_local frames
        launch_ms = lvl0_ms / max(lvl0_n, 1)
        alg_launch = alg_bytes * (n_local / N)
        achieved = alg_launch / (launch_ms * 1e-3) / 1e9 if launch_ms > 0 else 0.0
        line = {
            "metric": METRIC, "value": mp / (dt_res / a.steps), "unit": UNIT, "n_gpus": world,
            "steps": a.steps, "warmup": a.warmup, "ms_per_step": dt_res / a.steps * 1e3,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": workload_name(a), "levels": 6, "frames_per_gpu": n_local,
                       "l2": f"inputs ({N * frame_bytes / 1e9:.2f} GB) larger than the 126 MB L2; every "
                             "step re-reads all frames from HBM",
                       "parallelism": f"frames sharded over {world} GPU(s)" + ("; fused peer exchange over NVLink "
                                      "(banded merge + collapse, epoch flags), no collective on the data path" if world > 1 else ""),
                       "output": "value: fused image left in HBM; e2e: downloaded to pinned host memory",
                       "numa_node_rank0": numa_node},
            "clocks": clocks,
            "e2e": {"value": mp / (dt_e2e / a.steps), "unit": UNIT, "ms_per_step": dt_e2e / a.steps * 1e3,
                    # frames up + log table up; fused image down + histogram (u8) / gray base levels (u16) down
                    "h2d_bytes_per_step": N * frame_bytes + N * 4 * (256 if b == 1 else 48 * 64),
                    "d2h_bytes_per_step": frame_bytes + N * 4 * (256 if b == 1 else 48 * 64),
                    "matches_resident_result": same},
            "output_crc32": int(zlib.crc32(out.tobytes())),
            "gpu_launches": launches,
            "device_chain_ms_per_step": chain_ms / a.steps,
            "roofline": {"bound": "hbm", "kernel": "ss_level_kernel (level 0: reduce + expand/Laplacian + "
                                                   "energy + arg-max select, all frames of the launch)",
                         "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "traffic_source": traffic_src, "launch_ms": launch_ms, "algorithmic_bytes_per_launch": alg_launch,
                         "peak_source": peak_src},
        }
        # the bound that actually limits this kernel: bit-exact sequential FP32 chains.
        # Algorithmic chain steps per level-0 pixel and frame: reduce 25*3/4, expand 6.25*3 + 3, gray^2 4,
        # energy 25, convert 3, compare 1 = 73.5; peak = scalar FFMA rate measured on this pool's B200s with
        # tools/microbench.cu (profiles/r1_microbench_fp32_issue.txt): 36.2 TFMA/s.
        fp32_ops = 73.5 * n_local * H * W
        fp32_rate = fp32_ops / (launch_ms * 1e-3) / 1e12 if launch_ms > 0 else 0.0
        line["roofline_fp32"] = {"bound": "fp32 issue", "achieved": fp32_rate, "peak": 36.2, "unit": "TFMA/s",
                                 "frac": fp32_rate / 36.2, "algorithmic_chain_steps_per_pixel": 73.5,
                                 "peak_source": "tools/microbench.cu on B200 (scalar FFMA, 124.5 of 128 lanes/clk/SM)"}
        # ---- CPU side: reference baseline, oracle port, parity of the timed GPU result ------------------
        n_cpu = N if a.cpu_sample_frames <= 0 else max(1, min(a.cpu_sample_frames, N))
        want_parity = a.check_parity == "on" or (a.check_parity == "auto" and world == 1 and
                                                 not a.no_cpu_baseline and n_cpu == N)
        oracle_out = oracle_info = None
        if world == 1 and not a.no_cpu_baseline:
            port, oracle_out, oracle_info = cpu_port(a, n_cpu, frames=list(host))
            ref_entry = cpu_baseline(a, list(host))
            line["cpu_baseline"] = ref_entry if ref_entry is not None else port
            line["cpu_port"] = port
        elif want_parity:      # N > 1: rank 0 streams the whole stack through the oracle (not a timed baseline)
            from oracle import pyramid_oracle as po
            po.build()
            po.set_threads(os.cpu_count() or 1)
            sharp, blurred = synthetic.make_scene(H, W, dtype, seed=2)
            sharp = np.clip(sharp, 0, maxv).astype(dtype)
            blurred = np.clip(blurred, 0, maxv).astype(dtype)
            oracle_out, oracle_info = po.focus_stack(
                synthetic.frame_from_scene(sharp, blurred, i, N, dtype) for i in range(N))
        if want_parity and oracle_out is not None:
            fused_equal = bool(np.array_equal(fused_resident, oracle_out))
            e2e_equal = bool(np.array_equal(out, oracle_out))
            argmax_equal = bool(np.array_equal(argmax_l0, oracle_info["best"][0]))
            line["parity"] = {"fused_equal": fused_equal, "e2e_fused_equal": e2e_equal,
                              "argmax_l0_equal": argmax_equal,
                              "oracle_crc32": int(zlib.crc32(np.ascontiguousarray(oracle_out).tobytes())),
                              "frames_compared": N, "checker": "oracle/pyramid_oracle on the same full stack"}
            parity_ok = fused_equal and e2e_equal and argmax_equal
        else:
            parity_ok = True
        if world == 1 and not a.no_next_rows:
            try:
                line["next_rows"] = next_rows(a, host)
            except Exception as e:  # pylint: disable=broad-except
                line["next_rows"] = {"error": f"{type(e).__name__}: {e}"}
        print(json.dumps(line))
        if not parity_ok:
            sys.stdout.flush()
            raise SystemExit("bench.py: the GPU result differs from the oracle (see \"parity\" in the line above)")
    lib.ss_host_free(pin)
    lib.ss_host_free(out_pin)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
