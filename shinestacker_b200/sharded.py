"""Frames of ONE stack sharded over the GPUs of a box, one process per GPU.

Rank r of G fuses the contiguous frame block [r*N/G, (r+1)*N/G) with the ordinary kernel chain, producing the running
state of its frames (per level: best energy, global frame index, winner coefficients; base level: best entropy /
deviation and the winners' pixels).  The ranks are then merged by the library's fused peer exchange
(include/shinestack_b200.h, ss_peer_*): every rank merges and collapses one band of image rows straight out of the
peers' memory over NVLink and stores its rows of the fused image into the root's output buffer; epoch flags in peer
memory are the only synchronisation.  The merge takes the first maximum in rank (= frame) order with a strict '>',
i.e. np.argmax over the whole stack (reference pyramid.py:51-54), so the result is bit-identical to the single-GPU
path.  No collective library and no PyTorch: the only host-side communication is one all-gather of CUDA IPC handles
per native handle (rendezvous.py).
"""
import ctypes
import os

import numpy as np

from .pyramid import PyramidStack, _check
from . import rendezvous as _rdv

PEER_BLOB = 384


def shard_range(n_frames, rank, world):
    """Contiguous block of frames for ``rank`` (lower ranks take the remainder): contiguity keeps
    'lowest global index wins' equal to 'lowest rank wins, then lowest local index'."""
    base, rem = divmod(n_frames, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def bind_to_gpu_numa_node(device):
    """Pin this process to the CPUs of the NUMA node its GPU hangs off, so that the pinned staging buffers
    (allocated afterwards) and the upload traffic stay on that socket.  Best effort: returns the node or None."""
    import subprocess
    try:
        bus = subprocess.run(["nvidia-smi", "--query-gpu=pci.bus_id", "--format=csv,noheader", "-i", str(device)],
                             capture_output=True, text=True, check=True).stdout.strip().lower()
        if bus.startswith("00000000:"):
            bus = bus[4:]
        node = int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read())
        if node < 0:
            return None
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return node
    except Exception:  # pylint: disable=broad-except
        pass
    return None


class ShardedPyramidStack:
    """One per rank.  ``focus_stack_arrays(local_frames, n_total)`` returns the fused image on the root rank and None
    elsewhere.  ``rendezvous``: an object with ``all_gather(bytes) -> list[bytes]`` (default: from the torchrun-style
    environment variables, see rendezvous.py)."""

    def __init__(self, rank, world, device=0, root=0, rendezvous=None, **kwargs):
        self.rank, self.world, self.device, self.root = rank, world, device, root
        self.kwargs = kwargs
        self.algo = None
        self.n_total = None
        self._rdv = rendezvous
        self._connected_handle = None

    def _algo(self, n_total):
        start, _ = shard_range(n_total, self.rank, self.world)
        if self.algo is None or self.n_total != n_total:
            if self.algo is not None:
                self.algo.close()
            self.algo = PyramidStack(device=self.device, frame_index_base=start, keep_depth_maps=False,
                                     **self.kwargs)
            self.n_total = n_total
            self._connected_handle = None
        return self.algo

    def _connect(self):
        """All-gather the IPC handles of the native handle's exchangeable buffers and open the peers' (once per
        native handle; the ranks create their handles in lock step, so they all get here together)."""
        a = self.algo
        if self.world == 1 or self._connected_handle == a._h.value:
            return
        if self._rdv is None:
            self._rdv = _rdv.TcpRendezvous.from_env()
        blob = (ctypes.c_uint8 * PEER_BLOB)()
        _check(a._lib.ss_peer_export(a._h, blob))
        blobs = b"".join(self._rdv.all_gather(bytes(blob)))
        _check(a._lib.ss_peer_connect(a._h, blobs, self.world, self.rank, self.root))
        self._rdv.barrier()           # nobody signals into a peer's flags before everybody has opened them
        self._connected_handle = a._h.value

    def begin(self, img0, n_local, n_total, resident=False):
        a = self._algo(n_total)
        a._set_metadata(img0)
        a._ensure_handle(img0.shape, img0.dtype, n_local, resident=resident)
        self._connect()
        return a

    def finish(self, n_local, shape, out=None, download=True):
        """Base selection on the own frames, peer exchange, banded collapse.  All ranks must call it; the root
        returns the fused image (``download=False``: the image stays on the device and True is returned)."""
        a = self.algo
        lib, h = a._lib, a._h
        if self.world == 1:
            _check(lib.ss_finish_base_device(h))
            if not download:
                _check(lib.ss_finish_collapse(h, None, 0))
                return True
            if out is None:
                out = np.empty((shape[0], shape[1], 3), dtype=a.dtype)
            _check(lib.ss_finish_collapse(h, out.ctypes.data_as(ctypes.c_void_p), out.strides[0]))
            return out
        if self.rank != self.root or not download:
            _check(lib.ss_peer_finish(h, None, 0))
            return True if self.rank == self.root else None
        if out is None:
            out = np.empty((shape[0], shape[1], 3), dtype=a.dtype)
        _check(lib.ss_peer_finish(h, out.ctypes.data_as(ctypes.c_void_p), out.strides[0]))
        return out

    def download(self, shape, out=None):
        """The fused image of the last ``finish(download=False)`` (root only)."""
        a = self.algo
        if out is None:
            out = np.empty((shape[0], shape[1], 3), dtype=a.dtype)
        _check(a._lib.ss_download_output(a._h, out.ctypes.data_as(ctypes.c_void_p), out.strides[0]))
        return out

    def focus_stack_arrays(self, local_frames, n_total):
        local_frames = list(local_frames)
        img0 = local_frames[0]
        a = self.begin(img0, len(local_frames), n_total)
        ring = a._staging(img0.nbytes, 2 * max(a.batch_frames, 1))
        for i, img in enumerate(local_frames):
            a._validate(img, img0.shape[:2], img0.dtype)
            a._push(i, img, ring)
        res = self.finish(len(local_frames), img0.shape[:2])
        return res if self.rank == self.root else None


class ShardedDepthMapStack:
    """``DepthMapStack`` with the frames of one stack sharded over the GPUs of a box, one per rank (SURVEY.md section 8f
    rank 1 + 8e).  Every rank runs the energy, smoothing and blend kernels on its own contiguous run of the frames; the
    three reductions over the stack (depth_map.py:88-115) cross the ranks: the maximum energy as one float through the
    rendezvous, the per-pixel sum of the smoothed energies and the blended pyramid out of the peers' device memory
    (include/shinestack_b200_depthmap.h, ssdm_shard_*).  ``focus_stack_arrays(local_frames, n_total)`` returns the
    fused image on the root and None elsewhere.  Sums are associated rank by rank, so the image equals the single-GPU
    one within float32 rounding (the stacker's 1-LSB contract)."""

    def __init__(self, rank, world, device=0, root=0, rendezvous=None, **kwargs):
        from .depth_map import DepthMapStack      # pylint: disable=import-outside-toplevel
        self.rank, self.world, self.device, self.root = rank, world, device, root
        self.algo = DepthMapStack(device=device, **kwargs)
        self._rdv = rendezvous
        self._connected_handle = None
        self.last_kernel_ms = None

    def _rendezvous(self):
        if self._rdv is None:
            self._rdv = _rdv.TcpRendezvous.from_env() if self.world > 1 else _rdv.Single()
        return self._rdv

    def _connect(self):
        a = self.algo
        if self._connected_handle == a._h.value:
            return
        lib, rdv = a._lib, self._rendezvous()
        count = lib.ssdm_shard_buffer_count(a._h)
        blob = (ctypes.c_uint8 * (64 * count))()
        _check(lib.ssdm_shard_export(a._h, blob))
        blobs = b"".join(rdv.all_gather(bytes(blob)))
        _check(lib.ssdm_shard_connect(a._h, blobs, self.world, self.rank, self.root))
        self._connected_handle = a._h.value

    def finish(self, n_local, shape, dtype, out=None):
        """Everything after the pushes; all ranks must call it."""
        a = self.algo
        lib, h, rdv = a._lib, a._h, self._rendezvous()
        local = ctypes.c_float()
        _check(lib.ssdm_shard_local_max(h, ctypes.byref(local)))
        gmax = max(np.frombuffer(b, np.float32)[0] for b in rdv.all_gather(np.float32(local.value).tobytes()))
        self._connect()
        _check(lib.ssdm_shard_finish_local(h, ctypes.c_float(gmax)))
        _check(lib.ssdm_sync(h))
        rdv.barrier()                       # every partial sum (average) / partial maximum (max) plane is complete
        if a.map_type == "max":
            _check(lib.ssdm_shard_softmax_rel(h))
            _check(lib.ssdm_sync(h))
            rdv.barrier()                   # every partial sum of the exponentials is complete
        _check(lib.ssdm_shard_finish_weights(h))
        for i in range(n_local):
            _check(lib.ssdm_blend_frame(h, i))
        _check(lib.ssdm_sync(h))
        rdv.barrier()                       # every rank's share of the blended pyramid is complete
        _check(lib.ssdm_shard_reduce_band(h))
        _check(lib.ssdm_sync(h))
        rdv.barrier()                       # every band of rows is summed
        res = None
        if self.rank == self.root:
            res = out if out is not None else np.empty((shape[0], shape[1], 3), dtype=dtype)
            _check(lib.ssdm_shard_finish_end(h, res.ctypes.data_as(ctypes.c_void_p), res.strides[0]))
        else:
            _check(lib.ssdm_shard_finish_end(h, None, 0))
        rdv.barrier()                       # the root has read the peers' pyramids: they may be overwritten again
        ms = ctypes.c_float()
        _check(lib.ssdm_kernel_time_ms(h, ctypes.byref(ms)))
        self.last_kernel_ms = ms.value
        return res

    def focus_stack_arrays(self, local_frames, n_total=None):
        a = self.algo
        local_frames = list(local_frames)
        img0 = local_frames[0]
        a._ensure_handle(img0.shape, img0.dtype, len(local_frames))
        a._pitch = img0.shape[1] * 3 * img0.dtype.itemsize
        a._report_steps = False
        shape2, dtype = img0.shape[:2], img0.dtype

        def load_into(i, view):
            from .pyramid import stage_frame      # pylint: disable=import-outside-toplevel
            img = local_frames[i]
            if img is not None and img.shape == img0.shape and img.dtype == dtype:
                stage_frame(img, view)
            return img

        try:
            ring = a._staging(img0.nbytes, 2 * max(a.decode_threads, 2))
            a._stream(len(local_frames), load_into, ring, lambda _i, img: a._validate(img, shape2, dtype))
            return self.finish(len(local_frames), shape2, dtype)
        except BaseException:
            if a._h is not None:
                a._lib.ssdm_reset(a._h)
            raise

    def close(self):
        self.algo.close()
