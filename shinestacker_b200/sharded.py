"""Frames of ONE stack sharded over the GPUs of a box.

Rank ``r`` of ``G`` fuses the contiguous frame block ``shard_range(n, r, G)`` with its own
``PyramidStack`` handle (global frame indices via ``frame_index_base``).  The exchange is
    1. ONE max all-reduce over packed keys ``float_bits(value) << 32 | (0xFFFFFFFF - global_index)``
       for every pyramid level plus the two base-level planes: unsigned/signed max over ranks ==
       "largest energy, ties to the lowest global frame index" == ``np.argmax`` over the whole stack
       (reference pyramid.py:51,104-105);
    2. every rank zeroes the coefficients it did not win; ONE float32 sum-reduce to the root then
       reproduces the winners exactly (x + 0 + ... + 0);
    3. the root blends the base and collapses the pyramid.
torch.distributed (NCCL on GPUs; gloo in the CPU tests) is only plumbing here: the tensors alias the
library's device buffers.
"""
import ctypes

import numpy as np

from . import _native
from .pyramid import PyramidStack, select_base, _check


def shard_range(n_frames, rank, world):
    """Contiguous block of frames for ``rank`` (lower ranks take the remainder): contiguity keeps
    'lowest global index wins' equal to 'lowest rank wins, then lowest local index'."""
    base, rem = divmod(n_frames, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def bind_to_gpu_numa_node(device):
    """Pin this process to the CPUs of the NUMA node its GPU hangs off, so that the pinned staging buffers
    (allocated afterwards) and the upload traffic stay on that socket.  Best effort: returns the node or None."""
    import os
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


def pack_keys(values, indices):
    """numpy restatement of ss_pack_keys_kernel (csrc/ss_lib.cu) for the CPU (gloo) tests."""
    v = np.asarray(values, dtype=np.float32).copy()
    v[v == 0] = 0.0
    bits = v.view(np.uint32).astype(np.int64)
    return (bits << 32) | (0xFFFFFFFF - np.asarray(indices).astype(np.int64))


def unpack_keys(keys):
    keys = np.asarray(keys, dtype=np.int64)
    idx = (0xFFFFFFFF - (keys & 0xFFFFFFFF)).astype(np.int32)
    val = (keys >> 32).astype(np.uint32).view(np.float32)
    return val, idx


def exchange(keys, mask_losers, root=0, group=None):
    """The collective part, backend-agnostic: ``keys`` is an int64 torch tensor (all levels
    concatenated); ``mask_losers()`` is called after the max all-reduce and must return the float32
    coefficient tensor with this rank's lost entries zeroed.  Returns the reduced coefficients
    (valid on ``root``)."""
    import torch.distributed as dist
    dist.all_reduce(keys, op=dist.ReduceOp.MAX, group=group)
    coeffs = mask_losers()
    dist.reduce(coeffs, dst=root, op=dist.ReduceOp.SUM, group=group)
    return coeffs


class _DevArray:
    def __init__(self, ptr, count, typestr):
        self.__cuda_array_interface__ = {"shape": (int(count),), "typestr": typestr,
                                         "data": (int(ptr), False), "version": 3}


class ShardedPyramidStack:
    """One per rank.  ``focus_stack_arrays(local_frames, n_total)`` returns the fused image on the
    root rank and None elsewhere."""

    def __init__(self, rank, world, device=0, group=None, root=0, exchange="p2p", **kwargs):
        """``exchange``: "p2p" = key max all-reduce (NCCL) + root-side gather of the winners' coefficients
        over NVLink peer memory (CUDA IPC); "nccl" = key max all-reduce + masked float32 sum-reduce."""
        self.rank, self.world, self.device, self.group, self.root = rank, world, device, group, root
        self.exchange_mode = exchange
        self.kwargs = kwargs
        self.algo = None
        self.n_total = None

    def _algo(self, n_total):
        start, _ = shard_range(n_total, self.rank, self.world)
        if self.algo is None or self.n_total != n_total:
            self.algo = PyramidStack(device=self.device, frame_index_base=start, keep_depth_maps=False,
                                     **self.kwargs)
            self.n_total = n_total
        return self.algo

    def begin(self, img0, n_local, n_total, resident=False):
        a = self._algo(n_total)
        a._set_metadata(img0)
        a._ensure_handle(img0.shape, img0.dtype, n_local, resident=resident)
        return a

    def finish(self, n_local, shape, out=None):
        """Local base selection, exchange, and (root) collapse.  All ranks must call it."""
        import torch
        a = self.algo
        lib, h = a._lib, a._h
        select_base(lib, h, n_local, a.num_pixel_values, a.float_type)
        if self.world > 1:
            ptr, cnt = ctypes.c_void_p(), ctypes.c_size_t()
            _check(lib.ss_exchange_keys(h, ctypes.byref(ptr), ctypes.byref(cnt)))
            dev = f"cuda:{self.device}"
            keys = torch.as_tensor(_DevArray(ptr.value, cnt.value, "<i8"), device=dev)

            def mask():
                torch.cuda.synchronize(self.device)
                p2, c2 = ctypes.c_void_p(), ctypes.c_size_t()
                _check(lib.ss_exchange_mask_losers(h, ctypes.byref(p2), ctypes.byref(c2)))
                return torch.as_tensor(_DevArray(p2.value, c2.value, "<f4"), device=dev)

            if self.exchange_mode == "p2p":
                import torch.distributed as dist
                dist.all_reduce(keys, op=dist.ReduceOp.MAX, group=self.group)
                mine = torch.empty(64, dtype=torch.uint8)
                _check(lib.ss_exchange_ipc_handle(h, ctypes.c_void_p(mine.data_ptr())))
                allh = torch.empty(64 * self.world, dtype=torch.uint8, device=dev)
                dist.all_gather_into_tensor(allh, mine.to(dev), group=self.group)
                allh = allh.cpu()                       # also waits for the key reduction
                if self.rank == self.root:
                    starts = np.array([shard_range(self.n_total, r, self.world)[0] for r in range(self.world)]
                                      + [self.n_total], dtype=np.int32)
                    _check(lib.ss_exchange_pull(h, ctypes.c_void_p(allh.data_ptr()),
                                                starts.ctypes.data_as(ctypes.c_void_p), self.world, self.rank))
                dist.barrier(group=self.group)          # peers keep their state until the root has gathered it
            else:
                exchange(keys, mask, self.root, self.group)
                torch.cuda.synchronize(self.device)
                if self.rank == self.root:
                    _check(lib.ss_exchange_adopt(h))
        if self.rank != self.root:
            return None
        if out is None:
            out = np.empty((shape[0], shape[1], 3), dtype=a.dtype)
        _check(lib.ss_finish_collapse(h, out.ctypes.data_as(ctypes.c_void_p), out.strides[0]))
        return out

    def focus_stack_arrays(self, local_frames, n_total):
        local_frames = list(local_frames)
        img0 = local_frames[0]
        a = self.begin(img0, len(local_frames), n_total)
        ring = a._staging(img0.nbytes, 2 * max(a.batch_frames, 1))
        for i, img in enumerate(local_frames):
            a._validate(img, img0.shape[:2], img0.dtype)
            a._push(i, img, ring)
        return self.finish(len(local_frames), img0.shape[:2])
