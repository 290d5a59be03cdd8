"""Multi-GPU plumbing: one process per GPU (torch.distributed, NCCL over NVLink; gloo in CPU tests).

The path shards the ray index (round-robin for balance, or contiguous ranges; rays are independent
through the whole chain) or,
for misalignment sweeps, by chain variants.  There is NO per-ray traffic between ranks; the only
exchanges are the all-reduces of two tiny rows per detector:
    central sums (10 doubles, SUM)  -> every rank places the identical detector (Detector.autoplace)
    moments      (24 doubles)       -> sums SUM; extents MIN/MAX packed into one MAX all-reduce
Both are latency-bound (a few hundred bytes), issued on the compute stream right after the
reduction kernel, with no host synchronisation in between.
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from . import _cabi

_MIN = _cabi.MOMENT_MIN
_MAX = _cabi.MOMENT_MAX
_SUM = list(range(0, 14))


def shard_range(n, rank, world):
    """(first, count) of rank's contiguous slice of n items: [rank*n//world, (rank+1)*n//world)."""
    lo = (rank * n) // world
    hi = ((rank + 1) * n) // world
    return lo, hi - lo


def shard_strided(n, rank, world):
    """(first, count, stride) of rank's round-robin share of n items: rank, rank + world, ...  Used for
    ray bundles: an aperture typically blocks a contiguous range of Vogel-spiral indices, so
    contiguous shards would leave some ranks with only dead rays."""
    return rank, max(0, (n - rank + world - 1) // world), world


def bind_to_gpu_numa(device_index):
    """Pin the calling process to the CPUs that are local to GPU `device_index` (NVML's ideal CPU affinity), so that
    pinned host staging allocated afterwards is first-touched on the GPU's own NUMA node: on a two-socket 8-GPU box
    eight ranks uploading ray columns at once otherwise share one socket's memory controllers.  Returns the number
    of CPUs bound to, or 0 when NVML is unavailable (nothing changes then)."""
    try:
        import os
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(int(device_index))
        n_cpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (n_cpu + 63) // 64)
        cpus = [64 * w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1]
        cpus = [c for c in cpus if c < n_cpu]
        if cpus:
            os.sched_setaffinity(0, cpus)
        return len(cpus)
    except Exception:
        return 0


def is_distributed(group=None):
    return dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1


def all_reduce_central(central, group=None):
    """Sum the central-ray rows (n_variants x 10) over the ranks, in place."""
    if is_distributed(group):
        dist.all_reduce(central, op=dist.ReduceOp.SUM, group=group)
    return central


def all_reduce_moments(moments, group=None, gather_buffer=None):
    """Merge the moments rows (n_variants x 24) over the ranks, in place.

    On CUDA: ONE all-gather of the rows followed by the library's merge kernel (art_moments_merge:
    sums in rank order, extents by min / max) -- a single collective that a CUDA graph can capture.
    On CPU tensors (gloo tests): additive entries summed, minima / maxima in one max-all-reduce."""
    if not is_distributed(group):
        return moments
    if moments.is_cuda:
        import ctypes as C
        world = dist.get_world_size(group)
        if gather_buffer is None:
            gather_buffer = torch.empty((world,) + tuple(moments.shape), dtype=moments.dtype, device=moments.device)
        dist.all_gather_into_tensor(gather_buffer, moments.contiguous(), group=group)
        _cabi.check(_cabi.lib().art_moments_merge(C.c_void_p(gather_buffer.data_ptr()), world, moments.shape[0],
                                                  C.c_void_p(moments.data_ptr()),
                                                  C.c_void_p(torch.cuda.current_stream().cuda_stream)))
        return moments
    sums = moments[:, _SUM].contiguous()
    ext = torch.cat([-moments[:, _MIN], moments[:, _MAX]], dim=1).contiguous()
    dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=group)
    dist.all_reduce(ext, op=dist.ReduceOp.MAX, group=group)
    moments[:, _SUM] = sums
    moments[:, _MIN] = -ext[:, : len(_MIN)]
    moments[:, _MAX] = ext[:, len(_MIN):]
    return moments


def all_reduce_histogram(hist, group=None):
    """Sum the int64 histograms of every rank's shard (exact: integer bins).  Every rank must have binned
    against the same detector and the same merged moments row (`all_reduce_moments` first)."""
    if is_distributed(group):
        dist.all_reduce(hist, op=dist.ReduceOp.SUM, group=group)
    return hist


class PeerExchange:
    """The exchange of the central sums / moments rows done inside one kernel over peer memory
    (art_peer_exchange) instead of NCCL: one node, every GPU reachable over NVLink / NVSwitch.  The exchange
    buffer is torch symmetric memory (one allocation per rank, mapped into every rank's address space).

        peer = PeerExchange.create(device)        # None when symmetric memory is not available -> use NCCL
        peer.all_reduce_central(central, distance, det)   # sum over ranks + Detector.autoplace, one launch
        peer.all_reduce_moments(moments)                  # merge over ranks, one launch
    Every rank must issue the same sequence of calls.  No host synchronisation; CUDA-graph capturable."""

    def __init__(self, buffer, handle, rank, world):
        import ctypes as C
        self._buffer = buffer      # keeps the symmetric allocation alive
        self._handle = handle
        self.rank, self.world = rank, world
        self._ptrs = (C.c_uint64 * world)(*[int(p) for p in handle.buffer_ptrs])

    @classmethod
    def create(cls, device, group=None):
        if not is_distributed(group):
            return None
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        if world > _cabi.PEER_MAX_RANKS:
            return None
        try:
            import torch.distributed._symmetric_memory as symm_mem
            n = _cabi.peer_buffer_bytes(world) // 8
            buf = symm_mem.empty(n, dtype=torch.float64, device=device)
            handle = symm_mem.rendezvous(buf, group if group is not None else dist.group.WORLD)
            buf.zero_()
            torch.cuda.synchronize(device)
            ok = torch.ones(1, device=device)
        except Exception as exc:  # no symmetric memory on this system: the caller keeps NCCL
            import warnings
            warnings.warn(f"peer-memory exchange unavailable ({exc}); using NCCL collectives")
            ok = torch.zeros(1, device=device)
            buf = handle = None
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)  # all ranks or none; also: every buffer is zeroed
        if float(ok) < 1.0:
            return None
        return cls(buf, handle, rank, world)

    def _call(self, kind, rows, distance, det, chain=None):
        import ctypes as C
        if not rows.is_contiguous() or rows.dtype != torch.float64:
            raise ValueError("rows must be a contiguous float64 tensor")
        nv = rows.shape[0]
        dp = C.c_void_p(det.data_ptr()) if det is not None else None
        st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        if chain is not None:
            if nv != 1:
                raise ValueError("folding inside the exchange is for one variant")
            _cabi.check(_cabi.lib().art_peer_exchange_fold(chain._handle, self._ptrs, self.rank, self.world, kind,
                                                           C.c_void_p(rows.data_ptr()), float(distance), dp, st))
        else:
            _cabi.check(_cabi.lib().art_peer_exchange(self._ptrs, self.rank, self.world, kind, nv,
                                                      C.c_void_p(rows.data_ptr()), float(distance), dp, st))
        return rows

    def all_reduce_central(self, central, distance=0.0, det=None, chain=None):
        """Sum the central rows (n_variants x 10) over the ranks in place; with `det` (n_variants x 23) also
        place every variant's detector at `distance` (Detector.autoplace) in the same kernel.
        chain: the DeviceChain whose last call was trace(..., fold=False): its per-block rows are folded into
        `central` inside the exchange kernel (art_peer_exchange_fold) -- no fold launch of its own."""
        return self._call(0, central, distance, det, chain)

    def all_reduce_moments(self, moments, chain=None):
        """Merge the moments rows (n_variants x 24) over the ranks in place (sums / minima / maxima).
        chain: the DeviceChain whose last call was moments(..., fold=False)."""
        return self._call(1, moments, 0.0, None, chain)

    def status(self):
        """0, or the epoch of an exchange in which a peer did not arrive (synchronises the stream)."""
        import ctypes as C
        out = C.c_uint64(0)
        _cabi.check(_cabi.lib().art_peer_status(self._ptrs, self.rank, self.world, C.byref(out),
                                                C.c_void_p(torch.cuda.current_stream().cuda_stream)))
        return int(out.value)


    def stats(self, reset=False):
        """{"exchanges", "wait_us", "kernel_us"}: per-exchange averages on THIS rank since the last reset -- time
        spent waiting for the peers' flags (rank skew + NVLink round trip) and inside the exchange kernel
        (art_peer_stats; synchronises the stream)."""
        import ctypes as C
        out = (C.c_uint64 * _cabi.PEER_STATS)()
        _cabi.check(_cabi.lib().art_peer_stats(self._ptrs, self.rank, self.world, out, 1 if reset else 0,
                                               C.c_void_p(torch.cuda.current_stream().cuda_stream)))
        n = int(out[0])
        return {"exchanges": n, "wait_us": (out[1] / n / 1e3) if n else 0.0, "kernel_us": (out[2] / n / 1e3) if n else 0.0}


def merge_moments(rows):
    """Host-side merge of moments rows from several shards (sequence of (24,) arrays / tensors)."""
    rows = torch.stack([torch.as_tensor(r, dtype=torch.float64) for r in rows])
    out = torch.zeros(_cabi.MOMENTS_LEN, dtype=torch.float64)
    out[_SUM] = rows[:, _SUM].sum(dim=0)
    out[_MIN] = rows[:, _MIN].min(dim=0).values
    out[_MAX] = rows[:, _MAX].max(dim=0).values
    return out
