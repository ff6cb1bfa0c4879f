"""Rows sharded over ranks, bounds replicated: the multi-GPU form of the propagation round.

One process per GPU.  Every rank holds a contiguous block of rows (balanced by nonzeros) in its own ``gpulin_t`` handle
and the full bound vectors.  One round is

    local sweep (filter + exact kernels)  ->  ONE all-reduce(MIN) over the int64 candidate keys  ->  dense apply

The key vector stores ``~key(lb)`` and ``key(ub)`` interleaved (csrc/gpulin_device.cuh), so a single MIN merges both
sides of every variable; its two spare entries carry the cutoff verdict.  The collective is the only exchange step of the
path -- the counterpart of the reference's ``syncstore`` min/max merge of global bounds (syncstore.c:921) -- and runs
through ``torch.distributed`` (NCCL over NVLink on GPUs; gloo in the CPU tests of the host logic).
"""
from __future__ import annotations

import numpy as np

FIXPOINT, CUTOFF, ROUNDLIMIT = 0, 1, 2


def partition_rows(rowptr, nparts: int):
    """contiguous row blocks with (almost) equal nonzero counts; returns nparts+1 cut positions"""
    rowptr = np.asarray(rowptr, dtype=np.int64)
    nrows = len(rowptr) - 1
    targets = np.linspace(0, rowptr[-1], nparts + 1)
    cuts = np.searchsorted(rowptr, targets, side="left").astype(np.int64)
    cuts[0], cuts[-1] = 0, nrows
    return np.maximum.accumulate(np.clip(cuts, 0, nrows))


def encode_keys(lb, ub):
    """host mirror of the device key layout: [2j] = ~key(lb_j), [2j+1] = key(ub_j), + 2 spare verdict keys"""
    def d2key(x):
        u = (np.asarray(x, dtype=np.float64) + 0.0).view(np.int64)
        return np.where(u >= 0, u, u ^ np.int64(0x7FFFFFFFFFFFFFFF))
    n = len(lb)
    keys = np.zeros(2 * n + 2, dtype=np.int64)
    keys[0:2 * n:2] = ~d2key(lb)
    keys[1:2 * n:2] = d2key(ub)
    return keys


def decode_keys(keys):
    def key2d(k):
        k = np.asarray(k, dtype=np.int64)
        return np.where(k >= 0, k, k ^ np.int64(0x7FFFFFFFFFFFFFFF)).view(np.float64)
    n = (len(keys) - 2) // 2
    return key2d(~keys[0:2 * n:2]), key2d(keys[1:2 * n:2]), bool(keys[2 * n] < 0)


class CudaEngine:
    """the local row block on one GPU: a thin adapter over LinearPropagator's single-round entry points"""

    def __init__(self, prob, rows, device: int, **numerics):
        import torch
        from .propagator import LinearPropagator
        self.torch = torch
        torch.cuda.set_device(device)
        self.lp = LinearPropagator(prob, device=device, rows=rows, **numerics)
        self.device = device
        # all library work goes to torch's current stream, so the collective is ordered with the kernels; it must be
        # a created stream (the library captures CUDA graphs, which the legacy default stream does not allow)
        if torch.cuda.current_stream(device).cuda_stream == 0:
            torch.cuda.set_stream(torch.cuda.Stream(device))
        self.lp.set_stream(torch.cuda.current_stream(device).cuda_stream)
        ptr, n = self.lp.exchange_buffer()

        class _Holder:
            pass
        h = _Holder()
        h.__cuda_array_interface__ = dict(shape=(n,), typestr="<i8", data=(ptr, False), version=3)
        self.keys = torch.as_tensor(h, device=f"cuda:{device}")
        self._holder = h

    def set_bounds(self, lb, ub):
        self.lp.set_bounds(lb, ub)

    def get_bounds(self):
        return self.lp.get_bounds()

    def round_begin(self):
        self.lp.round_begin()

    def round_sweep(self):
        self.lp.round_sweep()
        return self.keys

    def round_apply(self):
        return self.lp.round_apply(dense=True)

    def close(self):
        self.lp.close()


class PeerPropagator:
    """rows sharded over the GPUs of one node, candidates exchanged through peer memory: after the one-time handle
    exchange every call is a plain (collective) LinearPropagator call -- the round loop runs on the devices"""

    def __init__(self, prob, rank: int, world: int, device: int, group=None, **numerics):
        import torch.distributed as dist
        from .propagator import LinearPropagator
        cuts = partition_rows(prob["rowptr"], world)
        self.lp = LinearPropagator(prob, device=device, rows=(int(cuts[rank]), int(cuts[rank + 1])), **numerics)
        self.rank, self.world = rank, world
        if world > 1:
            blobs = [None] * world
            dist.all_gather_object(blobs, self.lp.peer_handles(), group=group)
            self.lp.peer_connect(rank, blobs)
            dist.barrier(group=group)

    def set_bounds(self, lb, ub):
        self.lp.set_bounds(lb, ub)

    def propagate(self, maxrounds: int = 0):
        return self.lp.propagate(maxrounds)

    def get_bounds(self):
        return self.lp.get_bounds()

    def close(self):
        self.lp.close()


class ShardedPropagator:
    """drives the rounds of one rank; ``engine`` owns the local rows (CudaEngine on GPUs)"""

    def __init__(self, engine, group=None):
        import torch.distributed as dist
        self.dist = dist
        self.engine = engine
        self.group = group

    def propagate(self, maxrounds: int = 0):
        dist = self.dist
        eng = self.engine
        eng.round_begin()
        rounds = 0
        total = 0
        status = ROUNDLIMIT
        while maxrounds <= 0 or rounds < maxrounds:
            keys = eng.round_sweep()
            if dist.is_initialized() and dist.get_world_size(self.group) > 1:
                dist.all_reduce(keys, op=dist.ReduceOp.MIN, group=self.group)
            nchg, cutoff = eng.round_apply()
            rounds += 1
            total += nchg
            if cutoff:
                status = CUTOFF
                break
            if nchg == 0:
                status = FIXPOINT
                break
        return dict(status=status, nrounds=rounds, nchanges=total)
