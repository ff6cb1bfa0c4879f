"""Several GPUs of one node: the multi-GPU forms of the propagation round.

``PeerPropagator`` / ``GroupPropagator`` (the product path): matrix and bounds replicated, the work of the dense rounds
shared, one packed exchange per dense round through peer memory inside the round's kernels.

``ShardedPropagator`` (kept for comparison and for hosts without peer access): every rank holds a contiguous block of rows
(balanced by nonzeros) in its own ``gpulin_t`` handle and the full bound vectors.  One round is

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


def pack_changes(keys_before, keys_after):
    """host mirror of peer_push_kernel: the columns whose keys a rank's own candidates moved, as (columns, key pairs)"""
    n = (len(keys_after) - 2) // 2
    kb = np.asarray(keys_before[:2 * n]).reshape(n, 2)
    ka = np.asarray(keys_after[:2 * n]).reshape(n, 2)
    cols = np.flatnonzero((ka != kb).any(axis=1)).astype(np.int32)
    return cols, ka[cols].copy()


def merge_changes(keys, cols, pairs):
    """host mirror of peer_merge_kernel: elementwise MIN of the received key pairs into the local keys (in place);
    returns the columns touched"""
    n = (len(keys) - 2) // 2
    view = keys[:2 * n].reshape(n, 2)
    np.minimum.at(view, cols, pairs)
    return cols


class ReplicatedPropagator:
    """host mirror of the product's multi-GPU round (csrc/gpulin_kernels.cuh, "dense rounds sharded over the GPUs of a
    node"): bounds replicated, every rank sweeps its share of the rows, packs the columns it touched, the ranks
    all-gather the packets (peer-memory stores on the GPUs) and merge them with MIN; every rank then applies the same
    changes.  ``engine`` needs set_keys / get_keys on top of the round interface (tests: the CPU oracle over gloo)"""

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
        world = dist.get_world_size(self.group) if dist.is_initialized() else 1
        while maxrounds <= 0 or rounds < maxrounds:
            before = eng.get_keys().copy()
            keys = np.asarray(eng.round_sweep()).copy()
            cols, pairs = pack_changes(before, keys)
            packet = (cols, pairs, bool(keys[-2] < 0))
            packets = [packet]
            if world > 1:
                packets = [None] * world
                dist.all_gather_object(packets, packet, group=self.group)
            for c, pr, cut in packets:
                merge_changes(keys, c, pr)
                if cut:
                    keys[-2] = -1
            eng.set_keys(keys)
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
    """several GPUs of one node, one process per GPU: every rank holds the whole problem; dense rounds are shared (a rank
    sweeps its share of the rows, the touched columns travel once per dense round through peer memory), small rounds run
    redundantly.  After the one-time handle exchange every call is a plain (collective) LinearPropagator call -- the
    round loop runs on the devices"""

    def __init__(self, prob, rank: int, world: int, device: int, group=None, **numerics):
        import torch.distributed as dist
        from .propagator import LinearPropagator
        self.lp = LinearPropagator(prob, device=device, **numerics)
        self.rank, self.world = rank, world
        if world > 1:
            blobs = [None] * world
            dist.all_gather_object(blobs, self.lp.peer_handles(world), group=group)
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


class GroupPropagator:
    """the same in ONE process: n handles on n devices (``gpulin_group_connect``), driven from one host thread -- what the
    SCIP plugin does with ``propagating/gpulinear/ndevices`` (SCIP is single threaded)"""

    def __init__(self, prob, devices, **numerics):
        import ctypes
        from .propagator import LinearPropagator, _check
        self.lps = [LinearPropagator(prob, device=d, **numerics) for d in devices]
        arr = (ctypes.c_void_p * len(self.lps))(*[lp._h for lp in self.lps])
        _check(self.lps[0]._lib.gpulin_group_connect(arr, len(self.lps)))

    def set_bounds(self, lb, ub):
        for lp in self.lps:
            lp.set_bounds(lb, ub)

    def update_bounds(self, idx, lb, ub):
        for lp in self.lps:
            lp.update_bounds(idx, lb, ub)

    def propagate(self, maxrounds: int = 0):
        for lp in self.lps:
            lp.propagate_async(maxrounds)
        res = [lp.propagate_wait() for lp in self.lps]
        for r in res[1:]:
            assert (r["status"], r["nrounds"], r["nchanges"]) == (res[0]["status"], res[0]["nrounds"], res[0]["nchanges"])
        return res[0]

    def get_bounds(self, rank: int = 0):
        return self.lps[rank].get_bounds()

    def close(self):
        for lp in self.lps:
            lp.close()


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
