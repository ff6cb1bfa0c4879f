"""ctypes binding of libgpulin.so (include/gpulin.h) -- the harness side of the C ABI.

The product is the CUDA library plus the SCIP plugin ``plugin/prop_gpulinear.c``; this module only lets tests,
``bench.py`` and the multi-GPU driver call the same C entry points from Python.  Nothing here computes bounds:
if the library or a CUDA device is missing every call raises -- there is no CPU fallback.
"""
from __future__ import annotations

import ctypes
import os

import numpy as np

from . import build as _build

FIXPOINT, CUTOFF, ROUNDLIMIT = 0, 1, 2
STATUS_NAMES = {0: "fixpoint", 1: "cutoff", 2: "roundlimit"}


class GpulinError(RuntimeError):
    pass


class Numerics(ctypes.Structure):
    _fields_ = [(n, ctypes.c_double) for n in
                ("infinity", "epsilon", "sumepsilon", "feastol", "boundstreps", "hugeval", "maxeasyactivitydelta")]


class Result(ctypes.Structure):
    _fields_ = [("status", ctypes.c_int32), ("nrounds", ctypes.c_int32), ("nchanges", ctypes.c_int64),
                ("nnz_processed", ctypes.c_int64), ("device_ms", ctypes.c_double)]


CHANGE_DTYPE = np.dtype([("var", np.int32), ("round", np.int32), ("newbound", np.float64), ("is_upper", np.int32),
                         ("reserved", np.int32)])

_lib = None

_P = ctypes.c_void_p
_SIGNATURES = {
    "gpulin_default_numerics": (None, [ctypes.POINTER(Numerics)]),
    "gpulin_last_error": (ctypes.c_char_p, []),
    "gpulin_device_count": (ctypes.c_int, []),
    "gpulin_create": (ctypes.c_int, [ctypes.c_int, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64, _P, _P, _P, _P, _P,
                                     _P, ctypes.POINTER(Numerics), ctypes.POINTER(_P)]),
    "gpulin_destroy": (None, [_P]),
    "gpulin_set_bounds": (ctypes.c_int, [_P, _P, _P]),
    "gpulin_set_bounds_device": (ctypes.c_int, [_P, _P, _P]),
    "gpulin_update_bounds": (ctypes.c_int, [_P, ctypes.c_int64, _P, _P, _P]),
    "gpulin_propagate": (ctypes.c_int, [_P, ctypes.c_int, ctypes.POINTER(Result)]),
    "gpulin_propagate_async": (ctypes.c_int, [_P, ctypes.c_int]),
    "gpulin_propagate_wait": (ctypes.c_int, [_P, ctypes.POINTER(Result)]),
    "gpulin_clone": (ctypes.c_int, [_P, ctypes.POINTER(_P)]),
    "gpulin_reset_from": (ctypes.c_int, [_P, _P]),
    "gpulin_probe_batch": (ctypes.c_int, [_P, ctypes.c_int, ctypes.c_int64, _P, _P, _P, ctypes.c_int, _P, _P, _P]),
    "gpulin_probe_batch_changes": (ctypes.c_int, [_P, ctypes.c_int, ctypes.c_int64, _P, _P, _P, ctypes.c_int, _P, _P, _P, _P, _P,
                                                  ctypes.c_int64, ctypes.POINTER(ctypes.c_int64)]),
    "gpulin_get_bounds": (ctypes.c_int, [_P, _P, _P]),
    "gpulin_get_bounds_device": (ctypes.c_int, [_P, _P, _P]),
    "gpulin_set_change_log": (ctypes.c_int, [_P, ctypes.c_int64]),
    "gpulin_get_changes": (ctypes.c_int, [_P, _P, ctypes.c_int64, ctypes.POINTER(ctypes.c_int64)]),
    "gpulin_get_redundant_rows": (ctypes.c_int, [_P, _P, ctypes.c_int64, ctypes.POINTER(ctypes.c_int64)]),
    "gpulin_get_round_stats": (ctypes.c_int, [_P, _P, _P, _P, ctypes.c_int32, ctypes.POINTER(ctypes.c_int32)]),
    "gpulin_get_layout": (ctypes.c_int, [_P, _P, ctypes.c_int32]),
    "gpulin_get_call_stats": (ctypes.c_int, [_P, _P, ctypes.c_int32]),
    "gpulin_algorithmic_bytes": (ctypes.c_int, [_P, ctypes.POINTER(ctypes.c_int64)]),
    "gpulin_profile_round": (ctypes.c_int, [_P, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double),
                                            ctypes.POINTER(ctypes.c_double)]),
    "gpulin_set_stream": (ctypes.c_int, [_P, _P]),
    "gpulin_sync": (ctypes.c_int, [_P]),
    "gpulin_exchange_buffer": (ctypes.c_int, [_P, ctypes.POINTER(_P), ctypes.POINTER(ctypes.c_int64)]),
    "gpulin_peer_handles": (ctypes.c_int, [_P, ctypes.c_int, _P, ctypes.POINTER(ctypes.c_int64)]),
    "gpulin_group_connect": (ctypes.c_int, [_P, ctypes.c_int]),
    "gpulin_set_reference_bounds": (ctypes.c_int, [_P, _P, _P]),
    "gpulin_set_bounds_packed": (ctypes.c_int, [_P, _P, ctypes.c_int64, _P, _P, _P]),
    "gpulin_get_changes_packed": (ctypes.c_int, [_P, _P, ctypes.c_int64, ctypes.POINTER(ctypes.c_int64)]),
    "gpulin_get_changes_compact": (ctypes.c_int, [_P, _P, ctypes.c_int64, ctypes.POINTER(ctypes.c_int64), _P, ctypes.c_int64,
                                                  ctypes.POINTER(ctypes.c_int64)]),
    "gpulin_set_rangedrow": (ctypes.c_int, [_P, ctypes.c_int, _P, _P, _P]),
    "gpulin_get_trace": (ctypes.c_int, [_P, _P, _P, ctypes.c_int32, ctypes.POINTER(ctypes.c_int32)]),
    "gpulin_get_exchange_stats": (ctypes.c_int, [_P, _P, _P, ctypes.c_int32, ctypes.POINTER(ctypes.c_int32)]),
    "gpulin_peer_connect": (ctypes.c_int, [_P, ctypes.c_int, ctypes.c_int, _P]),
    "gpulin_get_keys": (ctypes.c_int, [_P, _P]),
    "gpulin_set_keys": (ctypes.c_int, [_P, _P]),
    "gpulin_mark_all": (ctypes.c_int, [_P]),
    "gpulin_round_begin": (ctypes.c_int, [_P]),
    "gpulin_round_sweep": (ctypes.c_int, [_P]),
    "gpulin_round_apply": (ctypes.c_int, [_P, ctypes.c_int, ctypes.POINTER(ctypes.c_int64),
                                          ctypes.POINTER(ctypes.c_int32)]),
}


def library_path() -> str:
    return _build.LIB


def exported_symbols():
    """names declared in include/gpulin.h (kept in one place for the symbol test)"""
    return sorted(_SIGNATURES)


def load_library():
    """dlopen libgpulin.so (never builds implicitly on a GPU box: the in-tree library must be there)"""
    global _lib
    if _lib is None:
        path = library_path()
        if not os.path.exists(path):
            raise GpulinError(f"{path} is missing: run `python -m scip_b200.build` (no CPU fallback exists)")
        lib = ctypes.CDLL(path)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def _check(rc: int):
    if rc != 0:
        raise GpulinError(f"libgpulin error {rc}: {load_library().gpulin_last_error().decode()}")


def default_numerics(**kw) -> Numerics:
    num = Numerics()
    load_library().gpulin_default_numerics(ctypes.byref(num))
    for k, v in kw.items():
        setattr(num, k, float(v))
    return num


class LinearPropagator:
    """device-resident copy of the linear rows of one problem + its bound vectors (one gpulin_t handle)

    ``prob``: dict with rowptr[int64], colidx[int32], vals[f64], lhs/rhs[f64 per row], vartype[u8 per column] and
    optionally lb/ub.  ``rows=(begin,end)`` keeps only that row block (multi-GPU row partition)."""

    def __init__(self, prob, device: int = 0, rows=None, **numerics):
        lib = load_library()
        rowptr = np.ascontiguousarray(prob["rowptr"], dtype=np.int64)
        colidx = np.ascontiguousarray(prob["colidx"], dtype=np.int32)
        vals = np.ascontiguousarray(prob["vals"], dtype=np.float64)
        lhs = np.ascontiguousarray(prob["lhs"], dtype=np.float64)
        rhs = np.ascontiguousarray(prob["rhs"], dtype=np.float64)
        vartype = np.ascontiguousarray(prob["vartype"], dtype=np.uint8)
        if rows is not None:
            b, e = int(rows[0]), int(rows[1])
            k0, k1 = int(rowptr[b]), int(rowptr[e])
            rowptr = np.ascontiguousarray(rowptr[b:e + 1] - k0)
            colidx = np.ascontiguousarray(colidx[k0:k1])
            vals = np.ascontiguousarray(vals[k0:k1])
            lhs = np.ascontiguousarray(lhs[b:e])
            rhs = np.ascontiguousarray(rhs[b:e])
        self.nrows, self.ncols, self.nnz = len(lhs), len(vartype), len(vals)
        self.numerics = default_numerics(**numerics)
        self._h = _P()
        _check(lib.gpulin_create(int(device), self.nrows, self.ncols, self.nnz, rowptr.ctypes.data, colidx.ctypes.data,
                                 vals.ctypes.data, lhs.ctypes.data, rhs.ctypes.data, vartype.ctypes.data,
                                 ctypes.byref(self.numerics), ctypes.byref(self._h)))
        self._lib = lib
        self.device = int(device)
        if "lb" in prob and "ub" in prob:
            self.set_bounds(prob["lb"], prob["ub"])

    # -- lifetime ----------------------------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            self._lib.gpulin_destroy(self._h)
            self._h = _P()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # -- bounds ------------------------------------------------------------------------------------------------
    def set_bounds(self, lb, ub):
        lb = np.ascontiguousarray(lb, dtype=np.float64)
        ub = np.ascontiguousarray(ub, dtype=np.float64)
        assert lb.shape == (self.ncols,) and ub.shape == (self.ncols,)
        _check(self._lib.gpulin_set_bounds(self._h, lb.ctypes.data, ub.ctypes.data))
        # the host arrays may be pageable and temporary: the copy must be done before they go away
        _check(self._lib.gpulin_sync(self._h))

    def set_bounds_ptr(self, lb_ptr: int, ub_ptr: int, on_device: bool):
        """raw pointers: pinned host memory (asynchronous copy) or device memory"""
        fn = self._lib.gpulin_set_bounds_device if on_device else self._lib.gpulin_set_bounds
        _check(fn(self._h, lb_ptr, ub_ptr))

    def get_bounds_ptr(self, lb_ptr: int, ub_ptr: int, on_device: bool):
        fn = self._lib.gpulin_get_bounds_device if on_device else self._lib.gpulin_get_bounds
        _check(fn(self._h, lb_ptr, ub_ptr))

    def set_reference_bounds(self, lb, ub):
        lb = np.ascontiguousarray(lb, dtype=np.float64)
        ub = np.ascontiguousarray(ub, dtype=np.float64)
        assert lb.shape == (self.ncols,) and ub.shape == (self.ncols,)
        _check(self._lib.gpulin_set_reference_bounds(self._h, lb.ctypes.data, ub.ctypes.data))
        self._ref = (lb + 0.0, ub + 0.0)

    def pack_bounds(self, lb, ub):
        """(codes[uint32], idx, lb, ub of the explicit entries) of bounds against the reference bounds"""
        rl, ru = self._ref
        lb = np.asarray(lb, dtype=np.float64) + 0.0
        ub = np.asarray(ub, dtype=np.float64) + 0.0
        code = np.full(self.ncols, 3, dtype=np.uint32)
        code[(lb == rl) & (ub == rl)] = 1
        code[(lb == ru) & (ub == ru)] = 2
        code[(lb == rl) & (ub == ru)] = 0
        idx = np.flatnonzero(code == 3).astype(np.int32)
        pad = (-self.ncols) % 16
        c = np.concatenate([code, np.zeros(pad, dtype=np.uint32)]).reshape(-1, 16)
        words = (c << (2 * np.arange(16, dtype=np.uint32))).sum(axis=1, dtype=np.uint64).astype(np.uint32)
        return np.ascontiguousarray(words), idx, np.ascontiguousarray(lb[idx]), np.ascontiguousarray(ub[idx])

    def set_bounds_packed(self, words, idx, lb, ub):
        _check(self._lib.gpulin_set_bounds_packed(self._h, words.ctypes.data, len(idx), idx.ctypes.data if len(idx) else None,
                                                  lb.ctypes.data if len(idx) else None, ub.ctypes.data if len(idx) else None))
        _check(self._lib.gpulin_sync(self._h))

    def set_bounds_packed_ptr(self, words_ptr: int, nexplicit: int = 0, idx_ptr=None, lb_ptr=None, ub_ptr=None):
        """raw pointers (pinned host memory: asynchronous copies)"""
        _check(self._lib.gpulin_set_bounds_packed(self._h, words_ptr, int(nexplicit), idx_ptr, lb_ptr, ub_ptr))

    def changes_packed(self, maxn: int):
        """the change log as (var, is_upper, newbound) arrays + number of entries produced"""
        buf = np.empty(3 * max(maxn, 1), dtype=np.uint32)
        n = ctypes.c_int64(0)
        _check(self._lib.gpulin_get_changes_packed(self._h, buf.ctypes.data, maxn, ctypes.byref(n)))
        m = min(n.value, maxn)
        rec = buf[:3 * m].reshape(m, 3)
        var = (rec[:, 0] & 0x7FFFFFFF).astype(np.int32)
        upper = (rec[:, 0] >> 31).astype(np.int32)
        val = (rec[:, 1].astype(np.uint64) | (rec[:, 2].astype(np.uint64) << np.uint64(32))).view(np.float64)
        return var, upper, val, n.value

    def changes_compact(self, maxn: int):
        """the change log through gpulin_get_changes_compact (one word per entry + a side list for bounds other than 0 / 1),
        decoded to (var, is_upper, newbound) arrays + number of entries produced"""
        words = np.empty(max(maxn, 1), dtype=np.uint32)
        side = np.empty(3 * max(maxn, 1), dtype=np.uint32)
        n = ctypes.c_int64(0)
        nx = ctypes.c_int64(0)
        _check(self._lib.gpulin_get_changes_compact(self._h, words.ctypes.data, maxn, ctypes.byref(n), side.ctypes.data, maxn,
                                                    ctypes.byref(nx)))
        m = min(n.value, maxn)
        w = words[:m]
        var = (w & np.uint32(0x1FFFFFFF)).astype(np.int32)
        upper = (w >> np.uint32(31)).astype(np.int32)
        code = (w >> np.uint32(29)) & np.uint32(3)
        val = np.where(code == 1, 1.0, 0.0)
        rec = side[:3 * min(nx.value, maxn)].reshape(-1, 3)
        pos = rec[:, 0].astype(np.int64)
        assert int((code == 2).sum()) == len(pos)
        val[pos] = (rec[:, 1].astype(np.uint64) | (rec[:, 2].astype(np.uint64) << np.uint64(32))).view(np.float64)
        return var, upper, val, n.value

    def changes_compact_ptr(self, out_ptr: int, maxn: int, side_ptr: int, maxx: int):
        n = ctypes.c_int64(0)
        nx = ctypes.c_int64(0)
        _check(self._lib.gpulin_get_changes_compact(self._h, out_ptr, maxn, ctypes.byref(n), side_ptr, maxx, ctypes.byref(nx)))
        return n.value, nx.value

    def changes_packed_ptr(self, out_ptr: int, maxn: int) -> int:
        n = ctypes.c_int64(0)
        _check(self._lib.gpulin_get_changes_packed(self._h, out_ptr, maxn, ctypes.byref(n)))
        return n.value

    def set_rangedrow(self, enable: bool, glb=None, gub=None, tie=None):
        """ranged-row (gcd) propagation on / off; glb / gub: the global bounds the reference sorts its rows by, tie: the
        SCIPvarGetProbindex of every column (default: the column index)"""
        if not enable:
            _check(self._lib.gpulin_set_rangedrow(self._h, 0, None, None, None))
            return
        glb = np.ascontiguousarray(glb, dtype=np.float64)
        gub = np.ascontiguousarray(gub, dtype=np.float64)
        tb = None if tie is None else np.ascontiguousarray(tie, dtype=np.int32)
        _check(self._lib.gpulin_set_rangedrow(self._h, 1, glb.ctypes.data, gub.ctypes.data, None if tb is None else tb.ctypes.data))

    def update_bounds(self, idx, lb, ub):
        idx = np.ascontiguousarray(idx, dtype=np.int32)
        lb = np.ascontiguousarray(lb, dtype=np.float64)
        ub = np.ascontiguousarray(ub, dtype=np.float64)
        _check(self._lib.gpulin_update_bounds(self._h, len(idx), idx.ctypes.data, lb.ctypes.data, ub.ctypes.data))
        _check(self._lib.gpulin_sync(self._h))

    def get_bounds(self):
        lb = np.empty(self.ncols, dtype=np.float64)
        ub = np.empty(self.ncols, dtype=np.float64)
        _check(self._lib.gpulin_get_bounds(self._h, lb.ctypes.data, ub.ctypes.data))
        return lb, ub

    # -- propagation -------------------------------------------------------------------------------------------
    def propagate(self, maxrounds: int = 0) -> dict:
        res = Result()
        _check(self._lib.gpulin_propagate(self._h, int(maxrounds), ctypes.byref(res)))
        return dict(status=res.status, nrounds=res.nrounds, nchanges=res.nchanges, nnz_processed=res.nnz_processed,
                    device_ms=res.device_ms)

    def propagate_async(self, maxrounds: int = 0):
        _check(self._lib.gpulin_propagate_async(self._h, int(maxrounds)))

    def propagate_wait(self) -> dict:
        res = Result()
        _check(self._lib.gpulin_propagate_wait(self._h, ctypes.byref(res)))
        return dict(status=res.status, nrounds=res.nrounds, nchanges=res.nchanges, nnz_processed=res.nnz_processed,
                    device_ms=res.device_ms)

    def clone(self) -> "LinearPropagator":
        """another set of bound vectors on the same device matrix (probing)"""
        other = object.__new__(LinearPropagator)
        other._lib = self._lib
        other._h = _P()
        _check(self._lib.gpulin_clone(self._h, ctypes.byref(other._h)))
        other.nrows, other.ncols, other.nnz = self.nrows, self.ncols, self.nnz
        other.numerics, other.device = self.numerics, self.device
        return other

    def reset_from(self, base: "LinearPropagator"):
        _check(self._lib.gpulin_reset_from(self._h, base._h))

    def update_bounds_nosync(self, idx, lb, ub):
        """like update_bounds, but the (numpy, contiguous) arrays must stay alive until the stream has consumed them"""
        _check(self._lib.gpulin_update_bounds(self._h, len(idx), idx.ctypes.data, lb.ctypes.data, ub.ctypes.data))

    def probe_batch(self, var, lb, ub, nworkers: int = 32, maxrounds: int = 0) -> dict:
        """probe i: column var[i] set to [lb[i], ub[i]] on top of this handle's bounds, propagated to its fixpoint"""
        var = np.ascontiguousarray(var, dtype=np.int32)
        lb = np.ascontiguousarray(lb, dtype=np.float64)
        ub = np.ascontiguousarray(ub, dtype=np.float64)
        n = len(var)
        status = np.zeros(n, dtype=np.int32)
        nrounds = np.zeros(n, dtype=np.int32)
        nchanges = np.zeros(n, dtype=np.int64)
        _check(self._lib.gpulin_probe_batch(self._h, int(nworkers), n, var.ctypes.data, lb.ctypes.data, ub.ctypes.data,
                                            int(maxrounds), status.ctypes.data, nrounds.ctypes.data, nchanges.ctypes.data))
        return dict(status=status, nrounds=nrounds, nchanges=nchanges)

    def probe_batch_changes(self, var, lb, ub, nworkers: int = 32, maxrounds: int = 0, maxchg: int = 1 << 20) -> dict:
        """like probe_batch, plus what every probe implied: ``changes[chgbeg[i]:chgbeg[i+1]]`` are probe i's accepted
        bound changes in round order (the sparse form of SCIPapplyProbingVar's proplbs / propubs)"""
        var = np.ascontiguousarray(var, dtype=np.int32)
        lb = np.ascontiguousarray(lb, dtype=np.float64)
        ub = np.ascontiguousarray(ub, dtype=np.float64)
        n = len(var)
        status = np.zeros(n, dtype=np.int32)
        nrounds = np.zeros(n, dtype=np.int32)
        nchanges = np.zeros(n, dtype=np.int64)
        chgbeg = np.zeros(n + 1, dtype=np.int64)
        chg = np.zeros(max(maxchg, 1), dtype=CHANGE_DTYPE)
        nchg = ctypes.c_int64(0)
        _check(self._lib.gpulin_probe_batch_changes(self._h, int(nworkers), n, var.ctypes.data, lb.ctypes.data, ub.ctypes.data,
                                                    int(maxrounds), status.ctypes.data, nrounds.ctypes.data,
                                                    nchanges.ctypes.data, chgbeg.ctypes.data, chg.ctypes.data, int(maxchg),
                                                    ctypes.byref(nchg)))
        return dict(status=status, nrounds=nrounds, nchanges=nchanges, chgbeg=chgbeg, changes=chg[:nchg.value])

    def round_stats(self, maxn: int = 1024):
        ms = np.zeros(maxn, dtype=np.float64)
        nnz = np.zeros(maxn, dtype=np.int64)
        nchg = np.zeros(maxn, dtype=np.int64)
        n = ctypes.c_int32(0)
        _check(self._lib.gpulin_get_round_stats(self._h, ms.ctypes.data, nnz.ctypes.data, nchg.ctypes.data, maxn,
                                                ctypes.byref(n)))
        return ms[:n.value], nnz[:n.value], nchg[:n.value]

    TRACE_NAMES = {1: "begin", 2: "sweep_sell", 3: "sweep_stream", 4: "sweep_long", 5: "exact", 6: "push", 7: "push_end",
                   8: "merge", 9: "merge_ready", 10: "apply", 11: "sparse", 12: "sparse_end"}

    def trace(self, maxn: int = 256):
        """[(kernel name, microseconds since the device-side start of the last call)] in time order"""
        ids = np.zeros(maxn, dtype=np.int32)
        us = np.zeros(maxn, dtype=np.float64)
        n = ctypes.c_int32(0)
        _check(self._lib.gpulin_get_trace(self._h, ids.ctypes.data, us.ctypes.data, maxn, ctypes.byref(n)))
        return [(self.TRACE_NAMES.get(int(i), str(int(i))), float(t)) for i, t in zip(ids[:n.value], us[:n.value])]

    def exchange_stats(self, maxn: int = 1024):
        """several GPUs: per round (ms from the round's start to its exchange or -1, ms the merge waited for the peers)"""
        before = np.zeros(maxn)
        wait = np.zeros(maxn)
        n = ctypes.c_int32(0)
        _check(self._lib.gpulin_get_exchange_stats(self._h, before.ctypes.data, wait.ctypes.data, maxn, ctypes.byref(n)))
        return before[:n.value], wait[:n.value]

    def set_change_log(self, capacity: int):
        _check(self._lib.gpulin_set_change_log(self._h, int(capacity)))

    def changes(self, maxn: int):
        out = np.zeros(maxn, dtype=CHANGE_DTYPE)
        n = ctypes.c_int64(0)
        _check(self._lib.gpulin_get_changes(self._h, out.ctypes.data, maxn, ctypes.byref(n)))
        return out[:min(n.value, maxn)], n.value

    def changes_ptr(self, out_ptr: int, maxn: int) -> int:
        """The change log of the last call into a caller-owned (e.g. pinned) buffer of gpulin_change records; returns the
        number of changes produced."""
        n = ctypes.c_int64(0)
        _check(self._lib.gpulin_get_changes(self._h, out_ptr, maxn, ctypes.byref(n)))
        return n.value

    def redundant_rows(self):
        """rows (caller's numbering, ascending) that are redundant for the bounds on the device (cons_linear.c:7743)"""
        out = np.zeros(max(self.nrows, 1), dtype=np.int32)
        n = ctypes.c_int64(0)
        _check(self._lib.gpulin_get_redundant_rows(self._h, out.ctypes.data, self.nrows, ctypes.byref(n)))
        return out[:n.value].copy()

    def layout(self) -> dict:
        st = np.zeros(12, dtype=np.int64)
        _check(self._lib.gpulin_get_layout(self._h, st.ctypes.data, 12))
        keys = ("nnz", "stored_nnz", "rows_thread", "rows_stream", "rows_block", "device_bytes", "tiles",
                "blocks_thread", "blocks_stream", "maxlen", "rows_unit", "blocks_bittable")
        return dict(zip(keys, (int(x) for x in st)))

    def call_stats(self) -> dict:
        """what the last propagate call launched (kernel launches, dense / sparse rounds, one-block call, hand-over)"""
        st = np.zeros(6, dtype=np.int64)
        _check(self._lib.gpulin_get_call_stats(self._h, st.ctypes.data, 6))
        return dict(zip(("launches", "dense_rounds", "sparse_rounds", "small_call", "resumed", "fast_rows"), (int(x) for x in st)))

    def algorithmic_bytes(self) -> int:
        b = ctypes.c_int64(0)
        _check(self._lib.gpulin_algorithmic_bytes(self._h, ctypes.byref(b)))
        return b.value

    def profile_round(self):
        """one full round on the current bounds; CUDA-event ms of (filter sweep, exact kernel, apply kernel)"""
        a, b, c = ctypes.c_double(0), ctypes.c_double(0), ctypes.c_double(0)
        _check(self._lib.gpulin_profile_round(self._h, ctypes.byref(a), ctypes.byref(b), ctypes.byref(c)))
        return a.value, b.value, c.value

    # -- single rounds (multi-GPU, profiling) ------------------------------------------------------------------
    def set_stream(self, cuda_stream: int):
        _check(self._lib.gpulin_set_stream(self._h, cuda_stream))

    def sync(self):
        _check(self._lib.gpulin_sync(self._h))

    def exchange_buffer(self):
        ptr = _P()
        n = ctypes.c_int64(0)
        _check(self._lib.gpulin_exchange_buffer(self._h, ctypes.byref(ptr), ctypes.byref(n)))
        return ptr.value, n.value

    def get_keys(self):
        keys = np.empty(2 * self.ncols + 2, dtype=np.int64)
        _check(self._lib.gpulin_get_keys(self._h, keys.ctypes.data))
        return keys

    def set_keys(self, keys):
        keys = np.ascontiguousarray(keys, dtype=np.int64)
        assert keys.shape == (2 * self.ncols + 2,)
        _check(self._lib.gpulin_set_keys(self._h, keys.ctypes.data))

    def peer_handles(self, nranks: int) -> bytes:
        n = ctypes.c_int64(0)
        _check(self._lib.gpulin_peer_handles(self._h, int(nranks), None, ctypes.byref(n)))
        buf = ctypes.create_string_buffer(n.value)
        _check(self._lib.gpulin_peer_handles(self._h, int(nranks), buf, ctypes.byref(n)))
        return buf.raw

    def peer_connect(self, rank: int, handles):
        """``handles``: list of the blobs of all ranks in rank order"""
        blob = b"".join(handles)
        _check(self._lib.gpulin_peer_connect(self._h, int(rank), len(handles), blob))

    def mark_all(self):
        _check(self._lib.gpulin_mark_all(self._h))

    def round_begin(self):
        _check(self._lib.gpulin_round_begin(self._h))

    def round_sweep(self):
        _check(self._lib.gpulin_round_sweep(self._h))

    def round_apply(self, dense: bool = False, fetch: bool = True):
        if not fetch:
            _check(self._lib.gpulin_round_apply(self._h, int(dense), None, None))
            return None
        nchg = ctypes.c_int64(0)
        cutoff = ctypes.c_int32(0)
        _check(self._lib.gpulin_round_apply(self._h, int(dense), ctypes.byref(nchg), ctypes.byref(cutoff)))
        return nchg.value, cutoff.value


def propagate(prob, lb=None, ub=None, maxrounds: int = 0, device: int = 0, rangedrow: bool = False, tie=None, **numerics) -> dict:
    """one-shot: build, propagate to the fixpoint, read back; returns dict(status, lb, ub, nrounds, nchanges, ...)"""
    with LinearPropagator(prob, device=device, **numerics) as lp:
        lb = prob["lb"] if lb is None else lb
        ub = prob["ub"] if ub is None else ub
        if rangedrow:
            lp.set_rangedrow(True, lb, ub, tie)
        lp.set_bounds(lb, ub)
        res = lp.propagate(maxrounds)
        res["lb"], res["ub"] = lp.get_bounds()
    return res
