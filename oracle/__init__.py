"""CPU oracle for linear bound propagation -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may
import this package.  The product (``scip_b200``) never does.

Two checkers live here:

* ``oracle.propagate``  -- the plain-C Jacobi restatement (``linprop_oracle.c``) through ctypes;
* ``oracle.run_reference`` -- the UNMODIFIED reference (``oracle/_ref/ref_driver`` + ``libscip.so``, built by
  ``oracle/Makefile.ref`` from /root/reference) run as a subprocess on an ``.lpb`` or any file the reference reads.
"""
from __future__ import annotations

import ctypes
import json
import os
import struct
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

STATUS_FIXPOINT, STATUS_CUTOFF, STATUS_ROUNDLIMIT = 0, 1, 2


class _Numerics(ctypes.Structure):
    _fields_ = [(n, ctypes.c_double) for n in
                ("infinity", "epsilon", "sumepsilon", "feastol", "boundstreps", "hugeval", "maxeasyactivitydelta")]


class _Problem(ctypes.Structure):
    _fields_ = [("nrows", ctypes.c_int64), ("ncols", ctypes.c_int64), ("nnz", ctypes.c_int64),
                ("rowptr", ctypes.c_void_p), ("colidx", ctypes.c_void_p), ("vals", ctypes.c_void_p),
                ("lhs", ctypes.c_void_p), ("rhs", ctypes.c_void_p), ("vartype", ctypes.c_void_p)]


def build(force: bool = False) -> str:
    """compile liboracle.so (gcc) if missing or stale"""
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "linprop_oracle.c")
    hdr = os.path.join(_HERE, "linprop_oracle.h")
    if force or not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(["gcc", "-O2", "-std=c99", "-fPIC", "-shared", "-ffp-contract=off", "-o", so, src, "-lm"])
    return so


def _lib():
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(build())
        _LIB.oracle_propagate.restype = ctypes.c_int
        _LIB.oracle_propagate.argtypes = [ctypes.POINTER(_Problem), ctypes.POINTER(_Numerics), ctypes.c_void_p,
                                          ctypes.c_void_p, ctypes.c_int, ctypes.POINTER(ctypes.c_int),
                                          ctypes.POINTER(ctypes.c_int64)]
        _LIB.oracle_sweep.restype = ctypes.c_int
        _LIB.oracle_sweep.argtypes = [ctypes.POINTER(_Problem), ctypes.POINTER(_Numerics), ctypes.c_void_p,
                                      ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64,
                                      ctypes.c_int64]
        _LIB.oracle_dd_sum21.restype = None
        _LIB.oracle_dd_sum21.argtypes = [ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double),
                                         ctypes.c_double, ctypes.c_double, ctypes.c_double]
    return _LIB


def _numerics(**kw) -> _Numerics:
    d = dict(infinity=1e20, epsilon=1e-9, sumepsilon=1e-6, feastol=1e-6, boundstreps=0.05, hugeval=1e15,
             maxeasyactivitydelta=1e6)
    d.update(kw)
    return _Numerics(**d)


def _problem(prob):
    keep = dict(
        rowptr=np.ascontiguousarray(prob["rowptr"], dtype=np.int64),
        colidx=np.ascontiguousarray(prob["colidx"], dtype=np.int32),
        vals=np.ascontiguousarray(prob["vals"], dtype=np.float64),
        lhs=np.ascontiguousarray(prob["lhs"], dtype=np.float64),
        rhs=np.ascontiguousarray(prob["rhs"], dtype=np.float64),
        vartype=np.ascontiguousarray(prob["vartype"], dtype=np.uint8),
    )
    p = _Problem(len(keep["lhs"]), len(keep["vartype"]), len(keep["vals"]),
                 *[keep[k].ctypes.data for k in ("rowptr", "colidx", "vals", "lhs", "rhs", "vartype")])
    return p, keep


def propagate(prob, lb=None, ub=None, maxrounds: int = 0, rangedrow: bool = False, sortlb=None, sortub=None, tie=None,
              **numerics):
    """Jacobi fixpoint of the C restatement.  ``prob`` is a dict with rowptr/colidx/vals/lhs/rhs/vartype (+lb/ub).
    ``rangedrow``: with the gcd rule for ranged rows (rangedRowPropagation, cons_linear.c:5715-6696), which walks a row in
    the reference's sorted order: ``sortlb`` / ``sortub`` = the global bounds the reference sorted by (default: the
    bounds on entry), ``tie`` = SCIPvarGetProbindex of every column (default: the column index).
    Returns dict(status, lb, ub, nrounds, nchanges)."""
    p, keep = _problem(prob)
    num = _numerics(**numerics)
    lb = np.array(prob["lb"] if lb is None else lb, dtype=np.float64, copy=True)
    ub = np.array(prob["ub"] if ub is None else ub, dtype=np.float64, copy=True)
    nrounds = ctypes.c_int(0)
    nchg = ctypes.c_int64(0)
    if rangedrow:
        slb = None if sortlb is None else np.ascontiguousarray(sortlb, dtype=np.float64)
        sub = None if sortub is None else np.ascontiguousarray(sortub, dtype=np.float64)
        tb = None if tie is None else np.ascontiguousarray(tie, dtype=np.int32)
        fn = _lib().oracle_propagate_ranged
        fn.restype = ctypes.c_int
        fn.argtypes = [ctypes.POINTER(_Problem), ctypes.POINTER(_Numerics), ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int,
                       ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int64), ctypes.c_void_p, ctypes.c_void_p,
                       ctypes.c_void_p]
        st = fn(ctypes.byref(p), ctypes.byref(num), lb.ctypes.data, ub.ctypes.data, int(maxrounds), ctypes.byref(nrounds),
                ctypes.byref(nchg), None if slb is None else slb.ctypes.data, None if sub is None else sub.ctypes.data,
                None if tb is None else tb.ctypes.data)
        del keep
        return dict(status=int(st), lb=lb, ub=ub, nrounds=nrounds.value, nchanges=nchg.value)
    st = _lib().oracle_propagate(ctypes.byref(p), ctypes.byref(num), lb.ctypes.data, ub.ctypes.data, int(maxrounds),
                                 ctypes.byref(nrounds), ctypes.byref(nchg))
    del keep
    return dict(status=int(st), lb=lb, ub=ub, nrounds=nrounds.value, nchanges=nchg.value)


def sweep(prob, lb, ub, rowbegin=0, rowend=None, **numerics):
    """one synchronous sweep; returns (cutoff, newlb, newub)"""
    p, keep = _problem(prob)
    num = _numerics(**numerics)
    lb = np.ascontiguousarray(lb, dtype=np.float64) + 0.0
    ub = np.ascontiguousarray(ub, dtype=np.float64) + 0.0
    nlb, nub = lb.copy(), ub.copy()
    if rowend is None:
        rowend = p.nrows
    c = _lib().oracle_sweep(ctypes.byref(p), ctypes.byref(num), lb.ctypes.data, ub.ctypes.data, nlb.ctypes.data,
                            nub.ctypes.data, int(rowbegin), int(rowend))
    del keep
    return int(c), nlb, nub


def redundant_rows(prob, lb, ub, **numerics):
    """bool[nrows]: rows that are redundant for the bounds lb/ub (the verdict of propagateCons, cons_linear.c:7743)"""
    p, keep = _problem(prob)
    num = _numerics(**numerics)
    lb = np.ascontiguousarray(lb, dtype=np.float64) + 0.0
    ub = np.ascontiguousarray(ub, dtype=np.float64) + 0.0
    out = np.zeros(int(p.nrows), dtype=np.uint8)
    fn = _lib().oracle_redundant_rows
    fn.restype = ctypes.c_int64
    fn.argtypes = [ctypes.POINTER(_Problem), ctypes.POINTER(_Numerics), ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
    fn(ctypes.byref(p), ctypes.byref(num), lb.ctypes.data, ub.ctypes.data, out.ctypes.data)
    del keep
    return out.astype(bool)


def dd_sum21(ahi, alo, b):
    hi, lo = ctypes.c_double(0), ctypes.c_double(0)
    _lib().oracle_dd_sum21(ctypes.byref(hi), ctypes.byref(lo), ahi, alo, b)
    return hi.value, lo.value


# ---------------------------------------------------------------------------------------------------------------
# the unmodified reference (oracle/_ref)
# ---------------------------------------------------------------------------------------------------------------
REF_DRIVER = os.path.join(_HERE, "_ref", "ref_driver")


def have_reference() -> bool:
    return os.path.exists(REF_DRIVER) and os.path.exists(os.path.join(_HERE, "_ref", "lib", "libscip.so"))


def read_lpr(path):
    with open(path, "rb") as f:
        raw = f.read()
    assert raw[:8] == b"GPULPR01", "not an .lpr file"
    ncols, status, ncalls, ndomreds, proptime, solvetime = struct.unpack_from("<qiiqdd", raw, 8)
    off = 8 + 8 + 4 + 4 + 8 + 8 + 8
    lb = np.frombuffer(raw, dtype=np.float64, count=ncols, offset=off).copy()
    ub = np.frombuffer(raw, dtype=np.float64, count=ncols, offset=off + 8 * ncols).copy()
    return dict(infeasible=bool(status), lb=lb, ub=ub, prop_calls=ncalls, domreds=ndomreds, prop_time_s=proptime,
                solve_time_s=solvetime)


def run_reference(path, boundstreps=None, dump_lpb=None, out_lpr=None, is_lpb=None, timeout=3600, rangedrow=False,
                  dump_tie=None):
    """run the compiled reference on ``path`` (an .lpb, or any file a reference reader accepts); returns the
    parsed JSON summary (+ bounds if ``out_lpr`` is given)"""
    if not have_reference():
        raise RuntimeError("oracle/_ref is not built (make -C oracle ref; needs /root/reference)")
    if is_lpb is None:
        is_lpb = path.endswith(".lpb")
    cmd = [REF_DRIVER, "--lpb" if is_lpb else "--read", path]
    if boundstreps is not None:
        cmd += ["--boundstreps", repr(float(boundstreps))]
    if dump_lpb:
        cmd += ["--dump-lpb", dump_lpb]
    if out_lpr:
        cmd += ["--out", out_lpr]
    if rangedrow:
        cmd += ["--rangedrow"]
    if dump_tie:
        cmd += ["--dump-tie", dump_tie]
    out = subprocess.run(cmd, check=True, capture_output=True, text=True, timeout=timeout).stdout
    res = json.loads([ln for ln in out.splitlines() if ln.startswith("{")][-1])
    if out_lpr:
        res.update(read_lpr(out_lpr))
    return res
