"""``.lpb`` -- the flat binary problem format of the linear-propagation path.

One file = the linear rows of one (sub)problem in CSR plus the variable bounds and integrality flags, i.e. exactly
what ``prop_gpulinear`` extracts from SCIP through ``SCIPgetVarsLinear`` / ``SCIPgetValsLinear`` /
``SCIPgetLhsLinear`` / ``SCIPgetRhsLinear`` (reference ``cons_linear.h:263-307``) and ``SCIPvarGetLbLocal`` /
``SCIPvarGetUbLocal`` / ``SCIPvarIsIntegral``.  Layout (little endian)::

    char    magic[8] = "GPULPB01"
    int64   nrows, ncols, nnz
    int64   rowptr[nrows+1]
    int32   colidx[nnz]          (+ padding to a multiple of 8 bytes)
    float64 vals[nnz]
    float64 lhs[nrows], rhs[nrows]        (-1e20 / +1e20 = no side; SCIP's infinity)
    float64 lb[ncols],  ub[ncols]
    uint8   vartype[ncols]                (0 continuous, 1 integral)
"""
from __future__ import annotations

import numpy as np

MAGIC = b"GPULPB01"


def write_lpb(path, prob) -> None:
    rowptr = np.ascontiguousarray(prob["rowptr"], dtype=np.int64)
    colidx = np.ascontiguousarray(prob["colidx"], dtype=np.int32)
    vals = np.ascontiguousarray(prob["vals"], dtype=np.float64)
    nrows, ncols, nnz = len(rowptr) - 1, len(prob["lb"]), len(vals)
    with open(path, "wb") as f:
        f.write(MAGIC)
        f.write(np.array([nrows, ncols, nnz], dtype=np.int64).tobytes())
        f.write(rowptr.tobytes())
        f.write(colidx.tobytes())
        f.write(b"\0" * ((8 - (4 * nnz) % 8) % 8))
        f.write(vals.tobytes())
        for k in ("lhs", "rhs", "lb", "ub"):
            f.write(np.ascontiguousarray(prob[k], dtype=np.float64).tobytes())
        f.write(np.ascontiguousarray(prob["vartype"], dtype=np.uint8).tobytes())


def read_lpb(path) -> dict:
    with open(path, "rb") as f:
        raw = f.read()
    if raw[:8] != MAGIC:
        raise ValueError(f"{path}: not an .lpb file")
    nrows, ncols, nnz = (int(x) for x in np.frombuffer(raw, dtype=np.int64, count=3, offset=8))
    off = 32
    rowptr = np.frombuffer(raw, dtype=np.int64, count=nrows + 1, offset=off).copy()
    off += 8 * (nrows + 1)
    colidx = np.frombuffer(raw, dtype=np.int32, count=nnz, offset=off).copy()
    off += 4 * nnz + (8 - (4 * nnz) % 8) % 8
    out = dict(rowptr=rowptr, colidx=colidx)
    for k, n in (("vals", nnz), ("lhs", nrows), ("rhs", nrows), ("lb", ncols), ("ub", ncols)):
        out[k] = np.frombuffer(raw, dtype=np.float64, count=n, offset=off).copy()
        off += 8 * n
    out["vartype"] = np.frombuffer(raw, dtype=np.uint8, count=ncols, offset=off).copy()
    return out
