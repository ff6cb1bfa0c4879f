"""CPU: the C-ABI library builds for sm_100a, loads, and exports every symbol include/gpulin.h declares; without a
CUDA device every compute entry point fails loudly (no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import ROOT


def declared_symbols():
    with open(os.path.join(ROOT, "include", "gpulin.h")) as f:
        text = f.read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(gpulin_[a-z_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(gpulin):
    lib = ctypes.CDLL(gpulin.library_path())
    names = declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"libgpulin.so does not export {n}"
    assert set(gpulin.exported_symbols()) == set(names)


def test_default_numerics_are_the_reference_defaults(gpulin):
    num = gpulin.default_numerics()
    assert (num.infinity, num.epsilon, num.sumepsilon, num.feastol, num.boundstreps, num.hugeval,
            num.maxeasyactivitydelta) == (1e20, 1e-9, 1e-6, 1e-6, 0.05, 1e15, 1e6)


def test_invalid_arguments_are_rejected_before_touching_cuda(gpulin):
    lib = gpulin.load_library()
    h = ctypes.c_void_p()
    rowptr = np.array([0, 2], dtype=np.int64)
    col = np.array([0, 5], dtype=np.int32)          # column 5 out of range
    val = np.array([1.0, 1.0])
    side = np.array([0.0])
    vt = np.zeros(2, dtype=np.uint8)
    rc = lib.gpulin_create(0, 1, 2, 2, rowptr.ctypes.data, col.ctypes.data, val.ctypes.data, side.ctypes.data,
                           side.ctypes.data, vt.ctypes.data, None, ctypes.byref(h))
    assert rc == -2 and b"out of range" in lib.gpulin_last_error()
    col[1] = 1
    val[1] = 0.0                                     # zero coefficient (cons_linear.c:5422 asserts against it)
    rc = lib.gpulin_create(0, 1, 2, 2, rowptr.ctypes.data, col.ctypes.data, val.ctypes.data, side.ctypes.data,
                           side.ctypes.data, vt.ctypes.data, None, ctypes.byref(h))
    assert rc == -2


def test_no_cpu_fallback(gpulin):
    lib = gpulin.load_library()
    if lib.gpulin_device_count() > 0:
        pytest.skip("a CUDA device is present")
    prob = dict(rowptr=[0, 1], colidx=[0], vals=[1.0], lhs=[0.0], rhs=[1.0], vartype=[0], lb=[0.0], ub=[2.0])
    with pytest.raises(gpulin.GpulinError):
        gpulin.LinearPropagator(prob)


def test_product_does_not_reference_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "scip_b200")):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".c", ".h", ".cpp")):
                with open(os.path.join(dirpath, fn)) as f:
                    src = f.read()
                assert "import oracle" not in src and "liboracle" not in src and "linprop_oracle" not in src, fn
