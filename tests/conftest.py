import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _manifest():
    with open(os.path.join(GOLDEN, "manifest.json")) as f:
        return json.load(f)


def golden_names():
    """fixtures of the parity protocol proper (ranged-row propagation off on both sides)"""
    return sorted(n for n, e in _manifest().items() if not e["1e-9"].get("rangedrow"))


def golden_ranged_names():
    """fixtures the reference produced with constraints/linear/rangedrowpropagation = TRUE (rangedrowartcons = FALSE)"""
    return sorted(n for n, e in _manifest().items() if e["1e-9"].get("rangedrow"))


def load_golden_tie(name):
    """SCIPvarGetProbindex of every column of a ranged-row fixture: the last sort key of the row order its rule walks"""
    return np.fromfile(os.path.join(GOLDEN, name + ".tie"), dtype=np.int32)


def load_golden(name, bs):
    """(problem dict from <name>.lpb, reference result from <name>.bs<bs>.lpr)"""
    import oracle
    from scip_b200.lpb import read_lpb
    prob = read_lpb(os.path.join(GOLDEN, name + ".lpb"))
    ref = oracle.read_lpr(os.path.join(GOLDEN, f"{name}.bs{bs}.lpr"))
    return prob, ref


def load_golden_redundant(name):
    """bool[nrows]: rows of <name>.lpb the reference removed as redundant during its root propagation
    (tests/golden/make_golden_redundant.py, numerics/boundstreps = 1e-9)"""
    with open(os.path.join(GOLDEN, name + ".bs1e-9.red"), "rb") as f:
        return np.frombuffer(f.read(), dtype=np.uint8).astype(bool)


def rel_diff(a, b):
    return np.abs(a - b) / np.maximum(1.0, np.maximum(np.abs(a), np.abs(b)))


def assert_bounds_match(got_lb, got_ub, want_lb, want_ub, vartype, tol=1e-9, what=""):
    """the parity bar of BASELINE.json: integer-variable bounds exact (value-equal, -0.0 == 0.0), continuous bounds
    within 1e-9 relative"""
    integral = np.asarray(vartype) != 0
    bad_int = np.flatnonzero(((got_lb != want_lb) | (got_ub != want_ub)) & integral)
    assert bad_int.size == 0, (f"{what}: {bad_int.size} integer bound mismatches, first var {bad_int[0]}: "
                               f"got [{got_lb[bad_int[0]]!r},{got_ub[bad_int[0]]!r}] "
                               f"want [{want_lb[bad_int[0]]!r},{want_ub[bad_int[0]]!r}]")
    cont = ~integral
    if cont.any():
        d = max(rel_diff(got_lb[cont], want_lb[cont]).max(), rel_diff(got_ub[cont], want_ub[cont]).max())
        assert d <= tol, f"{what}: continuous bounds differ by {d:.3e} relative (> {tol})"


@pytest.fixture(scope="session")
def gpulin():
    """the ctypes binding, with the in-tree library built if it is stale (nvcc only; no GPU needed to build)"""
    from scip_b200 import build, propagator
    build.build_library()
    propagator.load_library()
    return propagator
