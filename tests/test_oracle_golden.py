"""CPU: the oracle (plain-C Jacobi restatement) against the fixtures produced by the UNMODIFIED reference."""
import numpy as np
import pytest

import oracle
from conftest import assert_bounds_match, golden_names, golden_ranged_names, load_golden, load_golden_redundant, load_golden_tie

NAMES = golden_names()


@pytest.mark.parametrize("name", NAMES)
def test_oracle_matches_reference_exact_protocol(name):
    # numerics/boundstreps = 1e-9 on both sides: the exact-parity protocol
    prob, ref = load_golden(name, "1e-9")
    res = oracle.propagate(prob, boundstreps=1e-9, maxrounds=1000)
    assert (res["status"] == oracle.STATUS_CUTOFF) == ref["infeasible"]
    if not ref["infeasible"]:
        assert res["status"] == oracle.STATUS_FIXPOINT
        assert_bounds_match(res["lb"], res["ub"], ref["lb"], ref["ub"], prob["vartype"], what=name)


@pytest.mark.parametrize("name", NAMES)
def test_oracle_default_threshold_verdict_and_validity(name):
    # default boundstreps = 0.05: the reference is order dependent (flugpl UE2); the verdict must agree and
    # every oracle bound must be implied by the tiny-threshold fixpoint (never tighter than it)
    prob, ref = load_golden(name, "0.05")
    res = oracle.propagate(prob, boundstreps=0.05, maxrounds=1000)
    assert (res["status"] == oracle.STATUS_CUTOFF) == ref["infeasible"]
    if ref["infeasible"]:
        return
    _, tight = load_golden(name, "1e-9")
    assert np.all(res["lb"] <= tight["lb"] + 1e-6 * np.maximum(1, np.abs(tight["lb"])))
    assert np.all(res["ub"] >= tight["ub"] - 1e-6 * np.maximum(1, np.abs(tight["ub"])))


def test_known_differences_at_default_threshold_are_the_documented_ones():
    prob, ref = load_golden("flugpl", "0.05")
    res = oracle.propagate(prob, boundstreps=0.05)
    bad = np.flatnonzero((res["lb"] != ref["lb"]) | (res["ub"] != ref["ub"]))
    assert bad.size == 1 and prob["vartype"][bad[0]] == 0
    assert res["ub"][bad[0]] == 1500.0 and ref["ub"][bad[0]] == 1440.0


def test_dbldbl_known_answers():
    # tests/src/misc/dbldblarith.c of the reference: 1.2345678e7 + 3.45670069507e-1 style sums must carry the
    # rounding error in the low word
    hi, lo = oracle.dd_sum21(1.0, 0.0, 1e-20)
    assert hi == 1.0 and lo == 1e-20
    hi, lo = oracle.dd_sum21(1e16, 0.0, 1.0)
    assert hi + lo == 1e16 and (hi - 1e16) + lo == 1.0
    acc = (0.0, 0.0)
    for _ in range(10):
        acc = oracle.dd_sum21(acc[0], acc[1], 0.1)
    # ten times the double 0.1, summed exactly: 1.0 + 10 * (0.1 - 1/10)
    from fractions import Fraction
    exact = 10 * Fraction(0.1)
    assert abs(Fraction(acc[0]) + Fraction(acc[1]) - exact) < Fraction(1, 2 ** 100)


def test_sweep_row_slices_compose():
    # sweeping two row blocks into the same candidate vectors == sweeping all rows (the multi-GPU row partition)
    prob, _ = load_golden("bell5", "1e-9")
    lb, ub = prob["lb"] + 0.0, prob["ub"] + 0.0
    c0, l0, u0 = oracle.sweep(prob, lb, ub, boundstreps=1e-9)
    n = len(prob["lhs"])
    c1, l1, u1 = oracle.sweep(prob, lb, ub, 0, n // 2, boundstreps=1e-9)
    c2, l2, u2 = oracle.sweep(prob, lb, ub, n // 2, n, boundstreps=1e-9)
    assert c0 == (c1 | c2)
    assert np.array_equal(np.maximum(l1, l2), l0) and np.array_equal(np.minimum(u1, u2), u0)


@pytest.mark.parametrize("name", NAMES)
def test_oracle_redundant_rows_are_the_rows_the_reference_deletes(name):
    # propagateCons removes a row whose activity bounds lie inside its sides (cons_linear.c:7743-7753); at the reference's
    # own root fixpoint the restated verdict marks exactly the rows the reference deleted
    prob, ref = load_golden(name, "1e-9")
    if ref["infeasible"]:
        pytest.skip("infeasible at the root: the reference stops at the cutoff")
    deleted = load_golden_redundant(name)
    assert deleted.shape == (len(prob["lhs"]),)
    assert np.array_equal(oracle.redundant_rows(prob, ref["lb"], ref["ub"], boundstreps=1e-9), deleted)


@pytest.mark.parametrize("name", golden_ranged_names())
@pytest.mark.parametrize("bs", ["1e-9", "0.05"])
def test_oracle_ranged_row_propagation_matches_the_reference(name, bs):
    # rangedRowPropagation (cons_linear.c:5715-6696; the gcd rule for equations and ranged rows) switched on in the
    # reference; the restatement walks every row in the reference's sorted order (consdataCompVarProp :3191)
    prob, ref = load_golden(name, bs)
    res = oracle.propagate(prob, boundstreps=float(bs), maxrounds=1000, rangedrow=True, tie=load_golden_tie(name))
    assert (res["status"] == oracle.STATUS_CUTOFF) == ref["infeasible"]
    if not ref["infeasible"]:
        if bs == "1e-9":
            assert_bounds_match(res["lb"], res["ub"], ref["lb"] + 0.0, ref["ub"] + 0.0, prob["vartype"], what=name)
        # the rule finds something the activity argument does not
        plain = oracle.propagate(prob, boundstreps=float(bs), maxrounds=1000)
        assert ((plain["lb"] != res["lb"]) | (plain["ub"] != res["ub"])).sum() > 0
    else:
        assert oracle.propagate(prob, boundstreps=float(bs), maxrounds=1000)["status"] == oracle.STATUS_FIXPOINT
