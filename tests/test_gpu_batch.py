"""GPU: a batch of probing bound vectors on one matrix (BASELINE config 5): clones of one handle share the device
matrix, every probe = backtrack to the node + one bound change + propagation, many probes in flight."""
import numpy as np
import pytest

import oracle
from conftest import assert_bounds_match
from scip_b200 import synth
from scip_b200.batch import ProbingBatch

pytestmark = pytest.mark.gpu


def _check_batch(gpulin, prob, var, lb, ub, nworkers, **numerics):
    with gpulin.LinearPropagator(prob, **numerics) as base:
        node = base.propagate()
        assert node["status"] == gpulin.FIXPOINT
        nlb, nub = base.get_bounds()
        pb = ProbingBatch(base, nworkers=nworkers)
        try:
            res = pb.run(var, lb, ub, want_bounds=True)
            again = pb.run(var, lb, ub)                      # the workers are reusable; results are reproducible
        finally:
            pb.close()
        # the native batch entry point (clones kept inside the handle, backtracking by undoing the change log)
        nat1 = base.probe_batch(var, lb, ub, nworkers=min(nworkers, 4))
        nat2 = base.probe_batch(var, lb, ub, nworkers=min(nworkers, 4))     # second batch: workers restore from their logs
        for k in ("status", "nrounds", "nchanges"):
            assert np.array_equal(nat1[k], res[k]) and np.array_equal(nat2[k], res[k]), k
        # ... and with the implied bounds of every probe (sparse: its change log replayed on the node's bounds)
        nat3 = base.probe_batch_changes(var, lb, ub, nworkers=min(nworkers, 4))
        for k in ("status", "nrounds", "nchanges"):
            assert np.array_equal(nat3[k], res[k]), k
        for i in range(len(var)):
            if res["status"][i] == gpulin.CUTOFF:
                continue
            plb, pub = nlb.copy(), nub.copy()
            plb[var[i]], pub[var[i]] = lb[i], ub[i]
            recs = nat3["changes"][nat3["chgbeg"][i]:nat3["chgbeg"][i + 1]]
            assert len(recs) == res["nchanges"][i] and np.all(np.diff(recs["round"]) >= 0)
            for rec in recs:
                if rec["is_upper"]:
                    pub[rec["var"]] = rec["newbound"]
                else:
                    plb[rec["var"]] = rec["newbound"]
            assert np.array_equal(plb, res["lb"][i]) and np.array_equal(pub, res["ub"][i]), f"probe {i}: implied bounds"
        blb, bub = base.get_bounds()
        assert np.array_equal(blb, nlb) and np.array_equal(bub, nub)      # the node itself is untouched
    assert np.array_equal(res["status"], again["status"]) and np.array_equal(res["nchanges"], again["nchanges"])
    ncut = 0
    for i in range(len(var)):
        plb, pub = nlb.copy(), nub.copy()
        plb[var[i]], pub[var[i]] = lb[i], ub[i]
        want = oracle.propagate(prob, lb=plb, ub=pub, **numerics)
        assert res["status"][i] == want["status"], f"probe {i}"
        if want["status"] == oracle.STATUS_CUTOFF:
            ncut += 1
            continue
        assert res["nrounds"][i] == want["nrounds"] and res["nchanges"][i] == want["nchanges"], f"probe {i}"
        assert_bounds_match(res["lb"][i], res["ub"][i], want["lb"], want["ub"], prob["vartype"], what=f"probe {i}")
    return ncut


def test_probing_batch_on_set_cover(gpulin):
    prob = synth.setcover(20_000, 20_000, 200_000, seed=3)
    node = oracle.propagate(prob)
    free = np.flatnonzero(node["lb"] < node["ub"])
    rng = np.random.default_rng(3)
    var = free[rng.integers(0, len(free), size=48)].astype(np.int32)
    val = rng.integers(0, 2, size=48).astype(np.float64)
    _check_batch(gpulin, prob, var, val, val, nworkers=8)


def test_probing_batch_on_mixed_integers(gpulin):
    prob = synth.mixed_knapsack(1500, 15_000, 300_000, seed=44, dense_range=(1200, 2500))
    node = oracle.propagate(prob, boundstreps=1e-9)
    free = np.flatnonzero((node["ub"] - node["lb"] >= 2.0) & (prob["vartype"] != 0))
    rng = np.random.default_rng(4)
    var = free[rng.integers(0, len(free), size=24)].astype(np.int32)
    mid = np.floor(0.5 * (node["lb"][var] + node["ub"][var]))
    up = rng.integers(0, 2, size=24).astype(bool)
    lb = np.where(up, mid + 1.0, node["lb"][var])
    ub = np.where(up, node["ub"][var], mid)
    _check_batch(gpulin, prob, var, lb, ub, nworkers=5, boundstreps=1e-9)


def test_c5_full_batch(gpulin):
    """BASELINE configs[4] at its full size: 1024 probes on the 5M-nnz set-cover matrix; verdict, rounds and number of
    changes of a seeded subset of 64 probes -- and their implied bounds, bound for bound -- against the CPU oracle"""
    prob = synth.setcover(500_000, 500_000, 5_000_000, seed=3)
    with gpulin.LinearPropagator(prob) as base:
        node = base.propagate()
        assert node["status"] == gpulin.FIXPOINT
        nlb, nub = base.get_bounds()
        free = np.flatnonzero(nlb < nub)
        rng = np.random.default_rng(3)
        var = free[rng.integers(0, len(free), size=1024)].astype(np.int32)
        val = rng.integers(0, 2, size=1024).astype(np.float64)
        res = base.probe_batch_changes(var, val, val, nworkers=64)
        blb, bub = base.get_bounds()
    assert np.array_equal(blb, nlb) and np.array_equal(bub, nub)
    assert res["chgbeg"][-1] == res["nchanges"].sum() == len(res["changes"])
    for i in rng.permutation(1024)[:64]:
        plb, pub = nlb.copy(), nub.copy()
        plb[var[i]] = pub[var[i]] = val[i]
        want = oracle.propagate(prob, lb=plb, ub=pub)
        assert res["status"][i] == want["status"], f"probe {i}"
        if want["status"] == oracle.STATUS_CUTOFF:
            continue
        assert res["nrounds"][i] == want["nrounds"] and res["nchanges"][i] == want["nchanges"], f"probe {i}"
        for rec in res["changes"][res["chgbeg"][i]:res["chgbeg"][i + 1]]:
            if rec["is_upper"]:
                pub[rec["var"]] = rec["newbound"]
            else:
                plb[rec["var"]] = rec["newbound"]
        assert np.array_equal(plb, want["lb"]) and np.array_equal(pub, want["ub"]), f"probe {i}: implied bounds"
