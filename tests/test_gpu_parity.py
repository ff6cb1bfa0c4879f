"""GPU: the CUDA path, called through the C ABI, against (a) the fixtures of the UNMODIFIED reference and (b) the CPU
oracle on the same seeded inputs.  Bar (BASELINE.json): integer bounds and verdicts exact, continuous bounds within
1e-9 relative."""
import numpy as np
import pytest

import oracle
from conftest import (assert_bounds_match, golden_names, golden_ranged_names, load_golden, load_golden_redundant,
                      load_golden_tie)
from scip_b200 import synth

pytestmark = pytest.mark.gpu
NAMES = golden_names()
INF = 1e20


def gpu_vs_oracle(gpulin, prob, maxrounds=0, what="", **numerics):
    want = oracle.propagate(prob, maxrounds=maxrounds, **numerics)
    got = gpulin.propagate(prob, maxrounds=maxrounds, **numerics)
    assert got["status"] == want["status"], f"{what}: verdict {got['status']} != {want['status']}"
    if want["status"] != oracle.STATUS_CUTOFF:
        assert got["nrounds"] == want["nrounds"], what
        assert got["nchanges"] == want["nchanges"], what
        assert_bounds_match(got["lb"], got["ub"], want["lb"], want["ub"], prob["vartype"], what=what)
    return got, want


@pytest.mark.parametrize("name", NAMES)
def test_reference_fixpoint_exact_protocol(gpulin, name):
    """CUDA fixpoint == the reference's own cons_linear fixpoint (numerics/boundstreps = 1e-9 on both sides)"""
    prob, ref = load_golden(name, "1e-9")
    got = gpulin.propagate(prob, maxrounds=1000, boundstreps=1e-9)
    assert (got["status"] == gpulin.CUTOFF) == ref["infeasible"]
    if not ref["infeasible"]:
        assert got["status"] == gpulin.FIXPOINT
        assert_bounds_match(got["lb"], got["ub"], ref["lb"], ref["ub"], prob["vartype"], what=name)


@pytest.mark.parametrize("name", NAMES)
@pytest.mark.parametrize("bs", [1e-9, 0.05])
def test_oracle_parity_on_golden_inputs(gpulin, name, bs):
    prob, _ = load_golden(name, "1e-9")
    gpu_vs_oracle(gpulin, prob, maxrounds=1000, what=f"{name} bs={bs}", boundstreps=bs)


@pytest.mark.parametrize("name", NAMES)
def test_redundant_rows_match_the_reference_and_the_oracle(gpulin, name):
    """redundancy feedback (gpulin_get_redundant_rows) after the fixpoint: the rows the reference deletes with
    SCIPdelConsLocal in propagateCons (cons_linear.c:7743-7753), and the oracle's verdict on the same bounds"""
    prob, ref = load_golden(name, "1e-9")
    with gpulin.LinearPropagator(prob, boundstreps=1e-9) as lp:
        lp.set_bounds(prob["lb"], prob["ub"])
        res = lp.propagate(1000)
        if ref["infeasible"]:
            assert res["status"] == gpulin.CUTOFF
            return
        got = np.zeros(len(prob["lhs"]), dtype=bool)
        got[lp.redundant_rows()] = True
        lb, ub = lp.get_bounds()
    assert np.array_equal(got, load_golden_redundant(name)), name
    assert np.array_equal(got, oracle.redundant_rows(prob, lb, ub, boundstreps=1e-9)), name


@pytest.mark.parametrize("gen", ["setcover", "mixedknap", "unitnet"])
def test_redundant_rows_on_synthetic_instances(gpulin, gen):
    prob = {"setcover": lambda: synth.setcover(100_000, 100_000, 1_000_000, seed=2),
            "mixedknap": lambda: synth.mixed_knapsack(4000, 40_000, 1_000_000, seed=21, dense_range=(1500, 6000), eq_frac=0.2),
            "unitnet": lambda: synth.unit_network(50_000, 40_000, 250_000, seed=8)}[gen]()
    with gpulin.LinearPropagator(prob) as lp:
        lp.set_bounds(prob["lb"], prob["ub"])
        before = lp.redundant_rows()
        assert np.array_equal(before, np.flatnonzero(oracle.redundant_rows(prob, prob["lb"], prob["ub"])))
        res = lp.propagate(500)
        assert res["status"] == gpulin.FIXPOINT
        after = lp.redundant_rows()
        lb, ub = lp.get_bounds()
    assert np.array_equal(after, np.flatnonzero(oracle.redundant_rows(prob, lb, ub)))
    assert set(before.tolist()) <= set(after.tolist())      # tightening never makes a redundant row matter again
    assert gen == "mixedknap" or len(after) > 0


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_setcover_100k(gpulin, seed):
    prob = synth.setcover(100_000, 100_000, 1_000_000, seed=seed)
    got, _ = gpu_vs_oracle(gpulin, prob, what=f"setcover seed {seed}")
    assert got["nchanges"] > 0


@pytest.mark.parametrize("seed,bs", [(4, 1e-9), (5, 0.05)])
def test_unit_rows_with_signs_take_the_bit_table_sweep(gpulin, seed, bs):
    # +1 / -1 rows over binaries, integers and (partly unbounded) continuous variables, large enough for the bit-table
    # variant of the thread-per-row sweep: the unit rows are stored first and their values are never read by the filter
    prob = synth.unit_network(200_000, 150_000, 1_000_000, seed=seed)
    with gpulin.LinearPropagator(prob) as lp:
        lay = lp.layout()
    assert lay["rows_unit"] >= 199_000 and lay["rows_unit"] % 32 == 0 and lay["blocks_bittable"] > 0
    got, _ = gpu_vs_oracle(gpulin, prob, maxrounds=500, what=f"unitnet seed {seed}", boundstreps=bs)
    assert got["nchanges"] > 0


def test_unit_and_weighted_rows_mixed_in_the_thread_per_row_bin(gpulin):
    # a set-cover instance (85 % unit rows, 15 % knapsack rows with weights) whose SELL bin holds both storage classes
    prob = synth.setcover(200_000, 200_000, 2_000_000, seed=6)
    with gpulin.LinearPropagator(prob) as lp:
        lay = lp.layout()
    assert 0 < lay["rows_unit"] < lay["rows_thread"] and lay["blocks_bittable"] > 0
    gpu_vs_oracle(gpulin, prob, what="setcover 200k")


def test_setcover_infeasible(gpulin):
    prob = synth.setcover(50_000, 50_000, 500_000, seed=5, infeasible=True)
    got, want = gpu_vs_oracle(gpulin, prob, what="setcover infeasible")
    assert got["status"] == gpulin.CUTOFF


@pytest.mark.parametrize("seed,bs", [(21, 1e-9), (22, 0.05), (23, 1e-9)])
def test_mixed_knapsack_all_row_bins(gpulin, seed, bs):
    # normal rows 50..350 (warp-per-row), 1 % dense rows 1500..6000 (block-per-row), equalities, general integers
    prob = synth.mixed_knapsack(4000, 40_000, 1_000_000, seed=seed, dense_range=(1500, 6000), eq_frac=0.2)
    with gpulin.LinearPropagator(prob) as lp:
        lay = lp.layout()
    assert lay["rows_stream"] > 0 and lay["rows_block"] > 0
    got, _ = gpu_vs_oracle(gpulin, prob, maxrounds=200, what=f"mixedknap seed {seed}", boundstreps=bs)
    assert got["nchanges"] > 0


def test_mixed_knapsack_infeasible(gpulin):
    prob = synth.mixed_knapsack(2000, 20_000, 400_000, seed=31, dense_range=(1500, 3000), infeasible=True)
    got, _ = gpu_vs_oracle(gpulin, prob, what="mixedknap infeasible")
    assert got["status"] == gpulin.CUTOFF


def test_very_long_row_beyond_shared_memory_staging(gpulin):
    # one row with 60k nonzeros: longer than the shared-memory alpha staging of the block-per-row kernel
    rng = np.random.default_rng(7)
    ncols = 80_000
    cols = rng.permutation(ncols)[:60_000].astype(np.int32)
    vals = rng.integers(1, 10, size=cols.size).astype(np.float64)
    rowptr = np.array([0, cols.size, cols.size + 3], dtype=np.int64)
    colidx = np.concatenate([cols, np.array([0, 1, 2], dtype=np.int32)])
    vals = np.concatenate([vals, np.array([1.0, 1.0, 1.0])])
    ub = np.full(ncols, 5.0)
    prob = dict(rowptr=rowptr, colidx=colidx, vals=vals, lhs=np.array([-INF, 2.0]), rhs=np.array([30.0, INF]),
                lb=np.zeros(ncols), ub=ub, vartype=(rng.random(ncols) < 0.5).astype(np.uint8))
    got, _ = gpu_vs_oracle(gpulin, prob, what="long row")
    assert got["nchanges"] > 10_000


def tiny(rows, lhs, rhs, lb, ub, vartype):
    rowptr = np.cumsum([0] + [len(r) for r in rows]).astype(np.int64)
    colidx = np.array([c for r in rows for c, _ in r], dtype=np.int32)
    vals = np.array([v for r in rows for _, v in r], dtype=np.float64)
    return dict(rowptr=rowptr, colidx=colidx, vals=vals, lhs=np.array(lhs, dtype=np.float64),
                rhs=np.array(rhs, dtype=np.float64), lb=np.array(lb, dtype=np.float64),
                ub=np.array(ub, dtype=np.float64), vartype=np.array(vartype, dtype=np.uint8))


def test_edge_cases(gpulin):
    cases = {
        # singleton row: force = TRUE (cons_linear.c:7029), tightens below the relative threshold
        "singleton": tiny([[(0, 2.0)]], [-INF], [7.0], [0.0], [100.0], [1]),
        # infinite bounds: one unbounded variable -> residual path (tightenVarBounds)
        "one_infinite": tiny([[(0, 1.0), (1, 1.0), (2, -1.0)]], [-INF], [10.0], [-INF, 0.0, 0.0], [INF, 4.0, 3.0], [0, 0, 0]),
        "two_infinite": tiny([[(0, 1.0), (1, 1.0)]], [-INF], [10.0], [-INF, -INF], [INF, INF], [0, 0]),
        # huge contributions (>= 1e15) are counted, not summed
        "huge": tiny([[(0, 1e10), (1, 1.0)]], [-INF], [5.0], [0.0, 0.0], [1e6, 10.0], [0, 0]),
        # empty row and a row that is infeasible on its own
        "empty_row": tiny([[], [(0, 1.0)]], [-1.0, -INF], [1.0, 3.0], [0.0], [9.0], [1]),
        "empty_row_infeasible": tiny([[], [(0, 1.0)]], [1.0, -INF], [INF, 3.0], [0.0], [9.0], [1]),
        "row_infeasible": tiny([[(0, 1.0), (1, 1.0)]], [5.0], [INF], [0.0, 0.0], [1.0, 1.0], [1, 1]),
        # equality with a continuous variable and negative zero bounds
        "negzero": tiny([[(0, 1.0), (1, 2.0)]], [4.0], [4.0], [-0.0, -0.0], [10.0, 10.0], [0, 1]),
        # integrality rounding: 3 x <= 10 -> x <= 3
        "rounding": tiny([[(0, 3.0)], [(0, 1.0), (1, 1.0)]], [-INF, 5.5], [10.0, INF], [0.0, 0.0], [100.0, 2.7], [1, 0]),
        # bounds that cross within one round by more than feastol -> cutoff in the apply step
        "crossing": tiny([[(0, 1.0), (1, 1.0)], [(0, 1.0), (1, -1.0)]], [-INF, 9.0], [5.0, INF], [0.0, 0.0], [10.0, 10.0], [0, 0]),
        # fixed variables only: maxactdelta == 0 -> row skipped
        "all_fixed": tiny([[(0, 1.0), (1, 1.0)]], [-INF], [5.0], [1.0, 2.0], [1.0, 2.0], [1, 1]),
    }
    for name, prob in cases.items():
        for bs in (0.05, 1e-9):
            gpu_vs_oracle(gpulin, prob, maxrounds=100, what=f"{name} bs={bs}", boundstreps=bs)


def test_no_rows_and_no_columns(gpulin):
    prob = tiny([], [], [], [0.0, 1.0], [2.0, 3.0], [0, 1])
    got = gpulin.propagate(prob)
    assert got["status"] == gpulin.FIXPOINT and got["nchanges"] == 0
    assert np.array_equal(got["lb"], prob["lb"]) and np.array_equal(got["ub"], prob["ub"])


def test_round_limit_and_resume(gpulin):
    prob, _ = load_golden("egout", "1e-9")          # needs 17 Jacobi rounds
    want = oracle.propagate(prob, boundstreps=1e-9)
    with gpulin.LinearPropagator(prob, boundstreps=1e-9) as lp:
        r1 = lp.propagate(5)
        assert r1["status"] == gpulin.ROUNDLIMIT and r1["nrounds"] == 5
        lb5, ub5 = lp.get_bounds()
        w5 = oracle.propagate(prob, boundstreps=1e-9, maxrounds=5)
        assert_bounds_match(lb5, ub5, w5["lb"], w5["ub"], prob["vartype"], what="egout after 5 rounds")
        r2 = lp.propagate(0)                          # continues from the device-resident state
        assert r2["status"] == gpulin.FIXPOINT and r1["nrounds"] + r2["nrounds"] == want["nrounds"]
        lb, ub = lp.get_bounds()
        assert_bounds_match(lb, ub, want["lb"], want["ub"], prob["vartype"], what="egout resumed")
        r3 = lp.propagate(0)                          # idempotent: nothing is marked any more
        assert r3["nchanges"] == 0 and r3["nnz_processed"] == 0


def test_host_loop_equals_graph_loop(gpulin, monkeypatch):
    prob = synth.setcover(20_000, 20_000, 200_000, seed=9)
    a = gpulin.propagate(prob)
    monkeypatch.setenv("GPULIN_LOOP", "host")
    b = gpulin.propagate(prob)
    assert a["status"] == b["status"] and a["nrounds"] == b["nrounds"] and a["nchanges"] == b["nchanges"]
    assert np.array_equal(a["lb"], b["lb"]) and np.array_equal(a["ub"], b["ub"])


def test_update_bounds_marks_only_affected_rows(gpulin):
    prob = synth.setcover(20_000, 20_000, 200_000, seed=10)
    with gpulin.LinearPropagator(prob) as lp:
        lp.propagate()
        lb, ub = lp.get_bounds()
        free = np.flatnonzero(lb < ub)[:50]
        # probing-style change: fix 50 free binaries to 1 (SCIPchgVarLbProbing, scip_probing.c:302)
        lp.update_bounds(free, np.ones(50), np.ones(50))
        res = lp.propagate()
        assert 0 < res["nnz_processed"] < lp.nnz // 10     # incremental: only rows of touched columns are swept
        glb, gub = lp.get_bounds()
    lb2, ub2 = lb.copy(), ub.copy()
    lb2[free] = 1.0
    want = oracle.propagate(prob, lb=lb2, ub=ub2)
    assert res["status"] == want["status"]
    if want["status"] != oracle.STATUS_CUTOFF:
        assert_bounds_match(glb, gub, want["lb"], want["ub"], prob["vartype"], what="after update_bounds")


@pytest.mark.parametrize("small", ["1", "0"])
def test_dive_of_incremental_calls_matches_oracle(gpulin, monkeypatch, small):
    """a dive (SCIPchgVarLb/UbProbing + SCIPpropagateProbing, step after step): every call after a few updated bounds runs
    in one block (probe_kernel, GPULIN_SMALL=1, the default) or through the general loop (=0); both equal the oracle"""
    monkeypatch.setenv("GPULIN_SMALL", small)
    prob = synth.setcover(20_000, 20_000, 200_000, seed=12)
    rng = np.random.default_rng(12)
    with gpulin.LinearPropagator(prob) as lp:
        lp.set_change_log(100_000)
        lp.propagate()
        lb, ub = lp.get_bounds()
        for step in range(12):
            free = np.flatnonzero(lb < ub)
            if len(free) == 0:
                break
            nfix = 1 if step < 8 else 40            # the last steps fix many columns: the cascade outgrows one block
            var = free[rng.permutation(len(free))[:nfix]]
            val = rng.integers(0, 2, size=len(var)).astype(np.float64)
            lp.update_bounds(var, val, val)
            res = lp.propagate()
            lb[var] = val
            ub[var] = val
            want = oracle.propagate(prob, lb=lb, ub=ub)
            assert res["status"] == want["status"], f"step {step}"
            if want["status"] == oracle.STATUS_CUTOFF:
                break
            assert res["nrounds"] == want["nrounds"] and res["nchanges"] == want["nchanges"], f"step {step}"
            log, n = lp.changes(100_000)
            assert n == res["nchanges"]
            glb, gub = lp.get_bounds()
            assert_bounds_match(glb, gub, want["lb"], want["ub"], prob["vartype"], what=f"dive step {step}")
            lb, ub = glb, gub


def test_change_log_replays_to_the_fixpoint(gpulin):
    prob, _ = load_golden("dcmulti", "1e-9")
    with gpulin.LinearPropagator(prob, boundstreps=1e-9) as lp:
        lp.set_change_log(10_000)
        res = lp.propagate()
        log, n = lp.changes(10_000)
        lb, ub = lp.get_bounds()
    assert n == res["nchanges"] == len(log)
    assert np.all(np.diff(log["round"]) >= 0)            # round ordered
    rl, ru = prob["lb"] + 0.0, prob["ub"] + 0.0
    for rec in log:
        if rec["is_upper"]:
            assert rec["newbound"] < ru[rec["var"]]
            ru[rec["var"]] = rec["newbound"]
        else:
            assert rec["newbound"] > rl[rec["var"]]
            rl[rec["var"]] = rec["newbound"]
    assert np.array_equal(rl, lb) and np.array_equal(ru, ub)


def test_single_round_api_matches_oracle_sweep(gpulin):
    prob = synth.mixed_knapsack(1500, 15_000, 300_000, seed=41, dense_range=(1200, 2500))
    c, wl, wu = oracle.sweep(prob, prob["lb"], prob["ub"], boundstreps=1e-9)
    with gpulin.LinearPropagator(prob, boundstreps=1e-9) as lp:
        lp.round_begin()
        lp.round_sweep()
        nchg, cutoff = lp.round_apply(dense=True)
        lb, ub = lp.get_bounds()
    assert cutoff == c
    assert nchg == int((wl != prob["lb"]).sum() + (wu != prob["ub"]).sum())
    assert_bounds_match(lb, ub, wl, wu, prob["vartype"], what="single round")


@pytest.mark.parametrize("nparts", [2, 4])
def test_row_partition_with_min_merge_equals_single_device(gpulin, nparts):
    """the multi-GPU protocol on one device: row blocks in separate handles, candidate keys merged by elementwise
    MIN (what ncclMin does), dense apply"""
    prob = synth.mixed_knapsack(1200, 12_000, 240_000, seed=43, dense_range=(1100, 2000))
    want = oracle.propagate(prob)
    nrows = len(prob["lhs"])
    cuts = np.searchsorted(prob["rowptr"], np.linspace(0, prob["rowptr"][-1], nparts + 1)).clip(0, nrows)
    cuts[0], cuts[-1] = 0, nrows
    parts = [gpulin.LinearPropagator(prob, rows=(cuts[i], cuts[i + 1])) for i in range(nparts)]
    try:
        for lp in parts:
            lp.set_bounds(prob["lb"], prob["ub"])
            lp.round_begin()
        rounds = 0
        while True:
            rounds += 1
            bufs = []
            for lp in parts:
                lp.round_sweep()
                bufs.append(lp.get_keys())
            merged = np.minimum.reduce(bufs)
            out = []
            for lp in parts:
                lp.set_keys(merged)
                out.append(lp.round_apply(dense=True))
            assert len(set(out)) == 1
            nchg, cutoff = out[0]
            if cutoff or nchg == 0 or rounds > 500:
                break
        assert (want["status"] == oracle.STATUS_CUTOFF) == bool(cutoff)
        if not cutoff:
            assert rounds == want["nrounds"]
            for lp in parts:
                lb, ub = lp.get_bounds()
                assert_bounds_match(lb, ub, want["lb"], want["ub"], prob["vartype"], what="row partition")
    finally:
        for lp in parts:
            lp.close()


# ---- BASELINE.json configs at their full size (the oracle needs 7 s for C3 and 14 s for C4 on one host core) -------------

def test_c3_full_size(gpulin):
    """BASELINE configs[2]: set cover 1M x 1M, 10M nonzeros, seed 1 -- every bound against the oracle"""
    prob = synth.setcover(1_000_000, 1_000_000, 10_000_000, seed=1)
    got, want = gpu_vs_oracle(gpulin, prob, what="c3 full size")
    assert got["status"] == gpulin.FIXPOINT and got["nchanges"] > 200_000


def test_c4_full_size(gpulin):
    """BASELINE configs[3]: mixed knapsack / general integer 200k x 2M, 50M nonzeros, 1 % dense rows, seed 2"""
    prob = synth.mixed_knapsack(200_000, 2_000_000, 50_000_000, seed=2)
    with gpulin.LinearPropagator(prob) as lp:
        lay = lp.layout()
    assert lay["rows_stream"] > 150_000 and lay["rows_block"] > 1000
    got, want = gpu_vs_oracle(gpulin, prob, what="c4 full size")
    assert got["status"] == gpulin.FIXPOINT and got["nchanges"] > 500_000


def test_row_longer_than_65536_nonzeros(gpulin):
    """the error bound of the interval filter is computed in double (an int n*n overflows from 46341 nonzeros on): a row
    of 70k nonzeros that is quiet, one that tightens, and one that sits within rounding distance of the gate"""
    rng = np.random.default_rng(17)
    ncols = 90_000
    n = 70_000
    rows, lhs, rhs = [], [], []
    for kind in range(3):
        cols = rng.permutation(ncols)[:n].astype(np.int32)
        vals = rng.integers(1, 10, size=n).astype(np.float64)
        rows.append((cols, vals))
        tot = float(vals.sum()) * 5.0            # maximal activity for x in [0, 5]
        lhs.append(-INF)
        rhs.append({0: tot + 100.0,              # quiet: the maximal activity stays below the side
                    1: 20.0,                     # tightens every variable with a coefficient of 5 and more
                    2: 45.0 + 5e-10}[kind])      # maxdelta = 45 vs slack 45 + 5e-10: on the gate (eps = 1e-9)
    rowptr = np.cumsum([0] + [len(c) for c, _ in rows]).astype(np.int64)
    prob = dict(rowptr=rowptr, colidx=np.concatenate([c for c, _ in rows]), vals=np.concatenate([v for _, v in rows]),
                lhs=np.array(lhs), rhs=np.array(rhs), lb=np.zeros(ncols), ub=np.full(ncols, 5.0),
                vartype=(rng.random(ncols) < 0.5).astype(np.uint8))
    for bs in (0.05, 1e-9):
        got, _ = gpu_vs_oracle(gpulin, prob, maxrounds=200, what=f"70k-nonzero rows bs={bs}", boundstreps=bs)
    assert got["nchanges"] > 10_000


def test_cancellation_and_huge_counts_follow_the_reference(gpulin):
    """SURVEY 8a row a7 (activities that cancel by twelve orders of magnitude once a variable is fixed) and rows with three
    and more huge contributions (the verdict uses count * hugeval): the generators behind the fixtures syn_edge_*"""
    for name, prob, cutoff in (("cancel", synth.edge_cancellation(), False), ("huge", synth.edge_huge(), False),
                               ("huge infeasible", synth.edge_huge(infeasible=True), True)):
        for bs in (0.05, 1e-9):
            got, _ = gpu_vs_oracle(gpulin, prob, maxrounds=100, what=f"{name} bs={bs}", boundstreps=bs)
            assert (got["status"] == gpulin.CUTOFF) == cutoff, name


def test_packed_bounds_and_packed_change_log(gpulin):
    """gpulin_set_bounds_packed (2 bits per column against resident reference bounds + an explicit list) gives the state
    gpulin_set_bounds gives; gpulin_get_changes_packed returns the log of gpulin_get_changes in 12-byte records"""
    prob = synth.mixed_knapsack(1500, 15_000, 300_000, seed=41, dense_range=(1200, 2500))
    rng = np.random.default_rng(5)
    n = len(prob["lb"])
    lb, ub = prob["lb"].copy(), prob["ub"].copy()
    fin = (np.abs(lb) < 1e19) & (np.abs(ub) < 1e19)
    u = rng.random(n)
    tolb = fin & (u < 0.03)
    toub = fin & (u >= 0.03) & (u < 0.05)
    shrink = fin & (u >= 0.05) & (u < 0.07) & (ub - lb >= 2.0)
    ub[tolb] = lb[tolb]
    lb[toub] = ub[toub]
    ub[shrink] = np.floor(0.5 * (lb[shrink] + ub[shrink]))
    with gpulin.LinearPropagator(prob) as lp:
        lp.set_change_log(200_000)
        lp.set_bounds(lb, ub)
        a = lp.propagate()
        alb, aub = lp.get_bounds()
        log, nlog = lp.changes(200_000)
        lp.set_reference_bounds(prob["lb"], prob["ub"])
        words, idx, elb, eub = lp.pack_bounds(lb, ub)
        assert len(words) == (n + 15) // 16 and 0 < len(idx) <= int(shrink.sum()) + 1
        lp.set_bounds_packed(words, idx, elb, eub)
        b = lp.propagate()
        blb, bub = lp.get_bounds()
        var, upper, val, npk = lp.changes_packed(200_000)
    want = oracle.propagate(prob, lb=lb, ub=ub)
    assert a["status"] == b["status"] == want["status"]
    if want["status"] != oracle.STATUS_CUTOFF:
        assert (a["nrounds"], a["nchanges"]) == (b["nrounds"], b["nchanges"]) == (want["nrounds"], want["nchanges"])
        assert np.array_equal(alb, blb) and np.array_equal(aub, bub)
        assert_bounds_match(blb, bub, want["lb"], want["ub"], prob["vartype"], what="packed bounds")
        assert npk == nlog == b["nchanges"]
        # same entries (the order inside a round is whatever the atomics made it: compare as sets per round)
        key = lambda v, up, x: sorted(zip(v.tolist(), up.tolist(), x.tolist()))  # noqa: E731
        assert key(var, upper, val) == key(log["var"], log["is_upper"], log["newbound"])


@pytest.mark.parametrize("gen", ["setcover", "mixedknap"])
def test_compact_change_log_is_the_change_log(gpulin, gen):
    """gpulin_get_changes_compact (one word per change whose new bound is 0 or 1, a side list for the others) returns the
    log of gpulin_get_changes entry by entry, in the same order"""
    prob = {"setcover": lambda: synth.setcover(60_000, 60_000, 600_000, seed=19),
            "mixedknap": lambda: synth.mixed_knapsack(1500, 15_000, 300_000, seed=42, dense_range=(1200, 2500))}[gen]()
    with gpulin.LinearPropagator(prob) as lp:
        lp.set_change_log(400_000)
        lp.set_bounds(prob["lb"], prob["ub"])
        res = lp.propagate()
        log, nlog = lp.changes(400_000)
        var, upper, val, ncp = lp.changes_compact(400_000)
    assert res["status"] == gpulin.FIXPOINT and nlog == ncp == res["nchanges"] > 0
    assert np.array_equal(var, log["var"]) and np.array_equal(upper, log["is_upper"]) and np.array_equal(val, log["newbound"])
    other = (log["newbound"] != 0.0) & (log["newbound"] != 1.0)
    assert (gen == "setcover") == (not other.any())       # binaries: nothing in the side list; general integers: something


# ---- ranged-row (gcd) propagation: rangedRowPropagation, cons_linear.c:5715-6696 ------------------------------------------

@pytest.mark.parametrize("name", golden_ranged_names())
def test_ranged_row_propagation_reaches_the_reference_fixpoint(gpulin, name):
    """the reference ran with constraints/linear/rangedrowpropagation = TRUE (rangedrowartcons = FALSE)"""
    prob, ref = load_golden(name, "1e-9")
    tie = load_golden_tie(name)
    got = gpulin.propagate(prob, maxrounds=1000, boundstreps=1e-9, rangedrow=True, tie=tie)
    assert (got["status"] == gpulin.CUTOFF) == ref["infeasible"]
    if not ref["infeasible"]:
        assert_bounds_match(got["lb"], got["ub"], ref["lb"] + 0.0, ref["ub"] + 0.0, prob["vartype"], what=name)
    else:
        assert gpulin.propagate(prob, maxrounds=1000, boundstreps=1e-9)["status"] == gpulin.FIXPOINT    # only the gcd rule sees it
    for bs in (1e-9, 0.05):
        want = oracle.propagate(prob, maxrounds=1000, boundstreps=bs, rangedrow=True, tie=tie)
        got = gpulin.propagate(prob, maxrounds=1000, boundstreps=bs, rangedrow=True, tie=tie)
        assert got["status"] == want["status"], (name, bs)
        if want["status"] != oracle.STATUS_CUTOFF:
            assert (got["nrounds"], got["nchanges"]) == (want["nrounds"], want["nchanges"]), (name, bs)
            assert_bounds_match(got["lb"], got["ub"], want["lb"], want["ub"], prob["vartype"], what=f"{name} bs={bs}")


@pytest.mark.parametrize("seed", [41, 42, 43, 44])
def test_ranged_row_propagation_matches_the_oracle_on_larger_instances(gpulin, seed):
    # dense rounds (the rule in front of the filter sweeps), small rounds and incremental calls (the rule in front of the
    # exact rules of the listed rows); rows longer than one warp pass
    prob = synth.ranged_rows(20_000 if seed < 43 else 3000, 30_000 if seed < 43 else 4000, seed=seed, infeasible=(seed == 44))
    want = oracle.propagate(prob, maxrounds=1000, rangedrow=True)
    plain = oracle.propagate(prob, maxrounds=1000)
    with gpulin.LinearPropagator(prob) as lp:
        lp.set_rangedrow(True, prob["lb"], prob["ub"])
        lp.set_bounds(prob["lb"], prob["ub"])
        got = lp.propagate(1000)
        assert got["status"] == want["status"]
        if want["status"] == oracle.STATUS_CUTOFF:
            return
        assert (got["nrounds"], got["nchanges"]) == (want["nrounds"], want["nchanges"])
        lb, ub = lp.get_bounds()
        assert_bounds_match(lb, ub, want["lb"], want["ub"], prob["vartype"], what=f"ranged seed {seed}")
        assert ((plain["lb"] != want["lb"]) | (plain["ub"] != want["ub"])).sum() > 0
        # an incremental call: a few integer variables fixed, one block (or the general loop) continues
        free = np.flatnonzero((ub - lb >= 2.0) & (prob["vartype"] != 0))[:4]
        lp.update_bounds(free, lb[free], lb[free])
        r2 = lp.propagate(1000)
        lb2, ub2 = lb.copy(), ub.copy()
        ub2[free] = lb[free]
        w2 = oracle.propagate(prob, lb=lb2, ub=ub2, maxrounds=1000, rangedrow=True, sortlb=prob["lb"], sortub=prob["ub"])
        assert r2["status"] == w2["status"]
        if w2["status"] != oracle.STATUS_CUTOFF:
            glb, gub = lp.get_bounds()
            assert_bounds_match(glb, gub, w2["lb"], w2["ub"], prob["vartype"], what=f"ranged seed {seed} incremental")


# ---- integer rows over integral columns: the exact kernel takes their activities from the filter sweep (FastAcc) ----------
def _fast_vs_oracle(gpulin, prob, what, expect_fast, **numerics):
    """propagates with the thread-per-row phase of the exact kernel forced on (GPULIN_FASTMIN=0 is set by the caller)
    and compares bounds, verdict, rounds and changes with the oracle; returns the number of rows that took the phase"""
    want = oracle.propagate(prob, **numerics)
    with gpulin.LinearPropagator(prob, **numerics) as lp:
        lp.set_bounds(prob["lb"], prob["ub"])
        got = lp.propagate(0)
        lb, ub = lp.get_bounds()
        fast = lp.call_stats()["fast_rows"]
    assert got["status"] == want["status"], what
    if want["status"] != oracle.STATUS_CUTOFF:
        assert got["nrounds"] == want["nrounds"] and got["nchanges"] == want["nchanges"], what
        assert_bounds_match(lb, ub, want["lb"], want["ub"], prob["vartype"], what=what)
    assert (fast > 0) == expect_fast, f"{what}: {fast} rows took the thread-per-row phase"
    return fast


@pytest.mark.parametrize("gen,bs", [("setcover", 1e-9), ("setcover", 0.05), ("setcover_small", 1e-9), ("infeasible", 1e-9),
                                    ("shortknap", 1e-9), ("shortknap", 0.05), ("unitnet", 1e-9)])
def test_integer_rows_take_their_activities_from_the_filter(gpulin, monkeypatch, gen, bs):
    monkeypatch.setenv("GPULIN_FASTMIN", "0")
    prob = {"setcover": lambda: synth.setcover(200_000, 200_000, 2_000_000, seed=12),       # bit-table sweep, unit + weighted
            "setcover_small": lambda: synth.setcover(20_000, 20_000, 200_000, seed=13),     # gather variant of the sweep
            "infeasible": lambda: synth.setcover(50_000, 50_000, 500_000, seed=14, infeasible=True),
            # short knapsack rows over binaries, general integers and 20 % continuous columns: rows with and without the flag
            "shortknap": lambda: synth.mixed_knapsack(60_000, 40_000, 900_000, seed=15, dense_frac=0.0, len_range=(3, 30)),
            "unitnet": lambda: synth.unit_network(100_000, 80_000, 500_000, seed=16)}[gen]()
    if gen == "infeasible":
        want = oracle.propagate(prob, boundstreps=bs)
        got = gpulin.propagate(prob, boundstreps=bs)
        assert got["status"] == want["status"] == oracle.STATUS_CUTOFF
    else:
        _fast_vs_oracle(gpulin, prob, f"{gen} bs={bs}", expect_fast=True, boundstreps=bs)


def test_fractional_bounds_of_integer_columns_switch_the_shortcut_off(gpulin, monkeypatch):
    """a bound of an integral column that is not an integer makes the filter's sums inexact: no row may skip its activity
    pass then (the flag is sticky until the next gpulin_set_bounds, which looks at every column again)"""
    monkeypatch.setenv("GPULIN_FASTMIN", "0")
    prob = synth.setcover(50_000, 50_000, 500_000, seed=17)
    frac = dict(prob)
    frac["ub"] = prob["ub"].copy()
    frac["lb"] = prob["lb"].copy()
    free = np.flatnonzero((prob["lb"] == 0.0) & (prob["ub"] == 1.0))
    frac["ub"][free[:200:2]] = 1.5        # (relaxations: the instance stays feasible)
    frac["lb"][free[1:200:2]] = -0.5
    want = oracle.propagate(frac)
    with gpulin.LinearPropagator(prob) as lp:
        lp.set_bounds(frac["lb"], frac["ub"])
        got = lp.propagate(0)
        lb, ub = lp.get_bounds()
        assert lp.call_stats()["fast_rows"] == 0
        assert want["status"] == oracle.STATUS_FIXPOINT and want["nchanges"] > 0
        assert got["status"] == want["status"] and got["nrounds"] == want["nrounds"] and got["nchanges"] == want["nchanges"]
        assert_bounds_match(lb, ub, want["lb"], want["ub"], prob["vartype"], what="fractional bounds")
        # integral bounds again on the same handle: the shortcut is back, the result is the oracle's
        want = oracle.propagate(prob)
        lp.set_bounds(prob["lb"], prob["ub"])
        got = lp.propagate(0)
        lb, ub = lp.get_bounds()
        assert lp.call_stats()["fast_rows"] > 0
        assert got["status"] == want["status"] and got["nrounds"] == want["nrounds"] and got["nchanges"] == want["nchanges"]
        assert_bounds_match(lb, ub, want["lb"], want["ub"], prob["vartype"], what="integral bounds again")


def test_rows_whose_sums_could_round_keep_the_double_double_pass(gpulin, monkeypatch):
    """integer coefficients beyond 2^20, or products whose absolute sum reaches 2^40: the flag / the magnitude test keep
    such rows on the double-double path; the result is the oracle's either way"""
    monkeypatch.setenv("GPULIN_FASTMIN", "0")
    rng = np.random.default_rng(18)
    nrows, ncols = 4000, 3000
    rows, lhs, rhs = [], [], []
    lb = np.zeros(ncols)
    ub = np.full(ncols, 1e9)
    ub[: ncols // 2] = rng.integers(1, 50, size=ncols // 2).astype(np.float64)
    for r in range(nrows):
        n = int(rng.integers(2, 12))
        big = r % 2 == 0
        cols = rng.choice(ncols // 2, size=n, replace=False) + (ncols // 2 if r % 4 == 1 else 0)
        coef = rng.integers(1, 9, size=n).astype(np.float64) * (float(2 ** 21 + 1) if big else 1.0)
        x = np.floor(rng.random(n) * np.minimum(ub[cols], 40.0))
        rows.append(list(zip(cols.tolist(), coef.tolist())))
        lhs.append(-INF)
        rhs.append(float((coef * x).sum() + rng.integers(0, 5) * coef.min()))
    prob = synth._from_rows(rows, np.array(lhs), np.array(rhs), lb, ub, np.ones(ncols, dtype=np.uint8))
    want = oracle.propagate(prob)
    with gpulin.LinearPropagator(prob) as lp:
        lp.set_bounds(prob["lb"], prob["ub"])
        got = lp.propagate(0)
        lbg, ubg = lp.get_bounds()
    assert got["status"] == want["status"] and got["nrounds"] == want["nrounds"] and got["nchanges"] == want["nchanges"]
    assert_bounds_match(lbg, ubg, want["lb"], want["ub"], prob["vartype"], what="large integer data")
    assert got["nchanges"] > 0
