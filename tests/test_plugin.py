"""prop_gpulinear as a SCIP plugin: the unmodified reference (oracle/_ref/lib/libscip.so) is the host application, the
propagator replaces cons_linear's bound tightening (constraints/linear/tightenboundsfreq = -1)."""
import json
import os
import subprocess

import numpy as np
import pytest

import oracle
from conftest import GOLDEN, ROOT, assert_bounds_match, golden_names, golden_ranged_names, load_golden

DRIVER = os.path.join(ROOT, "scip_b200", "plugin", "_build", "gpulinear_driver")
needs_driver = pytest.mark.skipif(not (os.path.exists(DRIVER) and oracle.have_reference()),
                                  reason="plugin driver / oracle/_ref not built (python -c 'import __graft_entry__ as g; g.build()')")


def run_driver(*args, check=True):
    res = subprocess.run([DRIVER, *args], capture_output=True, text=True, timeout=1200)
    if check:
        assert res.returncode == 0, res.stderr[-2000:]
        return json.loads([ln for ln in res.stdout.splitlines() if ln.startswith("{")][-1])
    return res


@needs_driver
def test_driver_cpu_mode_reproduces_the_golden_fixture(tmp_path):
    out = str(tmp_path / "o.lpr")
    run_driver("--lpb", os.path.join(GOLDEN, "dcmulti.lpb"), "--cpu", "--boundstreps", "1e-9", "--out", out)
    got = oracle.read_lpr(out)
    _, ref = load_golden("dcmulti", "1e-9")
    assert np.array_equal(got["lb"], ref["lb"]) and np.array_equal(got["ub"], ref["ub"])


@needs_driver
def test_plugin_fails_loudly_without_a_gpu(gpulin):
    if gpulin.load_library().gpulin_device_count() > 0:
        pytest.skip("a CUDA device is present")
    res = run_driver("--lpb", os.path.join(GOLDEN, "bell5.lpb"), check=False)
    assert res.returncode != 0 and "no CPU fallback" in res.stderr


@needs_driver
@pytest.mark.gpu
@pytest.mark.parametrize("name", golden_names())
def test_plugin_reaches_the_reference_fixpoint(tmp_path, name):
    """SCIP + prop_gpulinear (cons_linear tightening off) == SCIP's own cons_linear propagation, root node"""
    prob, ref = load_golden(name, "1e-9")
    out = str(tmp_path / "o.lpr")
    info = run_driver("--lpb", os.path.join(GOLDEN, name + ".lpb"), "--boundstreps", "1e-9", "--out", out)
    got = oracle.read_lpr(out)
    assert got["infeasible"] == ref["infeasible"], name
    assert info["gpu_prop_calls"] >= 1
    if not ref["infeasible"]:
        assert_bounds_match(got["lb"] + 0.0, got["ub"] + 0.0, ref["lb"] + 0.0, ref["ub"] + 0.0, prob["vartype"], what=name)
        if info["gpu_domreds"] > 0:
            assert info["linear_domreds"] == 0          # the replaced path really was off


@needs_driver
@pytest.mark.gpu
@pytest.mark.parametrize("gen", ["setcover", "mixedknap"])
def test_plugin_replays_a_long_change_log(tmp_path, gen):
    """a root propagation with thousands of changes: the plugin reads the log through gpulin_get_changes_compact (one
    word per change of a binary, a side list for general bounds) and replays it round by round -- SCIP ends with the
    bounds of the propagation fixpoint"""
    from scip_b200 import synth
    from scip_b200.lpb import write_lpb
    prob = {"setcover": lambda: synth.setcover(40_000, 40_000, 400_000, seed=53),
            "mixedknap": lambda: synth.mixed_knapsack(3000, 30_000, 600_000, seed=51, dense_range=(1200, 2500))}[gen]()
    want = oracle.propagate(prob, boundstreps=1e-9)
    assert want["status"] == oracle.STATUS_FIXPOINT and want["nchanges"] > 4096
    lpb = str(tmp_path / "p.lpb")
    out = str(tmp_path / "o.lpr")
    write_lpb(lpb, prob)
    info = run_driver("--lpb", lpb, "--boundstreps", "1e-9", "--out", out)
    got = oracle.read_lpr(out)
    assert not got["infeasible"] and info["gpu_prop_calls"] >= 1 and info["gpu_domreds"] > 4096
    assert_bounds_match(got["lb"] + 0.0, got["ub"] + 0.0, want["lb"] + 0.0, want["ub"] + 0.0, prob["vartype"], what=gen)


@needs_driver
@pytest.mark.gpu
def test_plugin_in_tree_search_with_conflict_analysis():
    """thousands of EXEC calls with local bounds, backtracking and PROPRESPROP: enigma is solved by propagation +
    branching (the reference build has no LP solver); both runs must prove the same status"""
    cpu = run_driver("--lpb", os.path.join(GOLDEN, "enigma.lpb"), "--cpu", "--solve")
    gpu = run_driver("--lpb", os.path.join(GOLDEN, "enigma.lpb"), "--solve")
    assert gpu["scip_status"] == cpu["scip_status"]
    assert gpu["gpu_prop_calls"] > 100 and gpu["linear_domreds"] == 0


@needs_driver
@pytest.mark.gpu
@pytest.mark.parametrize("presolve", [False, True])
def test_redundancy_feedback_in_tree_search(presolve):
    """propagating/gpulinear/delredundant: rows the device proves redundant for a node's bounds are deleted locally
    (SCIPdelConsLocal, the end of propagateCons cons_linear.c:7743-7753) -- same verdict and optimum"""
    extra = ["--presolve"] if presolve else []
    base = run_driver("--lpb", os.path.join(GOLDEN, "enigma.lpb"), "--solve", *extra)
    red = run_driver("--lpb", os.path.join(GOLDEN, "enigma.lpb"), "--solve", "--del-redundant", *extra)
    assert red["scip_status"] == base["scip_status"]
    assert abs(red["primal"] - base["primal"]) <= 1e-6
    assert red["gpu_builds"] >= 1 and red["gpu_prop_calls"] > 10


@needs_driver
@pytest.mark.gpu
def test_tree_search_is_reproducible():
    """the device logs the changes of a round in the order its atomics land; the plugin replays them sorted, so two runs
    of the same search visit the same nodes"""
    a = run_driver("--lpb", os.path.join(GOLDEN, "enigma.lpb"), "--solve", "--presolve")
    b = run_driver("--lpb", os.path.join(GOLDEN, "enigma.lpb"), "--solve", "--presolve")
    assert (a["nodes"], a["gpu_prop_calls"], a["gpu_domreds"]) == (b["nodes"], b["gpu_prop_calls"], b["gpu_domreds"])


@needs_driver
@pytest.mark.gpu
def test_device_copy_is_stable_across_the_tree():
    """constraint handlers delete rows locally that became redundant in a subtree (cons_linear.c:7743-7753) and backtracking
    brings them back: a device copy of the ACTIVE rows (--active-rows-only, the first policy of this plugin) is rebuilt
    whenever their number moves; the default copy of all EXISTING global rows is built once per change of the constraint
    set -- same verdict and optimum either way"""
    stable = run_driver("--lpb", os.path.join(GOLDEN, "enigma.lpb"), "--solve")
    active = run_driver("--lpb", os.path.join(GOLDEN, "enigma.lpb"), "--solve", "--active-rows-only")
    print("device copies built: stable", stable["gpu_builds"], "active-only", active["gpu_builds"],
          "calls", stable["gpu_prop_calls"], active["gpu_prop_calls"])
    assert stable["scip_status"] == active["scip_status"]
    assert abs(stable["primal"] - active["primal"]) <= 1e-6
    assert 1 <= stable["gpu_builds"] <= active["gpu_builds"]


@needs_driver
@pytest.mark.gpu
@pytest.mark.parametrize("name", ["p0548", "misc03", "lseu", "enigma"])
def test_plugin_reads_upgraded_rows_after_presolve(tmp_path, name):
    """with presolving on, most linear constraints are upgraded to knapsack / setppc / logicor / varbound constraints
    (consPresolLinear); the plugin reads their rows like SCIP's matrix view does (matrix.c:541-696), rewritten to active
    variables.  Root node: same verdict as the reference's own propagators, and -- the GPU rows are an additional
    propagator on the same transformed problem -- bounds at least as tight"""
    outs = {}
    info = {}
    for mode in ("cpu", "gpu"):
        out = str(tmp_path / f"{mode}.lpr")
        args = ["--lpb", os.path.join(GOLDEN, name + ".lpb"), "--presolve", "--boundstreps", "1e-9", "--out", out]
        info[mode] = run_driver(*(args + (["--cpu"] if mode == "cpu" else [])))
        outs[mode] = oracle.read_lpr(out)
    assert outs["gpu"]["infeasible"] == outs["cpu"]["infeasible"]
    rows = info["gpu"]["gpu_rows"]
    if info["gpu"]["gpu_prop_calls"] > 0:
        assert sum(rows) > 0 and sum(rows[1:]) > 0, rows      # upgraded rows are on the device
    if not outs["cpu"]["infeasible"]:
        tol = 1e-9 * np.maximum(1.0, np.abs(outs["cpu"]["lb"]))
        assert np.all(outs["gpu"]["lb"] >= outs["cpu"]["lb"] - tol)
        tol = 1e-9 * np.maximum(1.0, np.abs(outs["cpu"]["ub"]))
        assert np.all(outs["gpu"]["ub"] <= outs["cpu"]["ub"] + tol)


@needs_driver
@pytest.mark.gpu
def test_plugin_in_tree_search_after_presolve():
    """the tree search of enigma on the presolved problem (set partitioning / knapsack rows on the device, negated
    variables rewritten to active ones): same status and optimal value as the reference alone"""
    cpu = run_driver("--lpb", os.path.join(GOLDEN, "enigma.lpb"), "--cpu", "--solve", "--presolve")
    gpu = run_driver("--lpb", os.path.join(GOLDEN, "enigma.lpb"), "--solve", "--presolve")
    assert gpu["scip_status"] == cpu["scip_status"]
    assert abs(gpu["primal"] - cpu["primal"]) <= 1e-6
    assert gpu["gpu_prop_calls"] > 10


@needs_driver
@pytest.mark.gpu
@pytest.mark.parametrize("name", ["enigma", "p0548", "misc03", "lseu", "gt2"])
def test_batched_probing_entry_matches_scip_probing(name):
    """SCIPprobeBatchGpulinear (all probes of a node in one call: gpulin_probe_batch, one launch per probe) against SCIP's own
    probing cycle, probe by probe (SCIPstartProbing / SCIPchgVarLb|UbProbing / SCIPpropagateProbing / SCIPendProbing as in
    SCIPapplyProbingVar, prop_probing.c:1254-1279), at the propagated root node: same verdict for every probe"""
    res = subprocess.run([DRIVER, "--lpb", os.path.join(GOLDEN, name + ".lpb"), "--boundstreps", "1e-9", "--probe-batch", "400"],
                         capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [ln for ln in res.stdout.splitlines() if ln.startswith("PROBEBATCH ")]
    assert len(lines) == 1, res.stdout[-2000:]
    info = json.loads(lines[0][len("PROBEBATCH "):])
    assert info["probes"] > 0 and info["node_cutoff"] == 0
    assert info["mismatches"] == 0 and info["cutoffs_batch"] == info["cutoffs_scip"], info
    # SCIPprobeBatchBoundsGpulinear: the implied bounds of every probe (what SCIPapplyProbingVar returns as proplbs /
    # propubs, prop_probing.c:1203-1303) are the bounds SCIP's own probing cycle ends with
    assert info["implied_bound_mismatches"] == 0, info
    assert info["probes_with_changes"] == 0 or info["implied_bound_changes"] > 0, info


def _ngpus():
    try:
        from scip_b200 import propagator
        return propagator.load_library().gpulin_device_count()
    except Exception:
        return 0


@needs_driver
@pytest.mark.gpu
@pytest.mark.skipif(_ngpus() < 2, reason="needs two GPUs (gpurun --gpus 2)")
@pytest.mark.parametrize("name", ["bell5", "dcmulti", "egout", "p0548", "syn_mixedknap_300", "syn_setcover_2k_infeas"])
def test_plugin_with_two_devices_reaches_the_reference_fixpoint(tmp_path, name):
    """propagating/gpulinear/ndevices = 2: SCIP (one process, one thread) drives a copy of the rows on each of two GPUs
    through SCIP_DECL_PROPEXEC; the devices share the dense rounds (gpulin_group_connect) -- same fixpoint as the reference"""
    prob, ref = load_golden(name, "1e-9")
    out = str(tmp_path / "o.lpr")
    info = run_driver("--lpb", os.path.join(GOLDEN, name + ".lpb"), "--boundstreps", "1e-9", "--ndevices", "2", "--out", out)
    got = oracle.read_lpr(out)
    assert got["infeasible"] == ref["infeasible"], name
    assert info["gpu_prop_calls"] >= 1
    if not ref["infeasible"]:
        assert_bounds_match(got["lb"] + 0.0, got["ub"] + 0.0, ref["lb"] + 0.0, ref["ub"] + 0.0, prob["vartype"], what=name)


@needs_driver
@pytest.mark.gpu
@pytest.mark.skipif(_ngpus() < 2, reason="needs two GPUs (gpurun --gpus 2)")
def test_plugin_with_two_devices_in_tree_search():
    one = run_driver("--lpb", os.path.join(GOLDEN, "enigma.lpb"), "--solve")
    two = run_driver("--lpb", os.path.join(GOLDEN, "enigma.lpb"), "--solve", "--ndevices", "2")
    assert two["scip_status"] == one["scip_status"] and abs(two["primal"] - one["primal"]) <= 1e-6
    assert (two["nodes"], two["gpu_prop_calls"], two["gpu_domreds"]) == (one["nodes"], one["gpu_prop_calls"], one["gpu_domreds"])


@needs_driver
@pytest.mark.gpu
@pytest.mark.parametrize("name", golden_ranged_names())
def test_plugin_ranged_row_propagation(tmp_path, name):
    """propagating/gpulinear/rangedrow: the gcd rule of rangedRowPropagation (cons_linear.c:5715-6696) on the device, rows
    walked in cons_linear's sorted order -- the fixpoint of the reference run with rangedrowpropagation = TRUE"""
    prob, ref = load_golden(name, "1e-9")
    out = str(tmp_path / "o.lpr")
    info = run_driver("--lpb", os.path.join(GOLDEN, name + ".lpb"), "--boundstreps", "1e-9", "--rangedrow", "--out", out)
    got = oracle.read_lpr(out)
    assert got["infeasible"] == ref["infeasible"], name
    assert info["gpu_prop_calls"] >= 1 and info["linear_domreds"] == 0
    if not ref["infeasible"]:
        assert_bounds_match(got["lb"] + 0.0, got["ub"] + 0.0, ref["lb"] + 0.0, ref["ub"] + 0.0, prob["vartype"], what=name)
    # and the driver's own --cpu --rangedrow run is that reference
    cpu = str(tmp_path / "c.lpr")
    run_driver("--lpb", os.path.join(GOLDEN, name + ".lpb"), "--boundstreps", "1e-9", "--rangedrow", "--cpu", "--out", cpu)
    c = oracle.read_lpr(cpu)
    assert c["infeasible"] == ref["infeasible"]
    if not ref["infeasible"]:
        assert np.array_equal(c["lb"], ref["lb"]) and np.array_equal(c["ub"], ref["ub"])
