#!/usr/bin/env python
"""bench.py -- activity-based linear bound propagation to the fixpoint on B200(s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c3|c3small|c4]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

One step = one propagation of the workload from its initial bounds to the fixpoint.  Workload: BASELINE.json configs[2],
the synthetic set-cover MIP 1M rows x 1M binaries, 10M nonzeros (seed 1) -- the configuration the headline metric is
quoted on.  At N > 1 every rank holds the whole instance; the work of the dense rounds is shared, one packed exchange per
dense round through NVLink peer memory, small rounds run redundantly (strong scaling of one fixpoint).

metric  propagation_fixpoint_nnz_per_s = nonzeros of the instance / time to reach the propagation fixpoint
value   bounds resident in HBM when the timed region starts (device-side reset of the bounds inside it)
e2e     the same through the C ABI with HOST buffers, as the plugin calls it at a node: H2D of the bounds (2 bits per
        column against resident reference bounds + an explicit list, pinned), fixpoint, D2H of the verdict + the
        round-ordered change log (one word per change whose new bound is 0 or 1, 16 bytes per other change), every step
roofline  one FULL ROUND at the fixpoint bounds (filter sweep + exact rules + collect + apply; cold L2): algorithmic bytes
        (nnz*12 + nrows*20 + ncols*17) / its CUDA-event time against the measured HBM copy peak of
        MEASURED_PEAKS.json; the dominant kernel (the filter sweep) alone and the first real round are sub-keys
cpu_baseline / --impl reference   the UNMODIFIED reference (oracle/_ref = SCIP's cons_linear propagation, built from
        /root/reference) on the host cores (1 thread: SCIP is single threaded) on a bounded sample of the workload;
        if oracle/_ref is missing the arm says so (kind "port" = the C restatement; never silently)
extras  BASELINE configs[3] (c4, 50M nonzeros) and configs[4] (c5, 1024 probing bound vectors split over the ranks) at
        every N; parity of the final bounds against the CPU oracle, once, outside the timed region
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "propagation_fixpoint_nnz_per_s"
UNIT = "nnz/s"

WORKLOADS = {
    "c3": dict(desc="synthetic set-cover MIP 1M rows x 1M binaries, 10M nnz, row length 5-20, seed 1 (BASELINE configs[2])",
               gen=lambda synth: synth.setcover(1_000_000, 1_000_000, 10_000_000, seed=1)),
    "c3small": dict(desc="synthetic set-cover MIP 100k rows x 100k binaries, 1M nnz, seed 1",
                    gen=lambda synth: synth.setcover(100_000, 100_000, 1_000_000, seed=1)),
    "c4": dict(desc="synthetic mixed knapsack/general-integer MIP 200k rows x 2M vars, 50M nnz, 1% dense rows, seed 2 "
                    "(BASELINE configs[3])",
               gen=lambda synth: synth.mixed_knapsack(200_000, 2_000_000, 50_000_000, seed=2)),
}
# bounded sample of the c3 family for the CPU arm: a solve of the full instance costs the reference 30 s of model
# transformation and clean-up around 6 s of propagation; 1/5 of the instance keeps K + W steps within a few minutes
CPU_SAMPLE = dict(desc="set-cover 200k rows x 200k binaries, 2M nnz, seed 1 (1/5 of the c3 workload, same generator)",
                  gen=lambda synth: synth.setcover(200_000, 200_000, 2_000_000, seed=1))


def measured_traffic(kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel from the committed ncu --set full capture"""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            return json.load(f).get(kernel)
    except Exception:
        return None


def measured_peak_gbs():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region"""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows = []
        self.stop = threading.Event()
        self.index = index
        self.thread = threading.Thread(target=self._run, daemon=True)

    def _run_nvml(self):
        """NVML directly (nvidia_ml_py): a sample every 2 ms -- a timed region of 10 ms is over before one nvidia-smi
        process has started"""
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
        mx = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
        get_reasons = getattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons", None) or pynvml.nvmlDeviceGetCurrentClocksThrottleReasons
        bits = ((0x8, "hw_slowdown"), (0x40, "hw_thermal_slowdown"), (0x20, "sw_thermal_slowdown"), (0x4, "sw_power_cap"))
        while not self.stop.is_set():
            sm = pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
            mask = int(get_reasons(h))
            self.rows.append([str(sm), str(mx), ""] + [("Active" if mask & b else "Not Active") for b, _ in bits])
            self.stop.wait(0.002)

    def _run(self):
        try:
            self._run_nvml()
            return
        except Exception:
            pass
        while not self.stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i",
                                      str(self.index)], capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self.stop.wait(0.1)

    def __enter__(self):
        self.thread.start()
        return self

    def __exit__(self, *exc):
        self.stop.set()
        self.thread.join(timeout=6)

    def summary(self):
        if not self.rows:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["unavailable"])
        sm = [float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[3 + i].lower().startswith("active") for r in self.rows)]
        return dict(sm_mhz=statistics.median(sm) if sm else None, sm_max_mhz=max(mx) if mx else None, reasons=reasons,
                    samples=len(self.rows))


# ---------------------------------------------------------------------------------------------------------------------
# CPU arm: the unmodified reference (oracle/_ref); the C restatement only if that is missing, and then it says so
# ---------------------------------------------------------------------------------------------------------------------
def cpu_reference_steps(steps, warmup, full_instance=False):
    """`steps` root propagations of the bounded sample by the reference, the model built once (ref_driver --repeat);
    returns (cpu_baseline dict, mean seconds per step, nnz of the sample)"""
    import oracle
    from scip_b200 import synth
    from scip_b200.lpb import write_lpb
    prob = CPU_SAMPLE["gen"](synth)
    nnz = len(prob["vals"])
    cores_avail = os.cpu_count()
    times = []
    extra = {}
    if oracle.have_reference():
        kind = "reference"
        with tempfile.TemporaryDirectory() as tmp:
            path = os.path.join(tmp, "sample.lpb")
            write_lpb(path, prob)
            out = subprocess.run([oracle.REF_DRIVER, "--lpb", path, "--repeat", str(warmup + steps)], check=True,
                                 capture_output=True, text=True, timeout=3600).stdout
            recs = [json.loads(ln[len("REPEAT "):]) for ln in out.splitlines() if ln.startswith("REPEAT ")]
            recs.append(json.loads([ln for ln in out.splitlines() if ln.startswith("{")][-1]))
            times = [float(r["prop_time_s"]) for r in recs[warmup:]]
            if full_instance:
                # one solve of the FULL workload as well (46 s of wall clock for 6 s of propagation): the same-config number
                full = WORKLOADS["c3"]["gen"](synth)
                fpath = os.path.join(tmp, "full.lpb")
                write_lpb(fpath, full)
                fres = oracle.run_reference(fpath)
                extra["full_instance"] = dict(workload=WORKLOADS["c3"]["desc"], steps=1, prop_time_s=float(fres["prop_time_s"]),
                                              value=len(full["vals"]) / float(fres["prop_time_s"]), unit=UNIT,
                                              prop_calls=int(fres["prop_calls"]), domreds=int(fres["domreds"]))
        what = ("SCIPconshdlrGetPropTime(linear) of the root propagation, SCIP 11 built from /root/reference (oracle/_ref), "
                "parity settings, model built once and solved repeatedly")
    else:
        kind = "port"
        sys.stderr.write("bench.py: oracle/_ref is NOT built -- the CPU arm falls back to the C restatement (kind = \"port\"): "
                         "this is not the reference\n")
        oracle.build()
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            oracle.propagate(prob)
            if i >= warmup:
                times.append(time.perf_counter() - t0)
        what = "oracle/linprop_oracle.c (C restatement, Jacobi rounds) -- NOT the reference: oracle/_ref is missing"
    t = statistics.mean(times)
    cb = dict(value=nnz / t, unit=UNIT, cores=1, cores_available=cores_avail, kind=kind,
              sample=f"{CPU_SAMPLE['desc']}; {what}; fixpoint {t * 1e3:.1f} ms")
    cb.update(extra)
    return cb, t, nnz


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from scip_b200 import synth  # noqa: F401
    cb, t, nnz = cpu_reference_steps(args.steps, args.warmup, full_instance=not args.no_extras)
    wl = WORKLOADS[args.workload]
    line = dict(metric=METRIC, value=cb["value"], unit=UNIT, n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
                ms_per_step=t * 1e3, higher_is_better=True, scaling="strong", vs_baseline=None, dtype="f64",
                data="synthetic", impl="reference",
                config=dict(workload=wl["desc"],
                            step="every step is the reference's root propagation of a bounded sample of this workload "
                                 "(cpu_baseline.sample); cpu_baseline.full_instance = one solve of the whole instance"),
                cpu_baseline=cb,
                e2e=dict(value=cb["value"], unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(line))


def cpu_probing_reference(synth):
    """the reference's own probing cycle (SCIPstartProbing / SCIPchgVarLb|UbProbing / SCIPpropagateProbing / SCIPendProbing,
    oracle/ref_driver.c --probe) at the propagated root of a bounded sample of the same family, 1 core"""
    import oracle
    from scip_b200.lpb import write_lpb
    if not oracle.have_reference():
        return dict(unavailable="oracle/_ref is not built")
    prob = synth.setcover(200_000, 200_000, 2_000_000, seed=3)
    with tempfile.TemporaryDirectory() as tmp:
        path = os.path.join(tmp, "sample.lpb")
        write_lpb(path, prob)
        out = subprocess.run([oracle.REF_DRIVER, "--lpb", path, "--probe", "512"], capture_output=True, text=True,
                             timeout=600).stdout
    lines = [ln for ln in out.splitlines() if ln.startswith("PROBING ")]
    if not lines:
        return dict(unavailable="ref_driver printed no PROBING line")
    info = json.loads(lines[0][len("PROBING "):])
    return dict(us_per_probe=info["probing_s"] / max(info["probes"], 1) * 1e6, probes=info["probes"], cores=1, kind="reference",
                sample="set-cover 200k x 200k, 2M nnz, seed 3; 512 probes at the propagated root, SCIP 11 built from /root/reference")


# ---------------------------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------------------------
class Harness:
    """one workload on this rank's GPU: the propagator (connected to the peers at N > 1), resident and pinned buffers"""

    def __init__(self, torch, dist, prob, world, rank, local_rank, stream, with_e2e=True):
        from scip_b200 import propagator, sharded
        self.torch, self.dist, self.world, self.rank, self.stream = torch, dist, world, rank, stream
        self.prob = prob
        self.nrows, self.ncols, self.nnz = len(prob["lhs"]), len(prob["lb"]), len(prob["vals"])
        t0 = time.perf_counter()
        if world == 1:
            self.lp = propagator.LinearPropagator(prob, device=local_rank)
        else:
            self.sp = sharded.PeerPropagator(prob, rank, world, local_rank)
            self.lp = self.sp.lp
        self.setup_ms = (time.perf_counter() - t0) * 1e3          # gpulin_create (+ the peer connection)
        self.lp.set_stream(stream.cuda_stream)
        lb0 = torch.from_numpy(prob["lb"] + 0.0)
        ub0 = torch.from_numpy(prob["ub"] + 0.0)
        self.d_lb0, self.d_ub0 = lb0.cuda(), ub0.cuda()           # resident initial bounds (device-timed arm)
        self.logcap = 2 * self.ncols
        self.lp.set_change_log(self.logcap)                       # the plugin always logs: on in both arms
        self.d2h_bytes = 0
        self.h2d_bytes = 0
        if with_e2e:
            # the call the plugin makes at a node: bounds in (packed against the reference = the original domains of
            # the columns: [0,1] for the binaries the instance fixes), verdict + round-ordered change log out
            ref_lb, ref_ub = prob["lb"] + 0.0, prob["ub"] + 0.0
            fixed_bin = (prob["vartype"] != 0) & (ref_lb == ref_ub) & ((ref_lb == 0.0) | (ref_lb == 1.0))
            ref_lb[fixed_bin], ref_ub[fixed_bin] = 0.0, 1.0
            self.lp.set_reference_bounds(ref_lb, ref_ub)
            words, idx, elb, eub = self.lp.pack_bounds(prob["lb"], prob["ub"])
            self.h_words = torch.from_numpy(words.view(np.int32)).pin_memory()
            self.h_idx = torch.from_numpy(idx).pin_memory()
            self.h_elb = torch.from_numpy(elb).pin_memory()
            self.h_eub = torch.from_numpy(eub).pin_memory()
            self.nexplicit = len(idx)
            self.h2d_bytes = 4 * len(words) + 20 * len(idx)
            self.h_chg = torch.empty(self.logcap * 4, dtype=torch.uint8).pin_memory()       # one word per change ...
            self.h_side = torch.empty(self.logcap * 12, dtype=torch.uint8).pin_memory()     # ... + bounds other than 0 / 1

    def step_resident(self):
        self.lp.set_bounds_ptr(self.d_lb0.data_ptr(), self.d_ub0.data_ptr(), on_device=True)
        return self.lp.propagate(0)

    def step_e2e(self):
        if self.nexplicit:
            self.lp.set_bounds_packed_ptr(self.h_words.data_ptr(), self.nexplicit, self.h_idx.data_ptr(),
                                          self.h_elb.data_ptr(), self.h_eub.data_ptr())
        else:
            self.lp.set_bounds_packed_ptr(self.h_words.data_ptr())
        res = self.lp.propagate(0)                                            # verdict: 32 bytes, synchronises
        n, nx = self.lp.changes_compact_ptr(self.h_chg.data_ptr(), self.logcap, self.h_side.data_ptr(), self.logcap)
        self.d2h_bytes = 4 * min(n, self.logcap) + 12 * min(nx, self.logcap) + 4 + 32
        return res

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def timed(self, stepfn, K, W):
        torch = self.torch
        for _ in range(W):
            res = stepfn()
        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(self.stream)
        for _ in range(K):
            res = stepfn()
        e1.record(self.stream)
        self.barrier()
        ms = e0.elapsed_time(e1)
        if self.world > 1:
            t = torch.tensor([ms], device="cuda", dtype=torch.float64)
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, res

    def launches_per_step(self):
        return 1 + self.lp.call_stats()["launches"]          # set_bounds + the call

    def parity(self):
        """final bounds of this rank against the CPU oracle (rank 0), and equality of the bounds over the ranks"""
        import oracle
        torch, dist = self.torch, self.dist
        self.step_resident()
        lb, ub = self.lp.get_bounds()
        out = {}
        if self.world > 1:
            h = np.frombuffer(lb.tobytes() + ub.tobytes(), dtype=np.uint64)
            chk = int(np.bitwise_xor.reduce(h * np.arange(1, len(h) + 1, dtype=np.uint64)) >> np.uint64(1))
            t = torch.tensor([chk, -chk], device="cuda", dtype=torch.int64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            out["ranks_identical"] = bool(int(t[0].item()) == -int(t[1].item()))
        if self.rank == 0:
            t0 = time.perf_counter()
            want = oracle.propagate(self.prob)
            integral = self.prob["vartype"] != 0
            bad = int((((lb != want["lb"]) | (ub != want["ub"])) & integral).sum())
            rel = lambda a, b: np.abs(a - b) / np.maximum(1.0, np.maximum(np.abs(a), np.abs(b)))  # noqa: E731
            cont = ~integral
            mrel = float(max(rel(lb[cont], want["lb"][cont]).max(), rel(ub[cont], want["ub"][cont]).max())) if cont.any() else 0.0
            out.update(int_mismatch=bad, max_rel_cont=mrel, checked_against="oracle/linprop_oracle.c (CPU, all bounds)",
                       oracle_s=round(time.perf_counter() - t0, 1))
        if self.world > 1:
            dist.barrier()
        return out

    def close(self):
        self.lp.close()


def probing_batch_extra(torch, dist, propagator, synth, world, rank, device, with_cpu=True):
    """BASELINE configs[4]: 1024 probes -- one free binary fixed to 0 or 1 each, SCIPapplyProbingVar's pattern -- on the
    5M-nnz set-cover matrix at its root fixpoint, partitioned over the ranks (replicas of the matrix, probes[rank::world],
    no collective on the data path); 64 workers per GPU, one launch per worker; wall clock of the whole batch through the
    C ABI with host buffers, max over the ranks"""
    prob = synth.setcover(500_000, 500_000, 5_000_000, seed=3)
    with propagator.LinearPropagator(prob, device=device) as base:
        base.propagate()
        lb, ub = base.get_bounds()
        free = np.flatnonzero(lb < ub)
        rng = np.random.default_rng(3)
        var = free[rng.integers(0, len(free), size=1024)].astype(np.int32)
        val = rng.integers(0, 2, size=1024).astype(np.float64)
        mine = slice(rank, None, world)
        base.probe_batch(var[:256], val[:256], val[:256], nworkers=64)
        times = []
        for _ in range(5):
            if world > 1:
                dist.barrier()
            t0 = time.perf_counter()
            res = base.probe_batch(var[mine], val[mine], val[mine], nworkers=64)
            dt = time.perf_counter() - t0
            if world > 1:
                t = torch.tensor([dt], device="cuda", dtype=torch.float64)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                dt = float(t.item())
            times.append(dt)
        stats = torch.tensor([float((res["status"] == 1).sum()), float(res["nrounds"].sum()), float(res["nchanges"].sum()),
                              float(len(res["status"]))], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(stats, op=dist.ReduceOp.SUM)
        stats = stats.tolist()
    t = min(times)
    out = dict(workload=f"1024 probing bound vectors on a 5M-nnz set-cover MIP (500k x 500k, seed 3), split over {world} GPU(s), "
                        "64 workers each", ms_per_batch=t * 1e3, us_per_probe=t / 1024 * 1e6, probes_per_s=1024 / t,
               probes=int(stats[3]), cutoffs=int(stats[0]), mean_rounds=stats[1] / max(stats[3], 1),
               mean_changes=stats[2] / max(stats[3], 1), scaling="strong (the batch is fixed, the probes are split)")
    if with_cpu and rank == 0:
        out["cpu_reference"] = cpu_probing_reference(synth)
    return out


def run_ours(args):
    import torch
    import torch.distributed as dist
    from scip_b200 import build, propagator, synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))
    if not os.path.exists(build.LIB):
        raise SystemExit(f"bench.py: {build.LIB} is missing (python -m scip_b200.build)")

    wl = WORKLOADS[args.workload]
    K, W = args.steps, args.warmup
    # a created stream: the library captures its round loop into a CUDA graph, which the legacy default stream forbids
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    hz = Harness(torch, dist, wl["gen"](synth), world, rank, local_rank, stream)
    lp = hz.lp
    abytes = lp.algorithmic_bytes()

    with ClockSampler(local_rank) as clocks:
        ms_res, res = hz.timed(hz.step_resident, K, W)
        round_stats = lp.round_stats()
        xstats = lp.exchange_stats() if world > 1 else None
        trace = lp.trace()
        launches = K * hz.launches_per_step()
        ms_e2e, res2 = hz.timed(hz.step_e2e, K, W)
        # one full round at the fixpoint bounds (all rows marked, nothing changes): filter sweep | exact | collect + apply
        prof, prof_warm = [], []
        if world == 1:
            # the sweep of c3 moves 89 MB, less than the 126 MB L2: flush it between the profiled rounds
            l2flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
            for i in range(W + max(K, 10)):       # back to back, as the rounds of a fixpoint follow each other
                t = lp.profile_round()
                if i >= W:
                    prof_warm.append(t)
            for i in range(W + max(K, 10)):       # cold: the number `roofline` reports
                l2flush.zero_()                   # 256 MB written: nothing of the previous round is left in the 126 MB L2
                l2flush[: 192 << 20].max()        # ... and read again, so that the sweep does not pay for evicting dirty lines
                t = lp.profile_round()
                if i >= W:
                    prof.append(t)
    parity = hz.parity() if not args.no_parity else None

    peak, peak_src = measured_peak_gbs()
    ms, rn, rc = round_stats
    line = None
    if rank == 0:
        line = dict(metric=METRIC, value=hz.nnz * K / (ms_res * 1e-3), unit=UNIT, n_gpus=world, steps=K, warmup=W,
                    ms_per_step=ms_res / K, higher_is_better=True, scaling="strong", vs_baseline=None, dtype="f64",
                    data="synthetic",
                    config=dict(workload=wl["desc"], nrows=hz.nrows, ncols=hz.ncols, nnz=hz.nnz, rounds=res["nrounds"],
                                changes=res["nchanges"], verdict=propagator.STATUS_NAMES[res["status"]],
                                parallelism=("1 GPU" if world == 1 else
                                             f"{world} GPUs: matrix and bounds replicated, the dense rounds' rows shared, one packed "
                                             "exchange per dense round through NVLink peer memory (stores from inside the round's "
                                             "kernels), small rounds redundant"),
                                l2=("a step touches more than the 126 MB L2 (matrix 120-600 MB, row and column arrays, CSC, change "
                                    "log); the profiled rounds of `roofline` run after a 256 MB L2 flush each") if args.workload != "c3small" else "fits L2",
                                loop="CUDA graph WHILE node (device-side)"),
                    e2e=dict(value=hz.nnz * K / (ms_e2e * 1e-3), unit=UNIT, ms_per_step=ms_e2e / K, h2d_bytes_per_step=hz.h2d_bytes,
                             d2h_bytes_per_step=hz.d2h_bytes,
                             h2d="bounds as 2 bits per column against resident reference bounds + explicit list (gpulin_set_bounds_packed)",
                             d2h="verdict + round-ordered change log, 4 bytes per change whose new bound is 0 or 1 and 16 per other change "
                                 "(gpulin_get_changes_compact: what the plugin replays)"),
                    gpu_launches=launches, clocks=clocks.summary(), status_consistent=res["status"] == res2["status"],
                    setup_ms=hz.setup_ms,
                    round_us=[round(float(x) * 1e3, 1) for x in ms], round_nnz=[int(x) for x in rn], round_changes=[int(x) for x in rc],
                    first_round_us=float(ms[0]) * 1e3, first_round_frac=abytes / (float(ms[0]) * 1e-3) / 1e9 / peak,
                    kernel_starts_us=[[n_, round(t_, 1)] for n_, t_ in trace[:48]])
        if xstats is not None:
            line["exchange"] = dict(before_us=[round(float(x) * 1e3, 1) for x in xstats[0]],
                                    wait_us=[round(float(x) * 1e3, 1) for x in xstats[1]],
                                    note="per round on rank 0: device time from the start of the round to its exchange (-1000: the "
                                         "round ran redundantly, no exchange) and the wait for the slowest peer")
        if parity is not None:
            line["parity"] = parity
        if world == 1:
            sweep_ms = statistics.mean(p[0] for p in prof)
            exact_ms = statistics.mean(p[1] for p in prof)
            apply_ms = statistics.mean(p[2] for p in prof)
            round_ms = sweep_ms + exact_ms + apply_ms
            warm_ms = statistics.mean(sum(p) for p in prof_warm)
            kname = {"c3": "sweep_sell_bits_kernel", "c3small": "sweep_sell_kernel", "c4": "sweep_stream_kernel"}[args.workload]
            traffic = measured_traffic(f"{kname}:{args.workload}")
            line["roofline"] = dict(bound="hbm", what="one full round at the fixpoint bounds: filter sweep + exact rules + collect + apply",
                                    achieved=abytes / (round_ms * 1e-3) / 1e9, peak=peak, unit="GB/s",
                                    frac=abytes / (round_ms * 1e-3) / 1e9 / peak, peak_source=peak_src,
                                    traffic=traffic, algorithmic_bytes=abytes, round_us=round_ms * 1e3,
                                    l2="flushed before every profiled round (256 MB written, 192 MB read back)",
                                    warm_l2=dict(round_us=warm_ms * 1e3, frac=abytes / (warm_ms * 1e-3) / 1e9 / peak,
                                                 note="rounds back to back as inside a fixpoint: what fits stays in L2"),
                                    kernel=dict(name=f"{kname} (the filter sweep: the dominant kernel of a round)", kernel_us=sweep_ms * 1e3,
                                                achieved=abytes / (sweep_ms * 1e-3) / 1e9, frac=abytes / (sweep_ms * 1e-3) / 1e9 / peak,
                                                dram_frac=(traffic / (sweep_ms * 1e-3) / 1e9 / peak) if traffic else None),
                                    stages_us=dict(sweep=sweep_ms * 1e3, exact=exact_ms * 1e3, collect_apply=apply_ms * 1e3),
                                    first_round_us=float(ms[0]) * 1e3, first_round_frac=abytes / (float(ms[0]) * 1e-3) / 1e9 / peak)
            if not args.no_cpu:
                cb, _, _ = cpu_reference_steps(1, 0)
                line["cpu_baseline"] = cb
    hz.close()
    del hz

    # ---- the other BASELINE configs, at every N
    if not args.no_extras:
        extras = {}
        if args.workload != "c4":
            h4 = Harness(torch, dist, WORKLOADS["c4"]["gen"](synth), world, rank, local_rank, stream, with_e2e=False)
            ms4, res4 = h4.timed(h4.step_resident, K, W)
            m4, n4, c4 = h4.lp.round_stats()
            e4 = dict(workload=WORKLOADS["c4"]["desc"], ms_per_step=ms4 / K, value=h4.nnz * K / (ms4 * 1e-3), unit=UNIT,
                      rounds=res4["nrounds"], changes=res4["nchanges"], round_us=[round(float(x) * 1e3, 1) for x in m4],
                      setup_ms=h4.setup_ms, kernel_starts_us=[[n_, round(t_, 1)] for n_, t_ in h4.lp.trace()[:32]])
            if world == 1:
                pr = [h4.lp.profile_round() for _ in range(W + 5)][W:]
                a4 = h4.lp.algorithmic_bytes()
                sw = statistics.mean(p[0] for p in pr)
                rd = statistics.mean(sum(p) for p in pr)
                e4["roofline"] = dict(algorithmic_bytes=a4, sweep_us=sw * 1e3, sweep_frac=a4 / (sw * 1e-3) / 1e9 / peak,
                                      round_us=rd * 1e3, round_frac=a4 / (rd * 1e-3) / 1e9 / peak,
                                      note="one full round at the fixpoint bounds (638 MB: larger than L2), sweep = sweep_stream_kernel || sweep_long_kernel")
            if not args.no_parity:
                e4["parity"] = h4.parity()
            h4.close()
            del h4
            extras["c4"] = e4
        extras["c5_probing_batch"] = probing_batch_extra(torch, dist, propagator, synth, world, rank, local_rank,
                                                         with_cpu=not args.no_cpu)
        if rank == 0:
            line["extras"] = extras
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-extras", action="store_true", help="skip BASELINE configs[3] / configs[4] (and the reference's full-instance solve)")
    ap.add_argument("--no-parity", action="store_true", help="skip the comparison of the final bounds with the CPU oracle")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
