#!/usr/bin/env python
"""bench.py -- activity-based linear bound propagation to the fixpoint on B200(s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c3|c3small|c4]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

One step = one propagation of the workload from its initial bounds to the fixpoint.  Workload at N=1: BASELINE.json
configs[2], the synthetic set-cover MIP 1M rows x 1M binaries, 10M nonzeros (seed 1) -- the configuration the headline
metric is quoted on.  At N>1 the same instance is row-sharded over the ranks (strong scaling), one int64 MIN all-reduce
of the candidate keys per round (NCCL).

metric  propagation_fixpoint_nnz_per_s = nonzeros of the instance / time to reach the propagation fixpoint
value   bounds resident in HBM when the timed region starts (device-side reset of the bounds inside it)
e2e     the same through the C ABI with HOST buffers, as the plugin calls it at a node: H2D of lb/ub (pinned), fixpoint,
        D2H of the verdict + the round-ordered change log (N > 1: of the bound vectors), every step
roofline  the dominant kernel (the filter sweep of one full round): algorithmic bytes (nnz*12 + nrows*20 + ncols*17,
        SURVEY.md 8d) / its CUDA-event time, against the measured HBM copy peak of MEASURED_PEAKS.json
cpu_baseline / --impl reference   the UNMODIFIED reference (oracle/_ref, SCIP's cons_linear propagation) on the host
        cores, on a bounded sample of the same workload family (1 thread: SCIP is single threaded)
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
# bounded sample of the c3 family for the CPU arm (about 10-30 s of host work per step incl. model construction)
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
# CPU arm: the unmodified reference (oracle/_ref) or, if it is not built, the C restatement (oracle/liboracle.so)
# ---------------------------------------------------------------------------------------------------------------------
def cpu_reference_steps(steps, warmup):
    import oracle
    from scip_b200 import synth
    from scip_b200.lpb import write_lpb
    prob = CPU_SAMPLE["gen"](synth)
    nnz = len(prob["vals"])
    cores_avail = os.cpu_count()
    times = []
    if oracle.have_reference():
        kind = "reference"
        with tempfile.TemporaryDirectory() as tmp:
            path = os.path.join(tmp, "sample.lpb")
            write_lpb(path, prob)
            for i in range(warmup + steps):
                res = oracle.run_reference(path)
                if i >= warmup:
                    times.append(float(res["prop_time_s"]))
        what = "SCIPconshdlrGetPropTime(linear) of SCIP 11 built from /root/reference (oracle/_ref), parity settings"
    else:
        kind = "port"
        oracle.build()
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            oracle.propagate(prob)
            if i >= warmup:
                times.append(time.perf_counter() - t0)
        what = "oracle/linprop_oracle.c (C restatement, Jacobi rounds)"
    t = statistics.mean(times)
    return dict(value=nnz / t, unit=UNIT, cores=1, cores_available=cores_avail, kind=kind,
                sample=f"{CPU_SAMPLE['desc']}; {what}; fixpoint {t * 1e3:.1f} ms"), t, nnz


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cb, t, nnz = cpu_reference_steps(args.steps, args.warmup)
    line = dict(metric=METRIC, value=cb["value"], unit=UNIT, n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
                ms_per_step=t * 1e3, higher_is_better=True, scaling="strong", vs_baseline=None, dtype="f64",
                data="synthetic", impl="reference",
                config=dict(workload=CPU_SAMPLE["desc"], full_workload=WORKLOADS[args.workload]["desc"]),
                cpu_baseline=cb,
                e2e=dict(value=cb["value"], unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(line))


def probing_batch_extra(propagator, synth, device, with_cpu=True):
    """BASELINE configs[4] on one GPU (reported beside the headline, not part of it): 1024 probes -- one free binary fixed
    to 0 or 1 each, SCIPapplyProbingVar's pattern -- on the 5M-nnz set-cover matrix at its root fixpoint, 64 workers, one
    launch per worker (its probes run one after the other, each inside one block); wall clock of the whole batch through
    the C ABI with host buffers"""
    prob = synth.setcover(500_000, 500_000, 5_000_000, seed=3)
    with propagator.LinearPropagator(prob, device=device) as base:
        base.propagate()
        lb, ub = base.get_bounds()
        free = np.flatnonzero(lb < ub)
        rng = np.random.default_rng(3)
        var = free[rng.integers(0, len(free), size=1024)].astype(np.int32)
        val = rng.integers(0, 2, size=1024).astype(np.float64)
        base.probe_batch(var[:256], val[:256], val[:256], nworkers=64)
        times = []
        for _ in range(5):
            t0 = time.perf_counter()
            res = base.probe_batch(var, val, val, nworkers=64)
            times.append(time.perf_counter() - t0)
    t = min(times)
    out = dict(workload="1024 probing bound vectors on a 5M-nnz set-cover MIP (500k x 500k, seed 3), 64 workers",
               ms_per_batch=t * 1e3, us_per_probe=t / len(var) * 1e6, probes=int(len(var)),
               cutoffs=int((res["status"] == 1).sum()), mean_rounds=float(res["nrounds"].mean()),
               mean_changes=float(res["nchanges"].mean()))
    if with_cpu:
        out["cpu_reference"] = cpu_probing_reference(synth)
    return out


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
def run_ours(args):
    import torch
    import torch.distributed as dist
    from scip_b200 import build, propagator, sharded, synth

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
    prob = wl["gen"](synth)
    nrows, ncols, nnz = len(prob["lhs"]), len(prob["lb"]), len(prob["vals"])
    K, W = args.steps, args.warmup

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # a created stream: the library captures its round loop into a CUDA graph, which the legacy default stream forbids
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    lb0 = torch.from_numpy(prob["lb"] + 0.0)
    ub0 = torch.from_numpy(prob["ub"] + 0.0)
    d_lb0, d_ub0 = lb0.cuda(), ub0.cuda()                      # resident initial bounds (device-timed arm)
    h_lb, h_ub = lb0.clone().pin_memory(), ub0.clone().pin_memory()   # pinned host buffers (end-to-end arm)
    h_olb, h_oub = torch.empty_like(h_lb).pin_memory(), torch.empty_like(h_ub).pin_memory()

    result = {}
    if world == 1:
        lp = propagator.LinearPropagator(prob, device=local_rank)
        lp.set_stream(stream.cuda_stream)
        abytes = lp.algorithmic_bytes()
        # the call the plugin makes at a node (prop_gpulinear.c): bounds in, verdict + round-ordered change log out; the
        # log is on in both arms (it is part of the product path)
        logcap = 2 * ncols
        lp.set_change_log(logcap)
        h_chg = torch.empty(logcap * 24, dtype=torch.uint8).pin_memory()
        d2h_bytes = [0]

        def step_resident():
            lp.set_bounds_ptr(d_lb0.data_ptr(), d_ub0.data_ptr(), on_device=True)
            return lp.propagate(0)

        def step_e2e():
            lp.set_bounds_ptr(h_lb.data_ptr(), h_ub.data_ptr(), on_device=False)
            res = lp.propagate(0)                                                    # verdict: 32 bytes, synchronises
            d2h_bytes[0] = 24 * min(lp.changes_ptr(h_chg.data_ptr(), logcap), logcap) + 32
            return res
        launches_per_step = lambda res: 1 + lp.call_stats()["launches"]             # noqa: E731  (set_bounds + the call)
    elif args.exchange == "peer":
        # candidates are committed into every rank's key vector through NVLink peer memory by the kernel that produces
        # them; device-side barriers; the round loop stays in the CUDA graph on every GPU
        sp = sharded.PeerPropagator(prob, rank, world, local_rank)
        lp = sp.lp
        lp.set_stream(stream.cuda_stream)
        abytes = nnz * 12 + nrows * 20 + ncols * 17

        def step_resident():
            lp.set_bounds_ptr(d_lb0.data_ptr(), d_ub0.data_ptr(), on_device=True)
            return lp.propagate(0)

        def step_e2e():
            lp.set_bounds_ptr(h_lb.data_ptr(), h_ub.data_ptr(), on_device=False)
            res = lp.propagate(0)
            lp.get_bounds_ptr(h_olb.data_ptr(), h_oub.data_ptr(), on_device=False)
            return res
        launches_per_step = lambda res: 1 + lp.call_stats()["launches"]             # noqa: E731
    else:
        cuts = sharded.partition_rows(prob["rowptr"], world)
        eng = sharded.CudaEngine(prob, (int(cuts[rank]), int(cuts[rank + 1])), local_rank)
        sp = sharded.ShardedPropagator(eng)
        lp = eng.lp
        abytes = nnz * 12 + nrows * 20 + ncols * 17

        def step_resident():
            lp.set_bounds_ptr(d_lb0.data_ptr(), d_ub0.data_ptr(), on_device=True)
            return sp.propagate(0)

        def step_e2e():
            lp.set_bounds_ptr(h_lb.data_ptr(), h_ub.data_ptr(), on_device=False)
            res = sp.propagate(0)
            lp.get_bounds_ptr(h_olb.data_ptr(), h_oub.data_ptr(), on_device=False)
            return res
        launches_per_step = lambda res: 2 + 6 * res["nrounds"]                       # noqa: E731

    def timed(stepfn):
        for _ in range(W):
            res = stepfn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(K):
            res = stepfn()
        e1.record(stream)
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, res

    with ClockSampler(local_rank) as clocks:
        ms_res, res = timed(step_resident)
        ms_e2e, res2 = timed(step_e2e)
        # the dominant kernel: filter sweep of a full round at the fixpoint bounds (all rows marked, nothing changes)
        prof = []
        if world == 1:
            # the sweep of c3 moves 89 MB, less than the 126 MB L2: flush it between the profiled rounds
            l2flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
            prof_warm = []
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
            round_stats = lp.round_stats()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    status_ok = res["status"] == res2["status"]
    peak, peak_src = measured_peak_gbs()
    line = dict(metric=METRIC, value=nnz * K / (ms_res * 1e-3), unit=UNIT, n_gpus=world, steps=K, warmup=W,
                ms_per_step=ms_res / K, higher_is_better=True, scaling="strong", vs_baseline=None, dtype="f64",
                data="synthetic",
                config=dict(workload=wl["desc"], nrows=nrows, ncols=ncols, nnz=nnz, rounds=res["nrounds"],
                            changes=res["nchanges"], verdict=propagator.STATUS_NAMES[res["status"]],
                            parallelism=("1 GPU" if world == 1 else
                                         (f"rows sharded over {world} GPUs, candidates committed to all ranks through NVLink peer memory "
                                          "inside the exact kernel, 2 device barriers/round" if args.exchange == "peer" else
                                          f"rows sharded over {world} GPUs, 1 int64 MIN all-reduce/round (NCCL)")),
                            l2=("a step touches more than the 126 MB L2 (matrix 120-600 MB, row and column arrays, CSC, change log); "
                                "the profiled rounds of `roofline` run after a 256 MB L2 flush each") if args.workload != "c3small" else "fits L2",
                            loop="CUDA graph WHILE node (device-side)" if (world == 1 or args.exchange == "peer") else "host loop, NCCL per round"),
                e2e=dict(value=nnz * K / (ms_e2e * 1e-3), unit=UNIT, ms_per_step=ms_e2e / K, h2d_bytes_per_step=16 * ncols,
                         d2h_bytes_per_step=(d2h_bytes[0] if world == 1 else 16 * ncols + 32),
                         result=("verdict + change log (gpulin_get_changes)" if world == 1 else "verdict + bound vectors")),
                gpu_launches=K * launches_per_step(res), clocks=clocks.summary(), status_consistent=status_ok)
    if world > 1 and args.exchange == "peer":
        ms, rn, rc = lp.round_stats()
        line["round_us"] = [round(float(x) * 1e3, 1) for x in ms]
        line["round_nnz_local"] = [int(x) for x in rn]
    if world == 1:
        sweep_ms = statistics.mean(p[0] for p in prof)
        exact_ms = statistics.mean(p[1] for p in prof)
        apply_ms = statistics.mean(p[2] for p in prof)
        achieved = abytes / (sweep_ms * 1e-3) / 1e9
        kname = {"c3": "sweep_sell_bits_kernel", "c3small": "sweep_sell_kernel", "c4": "sweep_stream_kernel"}[args.workload]
        traffic = measured_traffic(f"{kname}:{args.workload}")
        line["roofline"] = dict(bound="hbm", kernel=f"{kname} (filter sweep of one full round)", achieved=achieved,
                                peak=peak, unit="GB/s", frac=achieved / peak, peak_source=peak_src,
                                traffic=traffic,
                                # what actually crossed the DRAM pins (ncu) over the live kernel time: the unit rows of c3
                                # are swept without reading their values, so this is below `achieved` there
                                dram_frac=(traffic / (sweep_ms * 1e-3) / 1e9 / peak) if traffic else None,
                                algorithmic_bytes=abytes, kernel_us=sweep_ms * 1e3,
                                l2="flushed before every profiled round (256 MB written, 192 MB read back)",
                                warm_l2=dict(kernel_us=statistics.mean(p[0] for p in prof_warm) * 1e3,
                                             frac=abytes / (statistics.mean(p[0] for p in prof_warm) * 1e-3) / 1e9 / peak,
                                             note="rounds back to back as inside a fixpoint: what fits stays in L2"),
                                full_round_us=(sweep_ms + exact_ms + apply_ms) * 1e3,
                                full_round_frac=abytes / ((sweep_ms + exact_ms + apply_ms) * 1e-3) / 1e9 / peak)
        ms, rn, rc = round_stats
        line["rounds"] = dict(full_round_nnz_per_s=nnz / (sweep_ms * 1e-3),
                              profile_round_us=dict(sweep=sweep_ms * 1e3, exact=exact_ms * 1e3, apply=apply_ms * 1e3))
        if not args.no_cpu:
            cb, _, _ = cpu_reference_steps(1, 0)
            line["cpu_baseline"] = cb
        if not args.no_extras:
            line["extras"] = dict(c5_probing_batch=probing_batch_extra(propagator, synth, local_rank, with_cpu=not args.no_cpu))
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
    ap.add_argument("--no-extras", action="store_true", help="skip the extra measurements (BASELINE configs[4], N=1 only)")
    ap.add_argument("--exchange", default="peer", choices=["peer", "nccl"],
                    help="N>1: how the ranks merge candidate bounds (peer memory inside the kernel | NCCL all-reduce)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
