#!/usr/bin/env python
"""GPU probe: builds one synthetic config, runs the CUDA fixpoint, compares with the CPU oracle, prints per-round
statistics and the roofline fraction of the full first round.  Usage: python scripts/probe.py [c3|c4|c3small] [--no-oracle]"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scip_b200 import propagator, synth  # noqa: E402


def main():
    which = sys.argv[1] if len(sys.argv) > 1 else "c3"
    t0 = time.time()
    if which == "c3":
        prob = synth.setcover()
    elif which == "c3small":
        prob = synth.setcover(100_000, 100_000, 1_000_000)
    elif which == "c4":
        prob = synth.mixed_knapsack()
    elif which == "c4small":
        prob = synth.mixed_knapsack(20_000, 200_000, 5_000_000)
    else:
        raise SystemExit("unknown config")
    print(f"generated {prob['name']}: rows {len(prob['lhs'])} cols {len(prob['lb'])} nnz {len(prob['vals'])} in {time.time()-t0:.1f}s", flush=True)
    t0 = time.time()
    lp = propagator.LinearPropagator(prob)
    print(f"device build {time.time()-t0:.1f}s layout {lp.layout()}", flush=True)
    abytes = lp.algorithmic_bytes()
    for rep in range(int(os.environ.get("PROBE_REPS", "3"))):
        lp.set_bounds(prob["lb"], prob["ub"])
        res = lp.propagate()
        ms, nnz, nchg = lp.round_stats()
        print(json.dumps(dict(rep=rep, **res)))
        for i in range(len(ms)):
            print(f"   round {i}: {ms[i]*1e3:9.1f} us  nnz {nnz[i]:10d}  changes {nchg[i]:8d}  "
                  f"{nnz[i]/max(ms[i],1e-9)/1e6:8.1f} Gnnz/s")
    tr = lp.trace()
    print("trace: " + "  ".join(f"{name}@{t:.1f}" for name, t in tr[:80]))
    print("stages: " + "  ".join(f"{tr[i][0]}={tr[i + 1][1] - tr[i][1]:.1f}" for i in range(min(len(tr) - 1, 12))))
    print(f"algorithmic bytes/round {abytes/1e6:.1f} MB; first round {abytes/ms[0]/1e6:.1f} GB/s "
          f"= {abytes/ms[0]/1e6/6544.3:.3f} of measured HBM peak")
    lb, ub = lp.get_bounds()
    if "--no-oracle" not in sys.argv:
        import oracle
        t0 = time.time()
        want = oracle.propagate(prob)
        dt = time.time() - t0
        print(f"oracle: status {want['status']} rounds {want['nrounds']} changes {want['nchanges']} in {dt:.2f}s")
        it = prob["vartype"] != 0
        print("int mismatches:", int(((lb != want["lb"]) | (ub != want["ub"]))[it].sum()),
              "cont mismatches:", int(((lb != want["lb"]) | (ub != want["ub"]))[~it].sum()),
              "status equal:", res["status"] == want["status"], "rounds equal:", res["nrounds"] == want["nrounds"])


if __name__ == "__main__":
    main()
