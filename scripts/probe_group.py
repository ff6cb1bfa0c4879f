"""Several GPUs driven from ONE process (gpulin_group_connect): per-round device times of a fixpoint on every rank -- the
round, the part before its exchange (sweeps + exact rules of the rank's share), the wait for the slowest peer.
Usage: python scripts/probe_group.py c3|c4 [ngpus] [steps]"""
import sys
import os

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scip_b200 import sharded, synth  # noqa: E402

WL = {"c3": lambda: synth.setcover(1_000_000, 1_000_000, 10_000_000, seed=1),
      "c3small": lambda: synth.setcover(100_000, 100_000, 1_000_000, seed=1),
      "c4": lambda: synth.mixed_knapsack(200_000, 2_000_000, 50_000_000, seed=2)}


def main():
    wl = sys.argv[1] if len(sys.argv) > 1 else "c3"
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    steps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
    prob = WL[wl]()
    gp = sharded.GroupPropagator(prob, list(range(n)))
    for lp in gp.lps:
        lp.set_change_log(2 * len(prob["lb"]))
    for step in range(steps):
        gp.set_bounds(prob["lb"], prob["ub"])
        res = gp.propagate(0)
    print(f"{wl} on {n} GPUs: {res}")
    for r, lp in enumerate(gp.lps):
        ms, nnz, nchg = lp.round_stats()
        before, wait = lp.exchange_stats()
        print(f"rank {r}: fixpoint {ms.sum() * 1e3:.1f} us")
        for i in range(len(ms)):
            ex = f"before exchange {before[i] * 1e3:6.1f}  wait {wait[i] * 1e3:6.1f}" if before[i] >= 0 else "redundant (no exchange)"
            print(f"   round {i}: {ms[i] * 1e3:7.1f} us  nnz {nnz[i]:9d}  changes {nchg[i]:7d}  {ex}")
        tr = lp.trace()
        print("   trace: " + "  ".join(f"{name}@{t:.1f}" for name, t in tr[:60]))
    gp.close()


if __name__ == "__main__":
    main()
