#!/usr/bin/env python
"""Latency of incremental calls (the plugin's call at a branch-and-bound node): after the root fixpoint of C3, fix one free
binary, propagate, repeat.  Usage: [GPULIN_SMALL=0] python scripts/probe_dive.py [nsteps]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scip_b200 import propagator, synth  # noqa: E402


def main():
    nsteps = int(sys.argv[1]) if len(sys.argv) > 1 else 300
    prob = synth.setcover()
    lp = propagator.LinearPropagator(prob)
    lp.propagate()
    lb, ub = lp.get_bounds()
    free = np.flatnonzero(lb < ub)
    rng = np.random.default_rng(5)
    var = free[rng.permutation(len(free))[:nsteps]].astype(np.int32)
    val = np.zeros(nsteps)
    wall, dev, rounds, chg = [], [], [], []
    for i in range(nsteps):
        t0 = time.perf_counter()
        lp.update_bounds_nosync(var[i:i + 1], val[i:i + 1], val[i:i + 1])
        res = lp.propagate()
        wall.append(time.perf_counter() - t0)
        dev.append(res["device_ms"])
        rounds.append(res["nrounds"])
        chg.append(res["nchanges"])
        if res["status"] != 0:
            break
    wall, dev = np.array(wall[10:]), np.array(dev[10:])
    print(f"GPULIN_SMALL={os.environ.get('GPULIN_SMALL', '1')}: {len(wall)} incremental calls: wall median {np.median(wall)*1e6:.1f} us "
          f"(p90 {np.percentile(wall, 90)*1e6:.1f}), device median {np.median(dev)*1e3:.1f} us; mean rounds {np.mean(rounds):.2f}, "
          f"mean changes {np.mean(chg):.1f}, last status {res['status']}")


if __name__ == "__main__":
    main()
