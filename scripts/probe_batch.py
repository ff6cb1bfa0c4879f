#!/usr/bin/env python
"""BASELINE config 5: 1024 probing bound vectors on one 5M-nnz set-cover matrix (500k x 500k, seed 3).
Prints the batch wall time for several worker counts and, with --ranks N, what one of N ranks would do."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scip_b200 import propagator, synth  # noqa: E402
from scip_b200.batch import ProbingBatch  # noqa: E402


def main():
    prob = synth.setcover(500_000, 500_000, 5_000_000, seed=3)
    lbv, ubv, var, val = synth.probing_batch(prob, nvec=1024, seed=3)
    del lbv, ubv
    base = propagator.LinearPropagator(prob)
    node = base.propagate()
    nlb, nub = base.get_bounds()
    print(f"node: {node}")
    still = nlb[var] < nub[var]
    var, val = var[still].astype(np.int32), val[still].astype(np.float64)
    print(f"{len(var)} of 1024 probe variables are still free at the node")
    for nworkers in (1, 32):
        pb = ProbingBatch(base, nworkers=nworkers)
        pb.run(var[:64], val[:64], val[:64])
        t0 = time.perf_counter()
        res = pb.run(var, val, val)
        dt = time.perf_counter() - t0
        pb.close()
        print(f"workers {nworkers:3d}: {dt*1e3:8.1f} ms for {len(var)} probes = {dt/len(var)*1e6:7.1f} us/probe; "
              f"cutoffs {int((res['status']==1).sum())}, mean rounds {res['nrounds'].mean():.2f}, "
              f"mean changes {res['nchanges'].mean():.1f}")


    for nworkers in (1, 8, 32, 64, 128, 256):
        base.probe_batch(var[:256], val[:256], val[:256], nworkers=nworkers)
        t0 = time.perf_counter()
        res = base.probe_batch(var, val, val, nworkers=nworkers)
        dt = time.perf_counter() - t0
        print(f"native, workers {nworkers:3d}: {dt*1e3:8.1f} ms for {len(var)} probes = {dt/len(var)*1e6:7.1f} us/probe; "
              f"cutoffs {int((res['status']==1).sum())}, mean rounds {res['nrounds'].mean():.2f}")


if __name__ == "__main__":
    main()
