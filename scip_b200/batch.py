"""Batch of probing bound vectors on one matrix (BASELINE config 5, the SCIPapplyProbingVar pattern of
prop_probing.c:1254-1279): every probe starts from the propagated base node, changes the bounds of one (or a few)
variables and propagates to its own fixpoint.

On the device a probe is a *clone* of the base handle (``gpulin_clone``): it shares the matrix and owns a bound vector,
keys, marks and a CUDA graph on its own stream, so that many probes are in flight at once -- each of them touches only
the rows its changes reach, which is latency-bound work that overlaps well.  Across GPUs the probes are split between
the ranks (``probes[rank::world]``); there is no collective."""
from __future__ import annotations

import numpy as np

from .propagator import LinearPropagator


class ProbingBatch:
    def __init__(self, base: LinearPropagator, nworkers: int = 32):
        """``base`` must hold the bounds of the node (normally after ``base.propagate()``)"""
        self.base = base
        self.workers = [base.clone() for _ in range(max(1, nworkers))]

    def close(self):
        for w in self.workers:
            w.close()
        self.workers = []

    def run(self, var, lb, ub, maxrounds: int = 0, want_bounds: bool = False):
        """probe i sets variable var[i] to [lb[i], ub[i]].  Returns dict(status, nrounds, nchanges) as arrays over the
        probes (+ lists lb/ub of full bound vectors if ``want_bounds``)"""
        var = np.ascontiguousarray(var, dtype=np.int32)
        lb = np.ascontiguousarray(lb, dtype=np.float64)
        ub = np.ascontiguousarray(ub, dtype=np.float64)
        n = len(var)
        status = np.zeros(n, dtype=np.int32)
        nrounds = np.zeros(n, dtype=np.int32)
        nchanges = np.zeros(n, dtype=np.int64)
        out_lb, out_ub = ([None] * n, [None] * n) if want_bounds else (None, None)
        W = len(self.workers)
        inflight = [None] * W

        def finish(w):
            i = inflight[w]
            res = self.workers[w].propagate_wait()
            status[i], nrounds[i], nchanges[i] = res["status"], res["nrounds"], res["nchanges"]
            if want_bounds:
                out_lb[i], out_ub[i] = self.workers[w].get_bounds()
            inflight[w] = None

        for i in range(n):
            w = i % W
            if inflight[w] is not None:
                finish(w)
            wk = self.workers[w]
            wk.reset_from(self.base)                                   # backtrack to the node
            wk.update_bounds_nosync(var[i:i + 1], lb[i:i + 1], ub[i:i + 1])   # the probing bound change
            wk.propagate_async(maxrounds)
            inflight[w] = i
        for w in range(W):
            if inflight[w] is not None:
                finish(w)
        res = dict(status=status, nrounds=nrounds, nchanges=nchanges)
        if want_bounds:
            res["lb"], res["ub"] = out_lb, out_ub
        return res
