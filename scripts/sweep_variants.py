#!/usr/bin/env python
"""Times one full round (all rows marked, bounds at the fixpoint AND at the initial bounds) for several kernel variants
selected through environment variables.  Usage: python scripts/sweep_variants.py c3 "GPULIN_SELLBITS_VARIANT=0" "GPULIN_SELLBITS=0" ...
Each variant runs in a fresh process (the library reads the variables at gpulin_create)."""
import json
import os
import subprocess
import sys

CHILD = r'''
import sys, os, json, statistics
sys.path.insert(0, os.getcwd())
from scip_b200 import propagator, synth
which = sys.argv[1]
prob = {"c3": synth.setcover, "c4": synth.mixed_knapsack,
        "c3small": lambda: synth.setcover(100_000, 100_000, 1_000_000)}[which]()
lp = propagator.LinearPropagator(prob)
lp.set_bounds(prob["lb"], prob["ub"])
res = lp.propagate()
ms, nnz, nchg = lp.round_stats()
prof = [lp.profile_round() for _ in range(13)][3:]
out = dict(fixpoint_ms=res["device_ms"], rounds=res["nrounds"], changes=res["nchanges"], round0_us=ms[0] * 1e3,
           sweep_us=statistics.mean(p[0] for p in prof) * 1e3, exact_us=statistics.mean(p[1] for p in prof) * 1e3,
           apply_us=statistics.mean(p[2] for p in prof) * 1e3, abytes=lp.algorithmic_bytes())
fix = []
for _ in range(5):
    lp.set_bounds(prob["lb"], prob["ub"])
    fix.append(lp.propagate()["device_ms"])
out["fixpoint_ms_best"] = min(fix)
print(json.dumps(out))
'''


def main():
    which = sys.argv[1]
    for spec in sys.argv[2:] or [""]:
        env = dict(os.environ)
        for kv in spec.split():
            k, v = kv.split("=")
            env[k] = v
        r = subprocess.run([sys.executable, "-c", CHILD, which], env=env, capture_output=True, text=True, timeout=300)
        line = r.stdout.strip().splitlines()[-1] if r.stdout.strip() else r.stderr[-400:]
        try:
            d = json.loads(line)
            frac = d["abytes"] / (d["sweep_us"] * 1e-6) / 1e9 / 6544.3
            print(f"{spec or '(default)':40s} sweep {d['sweep_us']:6.1f} us ({frac:.3f})  exact {d['exact_us']:5.1f}  apply {d['apply_us']:5.1f}  "
                  f"round0 {d['round0_us']:6.1f}  fixpoint {d['fixpoint_ms_best']:.3f} ms  rounds {d['rounds']} changes {d['changes']}", flush=True)
        except Exception:
            print(spec, "FAILED", line, flush=True)


if __name__ == "__main__":
    main()
