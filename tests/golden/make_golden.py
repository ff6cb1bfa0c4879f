#!/usr/bin/env python
"""Regenerates tests/golden/ from the UNMODIFIED reference (oracle/_ref, built by `make -C oracle ref`).

Run in the build container (needs /root/reference):   python tests/golden/make_golden.py

For every instance it stores
  <name>.lpb          the linear rows + bounds exactly as a propagator sees them at the first root propagation
                      call of the reference (dumped by oracle/ref_driver.c through SCIPgetVarsLinear & co.)
  <name>.bs1e-9.lpr   global bounds + verdict after the reference's own cons_linear propagation ran to its
                      fixpoint with numerics/boundstreps = 1e-9 (the exact-parity protocol)
  <name>.bs0.05.lpr   the same at the reference's default numerics/boundstreps = 0.05 (order dependent)
Instances: the check/instances/MIP/*.mps files the reference ships (semicon1 is skipped: it is not purely linear)
and small seeded synthetic problems of the BASELINE shapes.  p0548 stands in for the absent p0201.
"""
import glob
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

import oracle  # noqa: E402
from scip_b200 import synth  # noqa: E402
from scip_b200.lpb import write_lpb  # noqa: E402

REFDIR = "/root/reference/check/instances/MIP"


def main():
    manifest = {}
    jobs = []
    for f in sorted(glob.glob(os.path.join(REFDIR, "*.mps"))):
        name = os.path.basename(f)[:-4]
        if name == "semicon1":
            continue
        jobs.append((name, f, False))
    synth_probs = {
        "syn_setcover_2k": synth.setcover(2000, 2000, 20000, seed=11),
        "syn_setcover_2k_infeas": synth.setcover(2000, 2000, 20000, seed=12, infeasible=True),
        "syn_mixedknap_300": synth.mixed_knapsack(300, 3000, 30000, seed=13, dense_range=(300, 900), eq_frac=0.3),
        "syn_mixedknap_300_infeas": synth.mixed_knapsack(300, 3000, 30000, seed=14, dense_range=(300, 900),
                                                         infeasible=True),
        # numerics: cancellation after an unreliable activity update, three and more huge contributions
        "syn_edge_cancel": synth.edge_cancellation(),
        "syn_edge_huge": synth.edge_huge(),
        "syn_edge_huge_infeas": synth.edge_huge(infeasible=True),
    }
    # ranged-row (gcd) propagation, rangedRowPropagation cons_linear.c:5715-6696: these instances are run with
    # constraints/linear/rangedrowpropagation = TRUE (and rangedrowartcons = FALSE: the rule's branches that add constraints
    # are not bound propagation); <name>.tie = SCIPvarGetProbindex of every column, the last sort key of the row order
    # the rule walks (consdataCompVarProp :3191)
    ranged_probs = {
        "syn_ranged_400": synth.ranged_rows(400, 600, seed=31),
        "syn_ranged_400b": synth.ranged_rows(400, 600, seed=33),
        "syn_ranged_infeas": synth.ranged_rows(200, 300, seed=34, infeasible=True),
    }
    synth_probs.update(ranged_probs)
    for name, prob in synth_probs.items():
        path = os.path.join("/tmp", name + ".gen.lpb")
        write_lpb(path, prob)
        jobs.append((name, path, True))
    for name, path, is_lpb in jobs:
        entry = {}
        for bs in ("1e-9", "0.05"):
            ranged = name in ranged_probs
            res = oracle.run_reference(path, boundstreps=float(bs), is_lpb=is_lpb,
                                       dump_lpb=os.path.join(HERE, name + ".lpb"),
                                       out_lpr=os.path.join(HERE, f"{name}.bs{bs}.lpr"), rangedrow=ranged,
                                       dump_tie=os.path.join(HERE, name + ".tie") if ranged else None)
            entry[bs] = dict(infeasible=bool(res["infeasible"]), prop_calls=int(res["prop_calls"]),
                             domreds=int(res["domreds"]))
            if ranged:
                entry[bs]["rangedrow"] = True
        manifest[name] = entry
        print(name, entry)
    with open(os.path.join(HERE, "manifest.json"), "w") as f:
        json.dump(manifest, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
