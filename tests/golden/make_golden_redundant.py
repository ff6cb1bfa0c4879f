#!/usr/bin/env python
"""Adds <name>.bs1e-9.red to tests/golden/: one byte per row of <name>.lpb, 1 if the UNMODIFIED reference (oracle/_ref)
removed that constraint during its root propagation -- propagateCons deletes the rows it finds redundant
(cons_linear.c:7743-7753).  Run in the build container (needs oracle/_ref):  python tests/golden/make_golden_redundant.py
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

import oracle  # noqa: E402


def main():
    for f in sorted(os.listdir(HERE)):
        if not f.endswith(".lpb"):
            continue
        name = f[:-4]
        out = os.path.join(HERE, name + ".bs1e-9.red")
        subprocess.run([oracle.REF_DRIVER, "--lpb", os.path.join(HERE, f), "--boundstreps", "1e-9", "--redundant", out],
                       check=True, capture_output=True, text=True, timeout=600)
        raw = open(out, "rb").read()
        print(name, "rows", len(raw), "removed by the reference", sum(raw))


if __name__ == "__main__":
    main()
