#!/usr/bin/env python
"""Shared-memory wavefronts per phase (split at the group barriers) and opcode, from
`ncu -i X.ncu-rep --page source --csv`:
    python tools/ncu_smem_phases.py src.csv kernel_idx units
The l1tex data pipe is the busiest unit of the front-end kernels, so this is the table to shrink."""
import csv
import sys

sys.path.insert(0, __file__.rsplit("/", 1)[0])
csv.field_size_limit(10**9)
from ncu_source_summary import split_kernels  # noqa: E402

k = split_kernels(sys.argv[1])[int(sys.argv[2])]
per = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
col = {h: i for i, h in enumerate(k["hdr"])}
rows = k["rows"]
W, X, N = "L1 Wavefronts Shared", "L1 Wavefronts Shared Excessive", "Instructions Executed"
G = next((c for c in col if c.startswith("L2 Theoretical Sectors Global")), None)
reg, acc, tot = 0, {}, [0.0, 0.0]
for r in rows:
    s = r[col["Source"]]
    if "BAR.SYNC" in s:
        reg += 1
    toks = s.split()
    op = toks[1] if toks[0].startswith("@") else toks[0]
    w, ex, n = (float(r[col[c]] or 0) for c in (W, X, N))
    if w or op.startswith(("LDG", "STG")):
        a = acc.setdefault((reg, op.split(".")[0] + ("." + op.split(".")[-1] if op[-1].isdigit() else "")), [0, 0, 0])
        a[0] += w
        a[1] += ex
        a[2] += n
    tot[0] += w
    tot[1] += ex
print(k["name"][:70])
print("shared wavefronts per unit %.1f (excess %.1f)" % (tot[0] / per, tot[1] / per))
for key in sorted(acc):
    a = acc[key]
    print("phase %2d %-8s wavefronts %7.1f  excess %6.1f  instr %7.1f" % (key[0], key[1], a[0] / per, a[1] / per, a[2] / per))
