#!/usr/bin/env python
"""Summarise `ncu -i X.ncu-rep --page source --csv` output per kernel launch:
opcode mix (warp-level instructions executed), stall samples by reason, the hottest SASS
instructions and shared-memory excess wavefronts.

    ncu -i prof.ncu-rep --page source --csv > src.csv
    python tools/ncu_source_summary.py src.csv [kernel-index] [--top N] [--per F]   (F: divide counts by F)
"""
import csv
import sys
from collections import Counter, defaultdict


def split_kernels(path):
    kernels, cur = [], None
    with open(path, newline="") as fh:
        for row in csv.reader(fh):
            if not row:
                continue
            if row[0] == "Kernel Name":
                cur = {"name": row[1], "hdr": None, "rows": []}
                kernels.append(cur)
            elif cur is not None and cur["hdr"] is None:
                cur["hdr"] = row
            elif cur is not None:
                cur["rows"].append(row)
    return kernels


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    path = args[0]
    which = int(args[1]) if len(args) > 1 else -1
    top = 25
    per = 1.0
    for i, a in enumerate(sys.argv):
        if a == "--top":
            top = int(sys.argv[i + 1])
        if a == "--per":
            per = float(sys.argv[i + 1])
    ks = split_kernels(path)
    print("%d kernel launches in file" % len(ks))
    k = ks[which]
    hdr = k["hdr"]
    col = {h: i for i, h in enumerate(hdr)}
    ops, samples_by_op = Counter(), Counter()
    stalls = Counter()
    stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    total_inst = 0
    excess = Counter()
    rows = []
    for r in k["rows"]:
        src = r[col["Source"]].strip()
        toks = src.split()
        op = toks[1] if toks and toks[0].startswith("@") and len(toks) > 1 else (toks[0] if toks else "?")
        base = op.split(".")[0]
        n = float(r[col["Instructions Executed"]] or 0)
        s = float(r[col["# Samples"]] or 0)
        ops[base] += n
        samples_by_op[base] += s
        total_inst += n
        for sc in stall_cols:
            stalls[sc] += float(r[col[sc]] or 0)
        ex = float(r[col["L1 Wavefronts Shared Excessive"]] or 0)
        if ex:
            excess[src] += ex
        rows.append((s, n, src, r))
    print("kernel:", k["name"], " SASS instructions:", len(rows))
    print("total warp instructions executed: %.0f  (per unit: %.1f)" % (total_inst, total_inst / per))
    print("\n-- opcode mix (warp instr, per unit, %% of total, stall samples) --")
    for op, n in ops.most_common(28):
        print("%-12s %12.1f %6.2f%%  samples %8.0f" % (op, n / per, 100 * n / total_inst, samples_by_op[op]))
    tot_s = sum(stalls.values())
    print("\n-- stall samples by reason (total %.0f) --" % tot_s)
    for sc, v in stalls.most_common():
        if v:
            print("%-26s %9.0f %6.2f%%" % (sc, v, 100 * v / tot_s))
    print("\n-- top %d instructions by samples --" % top)
    rows.sort(key=lambda t: -t[0])
    for s, n, src, r in rows[:top]:
        reasons = sorted(((float(r[col[sc]] or 0), sc) for sc in stall_cols), reverse=True)[:2]
        print("%7.0f  exec %10.1f  %-60s %s" % (s, n / per, src[:60], ", ".join("%s=%.0f" % (b, a) for a, b in reasons)))
    if excess:
        print("\n-- shared-memory excessive wavefronts (bank conflicts) --")
        for src, ex in excess.most_common(12):
            print("%12.1f  %s" % (ex / per, src[:80]))


if __name__ == "__main__":
    main()
