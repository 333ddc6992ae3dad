#!/usr/bin/env python
"""Split a kernel's SASS (ncu source page csv) at BAR instructions and report, per phase,
warp instructions executed and stall samples.  usage: ncu_phase_split.py src.csv kernel_idx per"""
import sys
sys.path.insert(0, __file__.rsplit('/', 1)[0])
from ncu_source_summary import split_kernels

k = split_kernels(sys.argv[1])[int(sys.argv[2])]
per = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
col = {h: i for i, h in enumerate(k['hdr'])}
seg, cur = [], {'n': 0, 's': 0, 'ops': {}}
for r in k['rows']:
    src = r[col['Source']].strip()
    toks = src.split()
    op = toks[1] if toks[0].startswith('@') else toks[0]
    n = float(r[col['Instructions Executed']] or 0)
    s = float(r[col['# Samples']] or 0)
    cur['n'] += n
    cur['s'] += s
    b = op.split('.')[0]
    cur['ops'][b] = cur['ops'].get(b, 0) + n
    if op.startswith('BAR'):
        seg.append(cur)
        cur = {'n': 0, 's': 0, 'ops': {}}
seg.append(cur)
for i, c in enumerate(seg):
    top = sorted(c['ops'].items(), key=lambda t: -t[1])[:8]
    print(i, 'instr/unit %.0f samples %.0f' % (c['n'] / per, c['s']), ' '.join('%s=%.0f' % (a, b / per) for a, b in top))
