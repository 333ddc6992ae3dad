#!/usr/bin/env python
"""Per-phase (split at BAR) instruction counts and stall reasons of one kernel in an ncu source-page csv.
usage: ncu_phase_stalls.py src.csv kernel_idx units"""
import sys
sys.path.insert(0, __file__.rsplit('/', 1)[0])
from ncu_source_summary import split_kernels
k = split_kernels(sys.argv[1])[int(sys.argv[2])]
per = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
col = {h: i for i, h in enumerate(k['hdr'])}
stall_cols = [h for h in k['hdr'] if h.startswith('stall_') and 'Not Issued' not in h]
seg, cur = [], {'n': 0, 's': 0, 'st': {}}
for r in k['rows']:
    toks = r[col['Source']].strip().split()
    op = toks[1] if toks[0].startswith('@') else toks[0]
    cur['n'] += float(r[col['Instructions Executed']] or 0)
    cur['s'] += float(r[col['# Samples']] or 0)
    for h in stall_cols:
        v = float(r[col[h]] or 0)
        if v:
            cur['st'][h] = cur['st'].get(h, 0) + v
    if op.startswith('BAR'):
        seg.append(cur)
        cur = {'n': 0, 's': 0, 'st': {}}
seg.append(cur)
tot_s = sum(c['s'] for c in seg)
tot_n = sum(c['n'] for c in seg)
print('kernel', k['name'][:60], 'instr/unit %.0f' % (tot_n / per))
for i, c in enumerate(seg):
    if c['s'] < 0.005 * tot_s:
        continue
    tot = sum(c['st'].values()) or 1
    print('phase %d: instr/unit %.0f (%.0f%%) time %.1f%%' % (i, c['n'] / per, 100 * c['n'] / tot_n, 100 * c['s'] / tot_s),
          ' '.join('%s=%.0f%%' % (a.replace('stall_', ''), 100 * b / tot) for a, b in sorted(c['st'].items(), key=lambda t: -t[1])[:7]))
