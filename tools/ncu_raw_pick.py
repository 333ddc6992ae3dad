#!/usr/bin/env python
"""Print selected metrics of every kernel in an `ncu --page raw --csv` dump: tools/ncu_raw_pick.py raw.csv [regex]"""
import csv, re, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[0]
pat = re.compile(sys.argv[2]) if len(sys.argv) > 2 else None
keys = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'launch__registers_per_thread',
        'launch__shared_mem_per_block_dynamic', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct', 'sm__cycles_elapsed.max',
        'l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed', 'l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum', 'l1tex__f_wavefronts.sum', 'l1tex__data_pipe_lsu_wavefronts.sum',
        'smsp__average_warp_latency_issue_stalled_barrier.ratio', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fmalite.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'smsp__inst_executed_pipe_fma.sum']
if pat:
    keys = [h for h in hdr if pat.search(h)]
names = [r[hdr.index('Kernel Name')][:40] for r in rows[2:]]
print('%-75s' % 'metric', *['%22s' % n[-22:] for n in names])
for k in keys:
    if k in hdr:
        i = hdr.index(k)
        print('%-75s' % k[:75], *['%22s' % r[i] for r in rows[2:]], rows[1][i])
