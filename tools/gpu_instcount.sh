#!/bin/bash
# warp-instruction counts + durations of the k_front launches of one reduced step (cheap ncu pass)
O=gpurun_out; mkdir -p $O
ARGS="--steps 1 --warmup 0 --clips 32 --clip-seconds 120 --no-e2e --no-cpu"
timeout 300 python bench.py $ARGS > $O/ic_plain.log 2>&1 &&
timeout 600 ncu --metrics smsp__inst_executed.sum,gpu__time_duration.sum,l1tex__data_pipe_lsu_wavefronts.sum --clock-control none -k regex:k_front -c 3 --csv --log-file $O/instcount.csv \
  python bench.py $ARGS > $O/ic_ncu.log 2>&1; echo "rc=$?"
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/instcount.csv')) if len(r)>10]
h=rows[0]; kn=h.index('Kernel Name'); mn=h.index('Metric Name'); mv=h.index('Metric Value')
for r in rows[1:]:
    print(r[kn][:40], r[mn], r[mv])
PY
