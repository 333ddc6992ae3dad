#!/bin/bash
# tools/gpu_ncu2.sh TAG CASE [CASE ...]: `ncu --set full` of the front-end kernels of each named case
# (tools/ncu_case.py).  Each case first runs once WITHOUT ncu and must exit 0 (B200_PROFILING.md); the first
# repetition is skipped (-s = launches of one repetition).  The reports are turned into the raw-metric and
# per-instruction CSV pages on the box (gpurun brings back at most 64 MiB) and the .ncu-rep is dropped unless
# KEEP_REP=1.
TAG=$1; shift
O=gpurun_out; mkdir -p $O
for c in "$@"; do
  n=1; [ "$c" = "step3" ] && n=3
  timeout 300 python tools/ncu_case.py $c > $O/${TAG}_${c}_plain.log 2>&1 &&
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_front -s $n -c $n -o $O/${TAG}_$c -f \
    python tools/ncu_case.py $c > $O/${TAG}_${c}_ncu.log 2>&1
  echo "$c: rc=$? $(tail -1 $O/${TAG}_${c}_plain.log)"
  ncu -i $O/${TAG}_$c.ncu-rep --page raw --csv > $O/${TAG}_${c}_raw.csv 2>/dev/null
  ncu -i $O/${TAG}_$c.ncu-rep --page source --csv > $O/${TAG}_${c}_source.csv 2>/dev/null
  [ "$KEEP_REP" = "1" ] || rm -f $O/${TAG}_$c.ncu-rep
done
ls -la $O/${TAG}_*
