#!/bin/bash
# ncu --set full of the three k_front launches of one reduced step (32 x 120 s): tools/gpu_ncu.sh TAG
TAG=${1:-cur}
O=gpurun_out; mkdir -p $O
ARGS="--steps 1 --warmup 0 --clips 32 --clip-seconds 120 --no-e2e --no-cpu"
timeout 300 python bench.py $ARGS > $O/${TAG}_plain.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_front -c 3 -o $O/${TAG}_prof -f \
  python bench.py $ARGS > $O/${TAG}_ncu_full.log 2>&1; echo "ncu full rc=$?"
ls -la $O/${TAG}_prof.ncu-rep
