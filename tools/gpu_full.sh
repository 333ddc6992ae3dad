#!/bin/bash
# Full GPU check of the current tree (run through gpurun): parity tests, bench (both arms), ncu launch
# list of the bench command and one `ncu --set full` capture of the k_front launches (reduced batch).
# usage: tools/gpu_full.sh TAG
TAG=${1:-cur}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $O/${TAG}_smi.txt
timeout 900 python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/${TAG}_pytest.log
tail -3 $O/${TAG}_pytest.log
timeout 600 python bench.py > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err; echo "bench rc=$?"
cat $O/${TAG}_bench.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/${TAG}_bench_ref.json 2>> $O/${TAG}_bench.err; echo "ref rc=$?"
cat $O/${TAG}_bench_ref.json
# launch list of the bench command: our kernels only (the synthetic-input generator's torch kernels are skipped)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -c 400 --csv --log-file $O/${TAG}_launches.csv \
  python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu > $O/${TAG}_ncu_launch.log 2>&1; echo "ncu launches rc=$?"
# full capture of the three k_front launches of one step at the FULL bench size (traffic per launch)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_front -c 3 -o $O/${TAG}_prof -f \
  python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu > $O/${TAG}_ncu_full.log 2>&1; echo "ncu full rc=$?"
ls -la $O | tail -8
