#!/bin/bash
# BASELINE configs[3] and configs[4] at full size on 8 GPUs (run with gpurun --gpus 8)
O=gpurun_out; mkdir -p $O
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29561 \
  bench.py --gpus 8 --clips 512 --steps 3 --warmup 3 --no-e2e --no-cpu > $O/cfg4_8gpu.json 2> $O/cfg4_8gpu.err; echo "cfg4 rc=$?"
tail -2 $O/cfg4_8gpu.err | cut -c1-300
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29562 \
  tools/bench_config5.py --jobs-per-gpu 64 > $O/cfg5_8gpu.json 2> $O/cfg5_8gpu.err; echo "cfg5 rc=$?"
tail -2 $O/cfg5_8gpu.err | cut -c1-300
grep -h "^{" $O/cfg4_8gpu.json $O/cfg5_8gpu.json | cut -c1-900
