#!/bin/bash
# N = 8 bench line (run with gpurun --gpus 8)
O=gpurun_out; mkdir -p $O
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29548 \
  bench.py --gpus 8 --steps 10 --warmup 3 > $O/scale8.json 2> $O/scale8.err; echo "rc=$?"
tail -2 $O/scale8.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/scale8.json') if l.startswith('{')][-1])
print('N=%d value %.0f ms/step %.3f e2e %s numa cpus %s'%(d['n_gpus'], d['value'], d['ms_per_step'], round(d['e2e']['value']), d['e2e'].get('numa_bound_cpus')))
PY
