#!/bin/bash
# multi-GPU bench (run with gpurun --gpus 8): N = 8 full line, N = 4 and 2 device-resident only
O=gpurun_out; mkdir -p $O
nvidia-smi --query-gpu=index,name --format=csv > $O/scale_smi.txt
for n in 8 4 2; do
  extra=""; [ $n -ne 8 ] && extra="--no-e2e"
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2954$n \
    bench.py --gpus $n --steps 10 --warmup 3 $extra > $O/scale_n$n.json 2> $O/scale_n$n.err; echo "n=$n rc=$?"
  python - $O/scale_n$n.json <<'PY'
import json,sys
try:
    d=json.loads([l for l in open(sys.argv[1]) if l.startswith('{')][-1])
    print('N=%d value %.0f ms/step %.3f e2e %s'%(d['n_gpus'], d['value'], d['ms_per_step'], d['e2e'] and round(d['e2e']['value'])))
except Exception as e: print('parse failed', e)
PY
done
