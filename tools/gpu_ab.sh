#!/bin/bash
# parity tests with the default (pair) kernel, then A/B bench: one-frame kernel vs pair kernel
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -12
for v in 0 1; do
  B200SPEC_PAIR=$v timeout 300 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu > $O/ab_$v.json 2> $O/ab_$v.err; echo "pair=$v rc=$?"
  python - $O/ab_$v.json <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); print('step %.3f ms  value %.0f'%(d['ms_per_step'], d['value']), ' '.join('%d:%.3f'%(k['frame_size'],k['ms']) for k in d['roofline']['per_kernel']))
except Exception as e: print('failed', e)
PY
done
