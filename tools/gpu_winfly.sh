#!/bin/bash
# parity tests, then A/B of the in-register Hann window against the table (same library, env switch)
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
B200SPEC_WINFLY=0 timeout 900 python -m pytest tests/test_gpu_frontends.py -m gpu -x -q 2>&1 | tail -2
for rep in 1 2; do for v in 0 1; do
  B200SPEC_WINFLY=$v timeout 300 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu > $O/wf_$v.json 2> $O/wf_$v.err
  python - $v $O/wf_$v.json <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[2])); print('winfly='+sys.argv[1], 'step %.3f ms'%d['ms_per_step'], ' '.join('%d:%.3f'%(k['frame_size'],k['ms']) for k in d['roofline']['per_kernel']))
except Exception as e: print('failed', e)
PY
done; done
