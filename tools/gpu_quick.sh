#!/bin/bash
# quick GPU check: parity tests + device-resident bench of the current build
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
timeout 300 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu --no-configs > $O/quick_bench.json 2> $O/quick_bench.err; echo "bench rc=$?"
tail -3 $O/quick_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/quick_bench.json'))
print('step %.3f ms  value %.0f'%(d['ms_per_step'], d['value']), ' '.join('%d:%.3f'%(k['frame_size'],k['ms']) for k in d['roofline']['per_kernel']))
PY
