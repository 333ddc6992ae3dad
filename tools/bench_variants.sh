#!/bin/bash
# run on the GPU box: quick device-resident bench of every library variant (per-kernel ms)
O=gpurun_out; mkdir -p $O
for so in audio_tabs_b200/lib/variants/*.so; do
  n=$(basename $so .so)
  B200SPEC_LIB=$PWD/$so timeout 300 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu --no-configs > $O/var_$n.json 2> $O/var_$n.err
  python - "$n" $O/var_$n.json <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[2]))
    print(sys.argv[1], "step %.2f ms"%d["ms_per_step"], " ".join("%s:%.3f"%(k["frame_size"],k["ms"]) for k in d["roofline"]["per_kernel"]))
except Exception as e:
    print(sys.argv[1],'FAILED',e)
PY
done
