#!/bin/bash
# run on the GPU box with a tuning build (tools/build_variant.sh tuning "-DB200SPEC_TUNING"): task size sweep and
# the warm-up-rows A/B (B200SPEC_SEAM=0), all in ONE call so that the numbers share a box
O=gpurun_out; mkdir -p $O
export B200SPEC_LIB=$PWD/audio_tabs_b200/lib/variants/tuning.so
run() {
  timeout 300 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu --no-configs > $O/sweep.json 2> $O/sweep.err
  python - "$1" $O/sweep.json <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[2]))
    print(sys.argv[1], "step %.3f ms"%d["ms_per_step"], " ".join("%s:%.3f"%(k["frame_size"],k["ms"]) for k in d["roofline"]["per_kernel"]), flush=True)
except Exception as e:
    print(sys.argv[1],'FAILED',e)
PY
}
B200SPEC_SEAM=0 run "warm-up rows (default chunk)"
run "seam fix-up (default chunk)"
for c in 8 12 16 20 24 48 56 64 72 80; do B200SPEC_CHUNK=$c run "seam fix-up chunk=$c"; done
B200SPEC_SEAM=0 run "warm-up rows (default chunk)"
run "seam fix-up (default chunk)"
