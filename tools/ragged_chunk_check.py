#!/usr/bin/env python
"""Task-size model on RAGGED batches (tuning build: B200SPEC_CHUNK forces the size): device-resident time of the beat
front end on clips of very different lengths, model-chosen task size against forced ones.
    B200SPEC_LIB=audio_tabs_b200/lib/variants/tuning.so python tools/ragged_chunk_check.py"""
import json, os, subprocess, sys
from pathlib import Path
import numpy as np

if len(sys.argv) > 1 and sys.argv[1] == "child":
    import torch
    sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
    from audio_tabs_b200.frontends import beat_specs
    from audio_tabs_b200.plan import FrontEnd, Packed
    from audio_tabs_b200.synth import synth_batch_device
    rng = np.random.default_rng(int(sys.argv[2]))
    kind = sys.argv[3]
    if kind == "ragged":
        secs = rng.uniform(3.0, 300.0, size=160)
    elif kind == "few_long":
        secs = np.full(5, 1200.0)
    else:
        secs = rng.uniform(0.2, 8.0, size=3000)
    lens = [int(s * 44100) for s in secs]
    dev = torch.device("cuda", 0)
    sig = synth_batch_device(1, sum(lens), seed=5, device=dev)
    fe = FrontEnd(beat_specs(), device=0)
    packed = Packed(sig, lens, fe.hop_size)
    out = fe.alloc_output(packed.total_frames)
    for _ in range(3):
        fe.run_packed(packed, out)
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fe.run_packed(packed, out); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    print(json.dumps({"kind": kind, "clips": len(lens), "audio_s": round(float(sum(secs))), "chunk": os.environ.get("B200SPEC_CHUNK", "model"),
                      "ms": round(float(np.median(ts)), 3)}))
    sys.exit(0)

for kind in ("ragged", "few_long", "many_short"):
    for chunk in ("", "16", "32", "48", "64", "96"):
        env = dict(os.environ)
        if chunk:
            env["B200SPEC_CHUNK"] = chunk
        else:
            env.pop("B200SPEC_CHUNK", None)
        r = subprocess.run([sys.executable, __file__, "child", "1", kind], env=env, capture_output=True, text=True, timeout=600)
        print(r.stdout.strip() or r.stderr[-300:], flush=True)
