#!/usr/bin/env python
"""e2e (pinned host in -> pinned host out) throughput of config 2 for different pipeline group sizes."""
import json, sys, time
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from audio_tabs_b200.frontends import beat_specs
from audio_tabs_b200.plan import FrontEnd
from audio_tabs_b200.synth import synth_batch_device

SR, NC, SEC = 44100, 64, 180
dev = torch.device("cuda", 0)
n = SEC * SR
DT = "i16" if "--i16" in sys.argv else "f32"
fe = FrontEnd(beat_specs(int16=(DT == "i16")), device=0, dtype=DT)
sig = synth_batch_device(NC, n, seed=2000, device=dev, dtype=DT)
host_in = torch.empty(sig.shape, dtype=sig.dtype, pin_memory=True); host_in.copy_(sig); del sig
T = 18000 * NC
host_out = torch.empty((T, fe.width), dtype=torch.float32, pin_memory=True)
lens = [n] * NC
CASES = [(2, 2), ([1, 2], 2), (1, 2), (2, 3), ([1, 2], 3), (4, 2), ([1, 4], 3), (8, 2), ([1, 2, 4], 3), ([1, 1, 2, 4, 8], 3)]
for rep in range(2):
    for gc, ns in CASES:
        for _ in range(2):
            fe.process_batch_pinned(host_in, lens, host_out, group_clips=gc, n_slots=ns)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(5):
            fe.process_batch_pinned(host_in, lens, host_out, group_clips=gc, n_slots=ns)
        b.record(); torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 5
        print(json.dumps({"dtype": DT, "group_clips": gc, "slots": ns, "ms_per_step": round(ms, 3), "audio_s_per_s": round(NC * SEC / ms * 1e3)}), flush=True)
