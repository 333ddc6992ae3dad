#!/usr/bin/env python
"""One named workload, device resident, at the full bench size -- the command `ncu` wraps (tools/gpu_ncu2.sh).

    python tools/ncu_case.py CASE [--reps 2]
      step3   config 2 (64 x 180 s), one kernel launch per resolution   (k_front_pair<1024|2048|4096>)
      multi   config 2, all three resolutions in one launch              (k_front_multi)
      3b      config 3 at hop 4410 (256 x 180 s, frame 4096 -> chroma)   (k_front_pair<4096>)
      N1      DeepChroma front end (256 x 180 s, frame 8192 @ fps 10)    (k_front<8192>)
      key     CNN key front end (256 x 180 s int16, frame 8192 @ fps 5)  (k_front<8192>)
Prints the CUDA-event time of the last repetition (a number taken under ncu is never a bench value).
"""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from audio_tabs_b200.frontends import beat_specs, log_filt_spec  # noqa: E402
from audio_tabs_b200.plan import FrontEnd, Packed  # noqa: E402
from audio_tabs_b200.synth import synth_batch_device  # noqa: E402

SR = 44100
case = sys.argv[1]
reps = int(sys.argv[sys.argv.index("--reps") + 1]) if "--reps" in sys.argv else 2
dev = torch.device("cuda", 0)
if case in ("step3", "multi"):
    nc, n, dtype = 64, 180 * SR, "f32"
    fe = FrontEnd(beat_specs(), device=0, one_launch=(case == "multi"))
    proj = None
elif case == "3b":
    nc, n, dtype = 256, 180 * SR, "f32"
    fe = FrontEnd([log_filt_spec(4096, 4410.0, 24, 65.0, 2100.0, fold=True)], device=0)
    proj = True
elif case == "N1":
    nc, n, dtype = 256, 180 * SR, "f32"
    fe = FrontEnd([log_filt_spec(8192, 4410.0, 24, 65.0, 2100.0)], device=0)
    proj = None
elif case == "key":
    nc, n, dtype = 256, 180 * SR, "i16"
    fe = FrontEnd([log_filt_spec(8192, 8820.0, 24, 65.0, 2100.0, int16=True)], device=0, dtype="i16")
    proj = None
else:
    raise SystemExit("unknown case " + case)
sig = synth_batch_device(nc, n, seed=7, device=dev, dtype=dtype)
packed = Packed(sig, [n] * nc, fe.hop_size)
if proj:
    pout = torch.empty((packed.total_frames, 12), dtype=torch.float32, device=dev)
    run = lambda: fe.run_packed(packed, out=False, proj=[pout])  # noqa: E731
else:
    out = fe.alloc_output(packed.total_frames)
    run = lambda: fe.run_packed(packed, out)  # noqa: E731
for _ in range(reps):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); run(); b.record()
    torch.cuda.synchronize()
print("%s: %.3f ms (last of %d repetitions)" % (case, a.elapsed_time(b), reps))
