#!/usr/bin/env python
"""Smallest run that touches every kernel family (for compute-sanitizer): the beat front end (pair kernels
1024/2048/4096 + diff), a ragged batch, the 8192 chain (k_front, pruned magnitudes), stereo int16 input,
the one-launch kernel, chroma projection, SuperFlux, onset strength, context stacking.  Results are checked against the oracle."""
import sys
from pathlib import Path
import numpy as np
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from audio_tabs_b200.audio.chroma import chord_chroma_frontend, context_stack_device, deep_chroma_frontend
from audio_tabs_b200.frontends import beat_specs, log_filt_spec
from audio_tabs_b200.onsets import OnsetStrength
from audio_tabs_b200.plan import FrontEnd
from audio_tabs_b200.synth import synth_guitar
from oracle import madmom_ref as ref

def close(a, b):
    return bool((np.abs(a.astype(np.float64) - b) <= 1e-5 + 1e-4 * np.abs(b)).all())

x = synth_guitar(1, 0.6)
fe = FrontEnd(beat_specs(), device=0)
outs = fe.process_batch([x, x[:5000], x[:300]])
ok = close(outs[0], ref.rnn_beat_preprocessor()(x)) and close(outs[2], ref.rnn_beat_preprocessor()(x[:300]))
one = FrontEnd(beat_specs(), device=0, one_launch=True)        # k_front_multi: all resolutions in one launch, with status
outs1, status = one.process_batch([x, x[:5000], x[:300]], return_status=True)
ok &= close(outs1[0], ref.rnn_beat_preprocessor()(x)) and close(outs1[1], ref.rnn_beat_preprocessor()(x[:5000])) and not status.any()
st = np.stack([x, x * 0.5], axis=1)
i16 = np.clip(np.round(st * 20000), -32768, 32767).astype(np.int16)
o16 = FrontEnd(beat_specs(int16=True), device=0, dtype="i16", channels=2).process_batch([i16])[0]
ok &= close(o16, ref.rnn_beat_preprocessor()(ref.Signal(i16, sample_rate=44100, num_channels=1)))
d = np.asarray(deep_chroma_frontend()(x))
ok &= close(d, ref.log_filt_chain(8192, fps=10)(x).data)
c = np.asarray(chord_chroma_frontend(4096, fps=10)(x))
ok &= c.shape[1] == 12
sf = FrontEnd([log_filt_spec(2048, 441.0, 12, diff_ratio=0.5, diff_max_bins=3)], device=0)
ok &= sf.process_batch([x])[0].shape[1] == 162
env = OnsetStrength(sr=44100, aggregate=np.median).process_batch([x])[0]
ok &= env.shape[0] == 1 + len(x) // 512
cs = context_stack_device(torch.from_numpy(d).cuda(), 15)
ok &= tuple(cs.shape) == (d.shape[0], 15 * 105)
torch.cuda.synchronize()
print("sanitize case:", "OK" if ok else "MISMATCH")
sys.exit(0 if ok else 1)
