#!/usr/bin/env python
"""Secondary measurements for the other BASELINE.json configs (bench.py measures configs[1]).

    python tools/bench_configs.py [--quick]

Device-resident inputs, CUDA-event timing (3 warm-ups, median of 5), one JSON line per config:
  config 1  1 x 30 s, frame 2048 / hop 441 / 12 bpo -> (3000, 81)            (latency bound)
  config 3a 256 x 180 s, frame 4096 hop 441, 24 bpo 65-2100 Hz -> chroma(12)
  config 3b same at hop 4410 (the app's chord rate): the HBM-bound variant
  N1        256 x 180 s, frame 8192 @ fps 10, 24 bpo 65-2100 Hz -> (1800, 105)  (DeepChroma front end)
  key       256 x 180 s int16, frame 8192 @ fps 5 -> (900, 105)                 (CNN key front end)
"""
import json
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from audio_tabs_b200.frontends import log_filt_spec  # noqa: E402
from audio_tabs_b200.plan import FrontEnd, Packed  # noqa: E402
from audio_tabs_b200.synth import synth_batch_device  # noqa: E402

SR = 44100
PEAK = 6547.5


def timeit(fn, reps=5):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts))


def run(name, spec, n_clips, seconds, dtype="f32", proj=False):
    dev = torch.device("cuda", 0)
    n = int(seconds * SR)
    sig = synth_batch_device(n_clips, n, seed=7, device=dev, dtype=dtype)
    fe = FrontEnd([spec], device=0, dtype=dtype)
    packed = Packed(sig, [n] * n_clips, spec.hop_size)
    if proj:
        out = torch.empty((packed.total_frames, 12), dtype=torch.float32, device=dev)
        fn = lambda: fe.run_packed(packed, out=False, proj=[out])  # noqa: E731
    else:
        out = fe.alloc_output(packed.total_frames)
        fn = lambda: fe.run_packed(packed, out)  # noqa: E731
    ms = timeit(fn)
    alg = sig.numel() * sig.element_size() + out.numel() * 4
    audio = n_clips * seconds
    print(json.dumps({"config": name, "clips": n_clips, "seconds": seconds, "frames": packed.total_frames,
                      "ms": ms, "audio_s_per_s": audio / ms * 1e3, "alg_gb": alg / 1e9,
                      "alg_gbs": alg / ms / 1e6, "hbm_frac_of_measured_peak": alg / ms / 1e6 / PEAK}), flush=True)
    del sig, out


def run_onset(nc, seconds):
    """librosa-style onset strength (N3): frame 2048, hop 512, 128 mel bands, dB, median envelope."""
    import numpy as np
    from audio_tabs_b200.onsets import OnsetStrength
    dev = torch.device("cuda", 0)
    n = int(seconds * SR)
    sig = synth_batch_device(nc, n, seed=9, device=dev)
    eng = OnsetStrength(sr=SR, aggregate=np.median)
    packed = Packed(sig, [n] * nc, 512.0, "extend")
    ms_mel = timeit(lambda: eng.mel_db(packed))
    ms_all = timeit(lambda: eng.envelope(packed))
    alg = sig.numel() * 4 + packed.total_frames * 4
    print(json.dumps({"config": "N3: %dx%ds onset_strength (2048/512, 128 mel, dB, median)" % (nc, seconds), "clips": nc,
                      "frames": packed.total_frames, "ms": ms_all, "ms_mel_db_only": ms_mel,
                      "audio_s_per_s": nc * seconds / ms_all * 1e3, "alg_gb": alg / 1e9, "alg_gbs": alg / ms_all / 1e6,
                      "hbm_frac_of_measured_peak": alg / ms_all / 1e6 / PEAK}), flush=True)


def main():
    quick = "--quick" in sys.argv
    nc = 32 if quick else 256
    run("1: 1x30s 2048/441/12bpo -> (3000,81)", log_filt_spec(2048, 441.0, 12), 1, 30)
    run("3a: %dx180s 4096 hop441 24bpo 65-2100 -> chroma12" % nc,
        log_filt_spec(4096, 441.0, 24, 65.0, 2100.0, fold=True), nc, 180, proj=True)
    run("3b: %dx180s 4096 hop4410 24bpo 65-2100 -> chroma12" % nc,
        log_filt_spec(4096, 4410.0, 24, 65.0, 2100.0, fold=True), nc, 180, proj=True)
    run("N1: %dx180s 8192 fps10 24bpo 65-2100 -> (1800,105)" % nc,
        log_filt_spec(8192, 4410.0, 24, 65.0, 2100.0), nc, 180)
    run("key: %dx180s int16 8192 fps5 -> (900,105)" % nc,
        log_filt_spec(8192, 8820.0, 24, 65.0, 2100.0, int16=True), nc, 180, dtype="i16")
    run_onset(nc, 180)


if __name__ == "__main__":
    main()
