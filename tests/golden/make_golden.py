#!/usr/bin/env python
"""Generate the committed golden fixtures under tests/golden/.

madmom 0.16.1 (the real reference implementation of this path) cannot be imported in the build
container (not installed, no network), so the fixtures are produced by the repo's oracle
(oracle/madmom_ref.py, a numpy restatement pinned by the structural constants in
tests/test_oracle_pins.py) on
  * seeded synthetic guitar clips (audio_tabs_b200.synth.synth_guitar), and
  * the first 3 s of the reference's own recorded job audio
    /root/reference/data/jobs/c34b660dfb454be486983b1913bab38c/work/audio_mono_44k.wav
    (int16 mono 44.1 kHz), when that file is present.
Run from the repo root:  python tests/golden/make_golden.py
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from audio_tabs_b200.synth import synth_guitar  # noqa: E402
from oracle import madmom_ref as ref  # noqa: E402

OUT = Path(__file__).resolve().parent
WAV = Path("/root/reference/data/jobs/c34b660dfb454be486983b1913bab38c/work/audio_mono_44k.wav")


def main():
    x = synth_guitar(4242, 2.0)
    np.save(OUT / "guitar_2s_f32.npy", x)
    np.save(OUT / "guitar_2s_beat314.npy", ref.rnn_beat_preprocessor()(x))
    np.save(OUT / "guitar_2s_onset266.npy", ref.rnn_onset_preprocessor()(x))
    np.save(OUT / "guitar_2s_logfilt81.npy", ref.log_filtered_spectrogram(x))
    stft = ref.ShortTimeFourierTransform(ref.FramedSignal(ref.Signal(x[:22050], sample_rate=44100), frame_size=2048)).data
    np.save(OUT / "guitar_0p5s_stft2048.npy", stft)
    if WAV.exists():
        from scipy.io import wavfile
        sr, data = wavfile.read(WAV)
        assert sr == 44100 and data.dtype == np.int16 and data.ndim == 1
        clip = np.ascontiguousarray(data[44100 * 2:44100 * 5])
        np.save(OUT / "refjob_3s_i16.npy", clip)
        np.save(OUT / "refjob_3s_deepchroma105.npy", ref.log_filt_chain(8192, fps=10)(clip).data)
        np.save(OUT / "refjob_3s_key105.npy", ref.log_filt_chain(8192, fps=5)(clip).data)
        np.save(OUT / "refjob_3s_beat314.npy", ref.rnn_beat_preprocessor()(clip))
        np.save(OUT / "refjob_total_frames.npy", np.array([len(data), ref.num_frames_for(len(data), 441.0)]))
    for p in sorted(OUT.glob("*.npy")):
        print(p.name, np.load(p).shape, p.stat().st_size)


if __name__ == "__main__":
    main()
