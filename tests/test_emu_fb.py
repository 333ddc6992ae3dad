"""The fused filterbank stage (audio_tabs_b200/csrc/fb_pack.h + fb_core.cuh) compiled for the host:
every filterbank the reference's callers use, plus awkward ones (duplicate bands, a pitch-class
bank whose bands span the whole spectrum, a dense full-band bank), is packed into slabs, run thread
by thread and compared with the dense product.  Test infrastructure only."""
import shutil
import subprocess
from pathlib import Path

import numpy as np
import pytest

from audio_tabs_b200.audio.stft import fft_frequencies
from audio_tabs_b200.filters import LogarithmicFilterbank, PitchClassProfileFilterbank, SlaneyMelFilterbank

ROOT = Path(__file__).resolve().parent.parent


def _banks():
    sr = 44100
    out = []
    for F, nb, fmin, fmax, uniq in [(1024, 3, 30, 17000, True), (2048, 6, 30, 17000, True), (4096, 12, 30, 17000, True),
                                    (2048, 12, 30, 17000, True), (2048, 12, 30, 17000, False), (1024, 6, 30, 17000, True),
                                    (4096, 6, 30, 17000, True), (4096, 24, 65, 2100, True), (8192, 24, 65, 2100, True),
                                    (8192, 24, 60, 2600, True), (8192, 24, 30, 20000, True), (4096, 24, 30, 17000, True)]:
        fb = LogarithmicFilterbank(fft_frequencies(F >> 1, sr), num_bands=nb, fmin=fmin, fmax=fmax, unique_filters=uniq)
        out.append((F, fb.banded()))
    out.append((4096, PitchClassProfileFilterbank(fft_frequencies(2048, sr)).banded()))
    out.append((2048, SlaneyMelFilterbank(sr, 2048, n_mels=128).banded()))      # librosa onset strength
    out.append((4096, SlaneyMelFilterbank(22050, 4096, n_mels=64).banded()))
    # one rectangular band over everything, and an empty band between two real ones
    out.append((2048, (np.array([0], np.int32), np.array([1024], np.int32), np.array([0], np.int32),
                       np.full(1024, 1.0 / 1024, np.float32))))
    out.append((1024, (np.array([3, 0, 40], np.int32), np.array([5, 0, 9], np.int32), np.array([0, 5, 5], np.int32),
                       np.linspace(0.1, 1.0, 14).astype(np.float32))))
    return out


@pytest.mark.skipif(shutil.which("g++") is None, reason="g++ not available")
def test_filterbank_slabs_on_host(tmp_path):
    exe = tmp_path / "emu_fb"
    subprocess.run(["g++", "-O2", "-std=c++17", "-o", str(exe), str(ROOT / "tests" / "emu" / "emu_fb.cpp")], check=True)
    spec = tmp_path / "banks.txt"
    with open(spec, "w") as fh:
        for F, (start, length, _woff, w) in _banks():
            fh.write("%d %d\n" % (F, len(start)))
            fh.write(" ".join(str(int(v)) for v in start) + "\n")
            fh.write(" ".join(str(int(v)) for v in length) + "\n")
            fh.write(" ".join(repr(float(v)) for v in w) + "\n")
    res = subprocess.run([str(exe), str(spec)], capture_output=True, text=True)
    assert res.returncode == 0, res.stdout + res.stderr
    assert res.stdout.strip().endswith("OK")
    print(res.stdout)
