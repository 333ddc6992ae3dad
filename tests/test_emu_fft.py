"""The thread-mapped FFT passes (audio_tabs_b200/csrc/fft_core.cuh) compiled for the host and run
thread by thread: index maps, in-place pass 2, the real-spectrum split and both forms of the
self-paired columns agree with a float64 DFT.  Test infrastructure only."""
import shutil
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


@pytest.mark.skipif(shutil.which("g++") is None, reason="g++ not available")
def test_fft_passes_on_host(tmp_path):
    exe = tmp_path / "emu_fft"
    subprocess.run(["g++", "-O2", "-std=c++17", "-o", str(exe), str(ROOT / "tests" / "emu" / "emu_fft.cpp")], check=True)
    res = subprocess.run([str(exe)], capture_output=True, text=True)
    assert res.returncode == 0, res.stdout + res.stderr
    assert res.stdout.strip().endswith("OK")


@pytest.mark.skipif(shutil.which("g++") is None, reason="g++ not available")
def test_pair_transform_on_host(tmp_path):
    """Two real frames as one complex FFT of F points (k_front_pair): passes 1-3 thread by thread."""
    exe = tmp_path / "emu_pair"
    subprocess.run(["g++", "-O2", "-std=c++17", "-o", str(exe), str(ROOT / "tests" / "emu" / "emu_pair.cpp")], check=True)
    res = subprocess.run([str(exe)], capture_output=True, text=True)
    assert res.returncode == 0, res.stdout + res.stderr
    assert res.stdout.strip().endswith("OK")
