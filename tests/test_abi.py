"""The C-ABI library loads and exports every symbol include/b200spec.h declares; the integer frame
geometry it computes is bit-exact with madmom's float64 arithmetic.  No compute calls (no GPU)."""
import ctypes as C
import re
from pathlib import Path

import numpy as np
import pytest

from oracle import madmom_ref as ref

ROOT = Path(__file__).resolve().parent.parent


def test_header_symbols_exported(lib_built):
    from audio_tabs_b200 import _ffi
    header = (ROOT / "include" / "b200spec.h").read_text()
    declared = set(re.findall(r"\b(b200spec_[a-z_0-9]+)\s*\(", header))
    declared -= {"b200spec_plan_desc", "b200spec_res_desc", "b200spec_out_desc"}
    assert len(declared) >= 16
    handle = C.CDLL(str(lib_built))
    for name in declared:
        assert hasattr(handle, name), name
    bound = {name for name, _, _ in _ffi.SYMBOLS}
    assert declared == bound, declared ^ bound
    assert _ffi.lib().b200spec_abi_version() == _ffi.ABI_VERSION


def test_num_frames_bit_exact(lib_built):
    from audio_tabs_b200 import _ffi
    rng = np.random.default_rng(0)
    hops = [441.0, 4410.0, 8820.0, 44100 / 100.0, 44100 / 30.0, 22050 / 7.0, 512.0, 1.0, 48000 / 23.976]
    for hop in hops:
        for n in list(rng.integers(0, 30_000_000, size=40)) + [0, 1, 440, 441, 442, 123481, 675192]:
            for end in ("normal", "extend"):
                assert _ffi.num_frames(int(n), hop, end) == ref.num_frames_for(int(n), hop, end)
    with pytest.raises(ValueError):
        _ffi.num_frames(100, 441.0, "bogus")
    with pytest.raises(ValueError):
        _ffi.num_frames(100, 0.0)


def test_frame_start_bit_exact(lib_built):
    from audio_tabs_b200 import _ffi
    rng = np.random.default_rng(1)
    for hop in (441.0, 44100 / 30.0, 22050 / 7.0, 4410.0):
        for idx in list(rng.integers(0, 200000, size=50)) + [0, 1, 2]:
            for fs, origin in ((1024, 0), (2048, 0), (4096, 0), (8192, 0), (2048, 1023), (2048, -1024)):
                assert _ffi.frame_start(int(idx), hop, fs, origin) == ref.frame_start(int(idx), fs, hop, origin)


def test_plan_create_fails_loudly_without_gpu(lib_built):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from audio_tabs_b200 import _ffi
    from audio_tabs_b200.frontends import beat_specs
    from audio_tabs_b200.plan import DevicePlan
    with pytest.raises((_ffi.B200SpecError, RuntimeError)):
        DevicePlan(0, "f32", 1, beat_specs())


def test_null_and_bad_arguments(lib_built):
    from audio_tabs_b200 import _ffi
    lib = _ffi.lib()
    assert lib.b200spec_plan_create(None, None) == _ffi.ERR_ARG
    assert b"NULL" in lib.b200spec_last_error()
    assert lib.b200spec_num_frames(10, 441.0, 0, None) == _ffi.ERR_ARG
    assert lib.b200spec_plan_num_res(None) == 0
    assert lib.b200spec_workspace_bytes(64) >= 4 * 65 + 16
    assert lib.b200spec_launch_count() >= 0


def test_missing_library_is_a_hard_error(monkeypatch, tmp_path):
    from audio_tabs_b200 import _ffi
    monkeypatch.setenv("B200SPEC_LIB", str(tmp_path / "nope.so"))
    monkeypatch.setattr(_ffi, "_LIB", None)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _ffi.lib()
