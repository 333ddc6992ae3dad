"""install() against a behavioural stand-in for madmom 0.16.1 (tests/madmom_stub.py).

The real madmom is not installable here, so the drop-in claim is tested against a package
that reproduces what decides whether swapped-in processors survive inside madmom's feature processors:
ndarray-subclass stages, ``Signal.__new__`` loading every non-ndarray as a file, ``_process`` forwarding
kwargs only to madmom's own Processor instances, lazy imports inside ``__init__``, ``_dcp_flatten`` and
``np.hstack``.  The CPU tests cover the mechanism; the ``gpu`` tests run the three feature-processor shapes
the reference calls (/root/reference/backend/app/services/grid/beats.py:71-75, chords/extract.py:54-57,
theory/key.py:99-101,143-144) before and after ``install()`` and compare the network outputs.
"""
import sys

import numpy as np
import pytest

import audio_tabs_b200 as b2
from audio_tabs_b200 import install as inst
from audio_tabs_b200.synth import synth_guitar
from oracle import madmom_ref as ref
import madmom_stub


@pytest.fixture
def stub(monkeypatch):
    mods = madmom_stub.activate(monkeypatch)
    inst._saved.clear()
    yield mods
    inst.uninstall()
    inst._saved.clear()


def test_stub_stock_path_is_the_oracle(stub):
    """the unpatched stub computes what the oracle computes (it is the expected value of the gpu tests)"""
    from madmom.features.beats import RNNBeatProcessor
    from madmom.audio.signal import Signal
    x = synth_guitar(77, 0.6)
    proc = RNNBeatProcessor()
    data = proc.processors[0](Signal(x, sample_rate=44100, num_channels=1))     # (pre_processor, nn)
    want = ref.rnn_beat_preprocessor()(x)
    assert isinstance(data, np.ndarray) and data.shape == want.shape == (60, 314)
    np.testing.assert_array_equal(np.asarray(data), want)


def test_stub_signal_loads_non_arrays_as_files(stub):
    """the hazard of VERDICT r1 weak #7: stock Signal() on a non-ndarray stage tries to open a file"""
    from madmom.audio.signal import Signal
    from madmom.io.audio import LoadAudioFileError

    class NotAnArray:
        shape = (4, 3)

        def __array__(self, dtype=None, copy=None):
            return np.zeros((4, 3), np.float32)

    with pytest.raises(LoadAudioFileError):
        Signal(NotAnArray(), sample_rate=10)


def test_install_swaps_and_restores(stub):
    sig, stft, spec = stub["madmom.audio.signal"], stub["madmom.audio.stft"], stub["madmom.audio.spectrogram"]
    stock = {(m, n): getattr(stub[m], n) for m, names in inst._SWAPS.items() for n in names}
    replaced = inst.install()
    assert len(replaced) == sum(len(v) for v in inst._SWAPS.values())
    madmom_processor = stub["madmom.processors"].Processor
    # processors are re-based onto madmom's Processor (kwargs forwarding); data classes are ours as they are
    for mod, name, ours in ((sig, "FramedSignalProcessor", b2.FramedSignalProcessor),
                            (sig, "SignalProcessor", b2.SignalProcessor),
                            (stft, "ShortTimeFourierTransformProcessor", b2.ShortTimeFourierTransformProcessor),
                            (spec, "FilteredSpectrogramProcessor", b2.FilteredSpectrogramProcessor),
                            (spec, "SpectrogramDifferenceProcessor", b2.SpectrogramDifferenceProcessor)):
        cls = getattr(mod, name)
        assert issubclass(cls, ours) and issubclass(cls, madmom_processor) and cls.__name__ == name
    assert sig.Signal is b2.Signal and sig.FramedSignal is b2.FramedSignal
    assert spec.LogarithmicFilteredSpectrogram is b2.LogarithmicFilteredSpectrogram
    first = {(m, n): getattr(stub[m], n) for m, names in inst._SWAPS.items() for n in names}
    inst.install()                                                # idempotent: same classes, the originals stay saved
    assert first == {(m, n): getattr(stub[m], n) for m, names in inst._SWAPS.items() for n in names}
    inst.uninstall()
    assert stock == {(m, n): getattr(stub[m], n) for m, names in inst._SWAPS.items() for n in names}
    assert not inst._saved


def test_kwargs_reach_swapped_processors(stub):
    """madmom's _process forwards **kwargs only to ITS Processor instances: the swapped classes must be such"""
    inst.install()
    from madmom.audio.signal import FramedSignalProcessor, SignalProcessor
    from madmom.processors import SequentialProcessor
    chain = SequentialProcessor([SignalProcessor(num_channels=1, sample_rate=44100),
                                 FramedSignalProcessor(frame_size=2048, fps=100)])
    x = synth_guitar(3, 0.25)
    assert len(chain(x)) == 25
    assert len(chain(x, fps=50)) == 13                            # the kwarg arrived at our FramedSignalProcessor
    assert chain(x, frame_size=1024).frame_size == 1024


def test_install_without_madmom_raises(monkeypatch):
    for k in [k for k in sys.modules if k == "madmom" or k.startswith("madmom.")]:
        monkeypatch.delitem(sys.modules, k)
    monkeypatch.setattr(sys, "path", [p for p in sys.path])       # madmom is not importable in this image
    inst._saved.clear()
    with pytest.raises(ImportError):
        inst.install()


# ---- through the GPU: the three feature-processor shapes the reference calls ------------------------------
def _launches():
    from audio_tabs_b200 import _ffi
    return _ffi.launch_count()


@pytest.mark.gpu
def test_rnn_beat_processor_after_install(stub, cuda_device):
    """grid/beats.py:28-32,71-75: Signal(arr, sample_rate=sr, num_channels=1) -> RNNBeatProcessor()(signal)"""
    x = synth_guitar(4242, 2.0)
    from madmom.features.beats import RNNBeatProcessor
    from madmom.audio.signal import Signal
    want = RNNBeatProcessor()(Signal(x, sample_rate=44100, num_channels=1))
    inst.install()
    from madmom.audio.signal import Signal as Signal2             # what beats.py imports at call time
    assert Signal2 is b2.Signal
    n0 = _launches()
    got = RNNBeatProcessor()(Signal2(x, sample_rate=44100, num_channels=1))
    assert _launches() - n0 >= 3                                   # three fused front-end launches (+ task tables)
    assert isinstance(got, np.ndarray) and got.shape == want.shape == (200,)
    np.testing.assert_allclose(got, want, rtol=1e-4, atol=1e-5)


@pytest.mark.gpu
def test_deep_chroma_processor_after_install(stub, cuda_device):
    """chords/extract.py:54-57: the chain that re-wraps the spectrogram with SignalProcessor(sample_rate=10),
    frames it (15, hop 1) and flattens it with _dcp_flatten -- the LazyArray hazard of VERDICT r1 weak #7"""
    x = synth_guitar(4243, 3.0)
    from madmom.audio.chroma import DeepChromaProcessor
    from madmom.audio.signal import Signal
    want = DeepChromaProcessor()(Signal(x, sample_rate=44100, num_channels=1))
    inst.install()
    from madmom.audio.signal import Signal as Signal2
    n0 = _launches()
    proc = DeepChromaProcessor()
    got = proc(Signal2(x, sample_rate=44100, num_channels=1))
    assert _launches() - n0 >= 1
    assert got.shape == want.shape == (30, 12)
    np.testing.assert_allclose(got, want, rtol=1e-4, atol=1e-5)
    # the (T, 1575) network input itself, stage by stage
    data = Signal2(x, sample_rate=44100, num_channels=1)
    for p in proc.processors[:-1]:
        data = p(data)
    assert data.shape == (30, 1575)
    np.testing.assert_allclose(data, ref.dcp_context(ref.log_filt_chain(8192, fps=10)(x).data), rtol=1e-4, atol=1e-5)


@pytest.mark.gpu
def test_cnn_key_processor_after_install(stub, cuda_device, tmp_path):
    """theory/key.py:143-144: a WAV PATH goes in (int16 PCM, window / 32767); the network layers touch
    data.ndim / data.shape / arithmetic / reshape on our lazy spectrogram"""
    from scipy.io import wavfile
    x = (synth_guitar(4244, 3.0) * 20000).astype(np.int16)
    path = str(tmp_path / "harmonic.wav")
    wavfile.write(path, 44100, x)
    from madmom.features.key import CNNKeyRecognitionProcessor
    want = CNNKeyRecognitionProcessor()(path)
    inst.install()
    n0 = _launches()
    got = CNNKeyRecognitionProcessor()(path)
    assert _launches() - n0 >= 1
    assert got.shape == want.shape == (1, 24)
    np.testing.assert_allclose(got, want, rtol=1e-4, atol=1e-5)


def test_madmom_golden_script_against_the_stub(stub, monkeypatch, capsys):
    """tests/golden/make_madmom_golden.py (the script that pins the oracle wherever the real madmom exists) runs end to
    end: against the stub -- whose numerics are the oracle's -- every comparison must pass, including the
    committed fixtures under tests/golden/."""
    import importlib.util
    from pathlib import Path
    path = Path(__file__).resolve().parent / "golden" / "make_madmom_golden.py"
    spec = importlib.util.spec_from_file_location("make_madmom_golden", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    monkeypatch.setattr(sys, "argv", [str(path)])
    rc = mod.main()
    out = capsys.readouterr().out
    assert rc == 0, out
    assert "FAIL" not in out and "guitar_2s_beat314.npy" in out and "refjob_3s_key105.npy" in out
