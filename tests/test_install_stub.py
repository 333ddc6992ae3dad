"""install() against a STUB madmom package: the real madmom 0.16.1 is not installable here, so this only
proves the mechanism -- the module attributes the feature processors import lazily are replaced, the
swap is idempotent and reversible, and a processor built AFTER install() picks up the replacements the
way madmom.features.beats.RNNBeatProcessor.__init__ does (``from ..audio.signal import ...`` inside
``__init__``).  The numerical side is covered by the parity tests."""
import sys
import types

import pytest

import audio_tabs_b200 as b2
from audio_tabs_b200 import install as inst


def _stub_madmom():
    """Minimal module tree with the names madmom 0.16.1 defines in audio/{signal,stft,spectrogram}.py."""
    mods = {}
    for name in ("madmom", "madmom.audio", "madmom.audio.signal", "madmom.audio.stft", "madmom.audio.spectrogram",
                 "madmom.features", "madmom.features.beats"):
        mods[name] = types.ModuleType(name)
    for modname, names in inst._SWAPS.items():
        for n in names:
            setattr(mods[modname], n, type("Stock" + n, (), {"stock": True}))
    mods["madmom.audio.signal"].SignalProcessor = type("StockSignalProcessor", (), {"stock": True})

    class RNNBeatProcessor:                       # imports lazily inside __init__, like madmom's
        def __init__(self):
            from madmom.audio.signal import FramedSignalProcessor, SignalProcessor
            from madmom.audio.spectrogram import (FilteredSpectrogramProcessor, LogarithmicSpectrogramProcessor,
                                                  SpectrogramDifferenceProcessor)
            from madmom.audio.stft import ShortTimeFourierTransformProcessor
            self.classes = [SignalProcessor, FramedSignalProcessor, ShortTimeFourierTransformProcessor,
                            FilteredSpectrogramProcessor, LogarithmicSpectrogramProcessor,
                            SpectrogramDifferenceProcessor]
    mods["madmom.features.beats"].RNNBeatProcessor = RNNBeatProcessor
    mods["madmom"].audio = mods["madmom.audio"]
    mods["madmom.audio"].signal = mods["madmom.audio.signal"]
    mods["madmom.audio"].stft = mods["madmom.audio.stft"]
    mods["madmom.audio"].spectrogram = mods["madmom.audio.spectrogram"]
    return mods


@pytest.fixture
def stub(monkeypatch):
    mods = _stub_madmom()
    for k, v in mods.items():
        monkeypatch.setitem(sys.modules, k, v)
    inst._saved.clear()
    yield mods
    inst._saved.clear()


def test_install_swaps_and_restores(stub):
    replaced = inst.install()
    assert len(replaced) == sum(len(v) for v in inst._SWAPS.values())
    sig, stft, spec = stub["madmom.audio.signal"], stub["madmom.audio.stft"], stub["madmom.audio.spectrogram"]
    assert sig.FramedSignalProcessor is b2.FramedSignalProcessor
    assert stft.ShortTimeFourierTransformProcessor is b2.ShortTimeFourierTransformProcessor
    assert spec.FilteredSpectrogramProcessor is b2.FilteredSpectrogramProcessor
    assert spec.SpectrogramDifferenceProcessor is b2.SpectrogramDifferenceProcessor
    assert getattr(sig.SignalProcessor, "stock", False)          # SignalProcessor is left alone (host object)
    proc = stub["madmom.features.beats"].RNNBeatProcessor()       # built after install(): sees our classes
    assert proc.classes[1] is b2.FramedSignalProcessor and proc.classes[5] is b2.SpectrogramDifferenceProcessor
    inst.install()                                                # idempotent: the originals stay saved
    inst.uninstall()
    assert getattr(sig.FramedSignalProcessor, "stock", False) and getattr(spec.Spectrogram, "stock", False)
    assert not inst._saved


def test_install_without_madmom_raises(monkeypatch):
    for k in [k for k in sys.modules if k == "madmom" or k.startswith("madmom.")]:
        monkeypatch.delitem(sys.modules, k)
    monkeypatch.setattr(sys, "path", [p for p in sys.path])       # madmom is not importable in this image
    with pytest.raises(ImportError):
        inst.install()
