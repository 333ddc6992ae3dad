"""Swap the CUDA processors into an importable madmom.

madmom's feature processors import the audio processors lazily inside ``__init__``
(``from ..audio.signal import SignalProcessor, FramedSignalProcessor`` ...), so replacing the module
attributes before ``RNNBeatProcessor()`` / ``DeepChromaProcessor()`` is constructed is sufficient:
the reference's callers (/root/reference/backend/app/services/grid/beats.py:71-75,
chords/extract.py:54, theory/key.py:101) then run this front end with the madmom networks untouched.
"""
from __future__ import annotations

_SWAPS = {
    "madmom.audio.signal": ["FramedSignalProcessor", "FramedSignal"],
    "madmom.audio.stft": ["ShortTimeFourierTransformProcessor", "ShortTimeFourierTransform"],
    "madmom.audio.spectrogram": [
        "SpectrogramProcessor", "FilteredSpectrogramProcessor", "LogarithmicSpectrogramProcessor",
        "LogarithmicFilteredSpectrogramProcessor", "SpectrogramDifferenceProcessor",
        "Spectrogram", "FilteredSpectrogram", "LogarithmicSpectrogram", "LogarithmicFilteredSpectrogram",
        "SpectrogramDifference"],
}
_saved = {}


def install():
    """Patch madmom in place; returns the list of replaced names. Raises ImportError without madmom."""
    import importlib
    from .audio import signal, spectrogram, stft
    ours = {"madmom.audio.signal": signal, "madmom.audio.stft": stft, "madmom.audio.spectrogram": spectrogram}
    replaced = []
    for modname, names in _SWAPS.items():
        mod = importlib.import_module(modname)
        for name in names:
            if (modname, name) not in _saved:
                _saved[(modname, name)] = getattr(mod, name)
            setattr(mod, name, getattr(ours[modname], name))
            replaced.append(modname + "." + name)
    return replaced


def uninstall():
    import importlib
    for (modname, name), obj in _saved.items():
        setattr(importlib.import_module(modname), name, obj)
    _saved.clear()
