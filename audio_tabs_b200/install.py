"""Swap the CUDA processors into an importable madmom.

madmom's feature processors import the audio processors lazily inside ``__init__``
(``from ..audio.signal import SignalProcessor, FramedSignalProcessor`` ...), so replacing the module
attributes before ``RNNBeatProcessor()`` / ``DeepChromaProcessor()`` / ``CNNKeyRecognitionProcessor()``
is constructed is sufficient: the reference's callers
(/root/reference/backend/app/services/grid/beats.py:71-75, chords/extract.py:54-57,
theory/key.py:99-101,143-144) then run this front end with the madmom networks untouched.

Two madmom behaviours shape what is swapped:

* ``madmom.audio.signal.Signal.__new__`` treats every object that is not an ``np.ndarray`` as a file name
  and tries to load it.  Our stages are lazy (``audio/lazy.py``: computed on the GPU when their values are
  first needed), not ndarrays, and ``DeepChromaProcessor`` re-wraps the spectrogram with
  ``SignalProcessor(sample_rate=10)`` before the context stacking -- so ``Signal`` and ``SignalProcessor``
  are swapped as well (ours materialise a lazy stage through ``np.asarray``).  The same swap serves
  ``beats.py:28-32`` / ``extract.py:37-41``, which build ``madmom.audio.signal.Signal(arr, sample_rate=sr,
  num_channels=1)`` themselves, and ``key.py:144``, which passes a WAV path.
* ``madmom.processors._process`` forwards ``**kwargs`` only to instances of madmom's own ``Processor``
  class; the swapped-in processor classes are therefore re-based onto it (a subclass of ours AND of
  ``madmom.processors.Processor``), so keyword arguments given to a feature processor still arrive.
"""
from __future__ import annotations

_SWAPS = {
    "madmom.audio.signal": ["Signal", "SignalProcessor", "FramedSignalProcessor", "FramedSignal"],
    "madmom.audio.stft": ["ShortTimeFourierTransformProcessor", "ShortTimeFourierTransform"],
    "madmom.audio.spectrogram": [
        "SpectrogramProcessor", "FilteredSpectrogramProcessor", "LogarithmicSpectrogramProcessor",
        "LogarithmicFilteredSpectrogramProcessor", "SpectrogramDifferenceProcessor",
        "Spectrogram", "FilteredSpectrogram", "LogarithmicSpectrogram", "LogarithmicFilteredSpectrogram",
        "SpectrogramDifference"],
}
_saved = {}
_adopted = {}     # (our class, madmom's Processor class) -> subclass of both


def _adopt(cls, madmom_processor):
    """``cls`` re-based onto madmom's Processor so that ``isinstance(obj, madmom.processors.Processor)`` holds."""
    from .processors import Processor
    if madmom_processor is None or not isinstance(cls, type) or not issubclass(cls, Processor) \
            or issubclass(cls, madmom_processor):
        return cls
    key = (cls, madmom_processor)
    if key not in _adopted:
        _adopted[key] = type(cls.__name__, (cls, madmom_processor),
                             {"__module__": cls.__module__, "__doc__": cls.__doc__, "_b200spec_base": cls})
    return _adopted[key]


def install():
    """Patch madmom in place; returns the list of replaced names. Raises ImportError without madmom."""
    import importlib
    from .audio import signal, spectrogram, stft
    ours = {"madmom.audio.signal": signal, "madmom.audio.stft": stft, "madmom.audio.spectrogram": spectrogram}
    try:
        madmom_processor = getattr(importlib.import_module("madmom.processors"), "Processor", None)
    except ImportError:
        madmom_processor = None
    if not isinstance(madmom_processor, type):
        madmom_processor = None
    replaced = []
    for modname, names in _SWAPS.items():
        mod = importlib.import_module(modname)
        for name in names:
            if (modname, name) not in _saved:
                _saved[(modname, name)] = getattr(mod, name)
            setattr(mod, name, _adopt(getattr(ours[modname], name), madmom_processor))
            replaced.append(modname + "." + name)
    return replaced


def uninstall():
    import importlib
    for (modname, name), obj in _saved.items():
        setattr(importlib.import_module(modname), name, obj)
    _saved.clear()
