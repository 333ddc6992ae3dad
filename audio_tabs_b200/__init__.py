"""audio_tabs_b200 -- B200-native madmom-style spectral front end (one hot path of audio-tabs).

Public surface mirrors madmom's processors for this path; the compute lives in libb200spec.so
(hand-written sm_100a CUDA behind the C ABI of include/b200spec.h).  No CPU fallback.
"""
from .processors import ParallelProcessor, Processor, SequentialProcessor  # noqa: F401
from .filters import (Filterbank, LogarithmicFilterbank, PitchClassProfileFilterbank,  # noqa: F401
                      SlaneyMelFilterbank, TriangularFilter)
from .audio.signal import FramedSignal, FramedSignalProcessor, Signal, SignalProcessor  # noqa: F401
from .audio.stft import ShortTimeFourierTransform, ShortTimeFourierTransformProcessor  # noqa: F401
from .audio.spectrogram import (FilteredSpectrogram, FilteredSpectrogramProcessor,  # noqa: F401
                                LogarithmicFilteredSpectrogram, LogarithmicFilteredSpectrogramProcessor,
                                LogarithmicSpectrogram, LogarithmicSpectrogramProcessor, Spectrogram,
                                SpectrogramDifference, SpectrogramDifferenceProcessor, SpectrogramProcessor)
from .audio.chroma import FoldedChroma, FoldedChromaProcessor, PitchClassProfile  # noqa: F401

__version__ = "0.1.0"
