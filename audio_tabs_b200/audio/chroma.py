"""Chroma projections of the front end (madmom ``audio/hpcp.py``, ``audio/chroma.py``).

* ``PitchClassProfile``: ``np.dot(spec, PitchClassProfileFilterbank)`` on linear FFT bins
  (madmom.audio.hpcp.PitchClassProfile; class 0 = pitch class of ``fref`` = A).
* ``FoldedChroma``: octave fold of a (log-)filtered spectrogram onto 12 classes with class 0 = C,
  the orientation the reference's chroma consumers assume
  (/root/reference/backend/app/services/chords/template.py:7,76-80, chords/extract.py:60).
* ``deep_chroma_frontend`` / ``cnn_key_frontend`` / ``cnn_chord_frontend``: the pre-network part of
  madmom's DeepChromaProcessor / CNNKeyRecognitionProcessor / CNNChordFeatureProcessor
  (chords/extract.py:54, theory/key.py:101, chords/deep_chords.py:79-81).
"""
from __future__ import annotations

import numpy as np

from ..filters import A4, PitchClassProfileFilterbank, fold_classes
from ..processors import Processor, SequentialProcessor
from .signal import FramedSignalProcessor, SignalProcessor
from .spectrogram import (FilteredSpectrogram, LogarithmicFilteredSpectrogramProcessor, _Stage, _as_spectrogram)
from .stft import ShortTimeFourierTransformProcessor


class PitchClassProfile(FilteredSpectrogram):
    def __init__(self, spectrogram, filterbank=PitchClassProfileFilterbank, num_classes=12, fmin=100.0,
                 fmax=5000.0, fref=A4, **kwargs):
        spectrogram = _as_spectrogram(spectrogram, **kwargs)
        if isinstance(filterbank, type):
            filterbank = filterbank(spectrogram.bin_frequencies, num_classes=num_classes, fmin=fmin, fmax=fmax,
                                    fref=fref)
        FilteredSpectrogram.__init__(self, spectrogram, filterbank=filterbank)
        self.bin_frequencies = None


class FoldedChroma(_Stage):
    """chroma[:, c] = sum of the bands whose centre frequency has pitch class c (0 = C)."""

    def __init__(self, spectrogram, num_classes=12, classes=None):
        spectrogram = _as_spectrogram(spectrogram)
        if classes is None:
            classes = fold_classes(spectrogram.bin_frequencies, num_classes)
        self.source = spectrogram
        self.stft = spectrogram.stft
        self.classes = np.asarray(classes, dtype=np.int64)
        self.num_classes = int(num_classes)
        self.bin_frequencies = None

    def _result_shape(self):
        return (self.source.shape[0], self.num_classes)


class FoldedChromaProcessor(Processor):
    def __init__(self, num_classes=12, **kwargs):
        self.num_classes = num_classes

    def process(self, data, **kwargs):
        return FoldedChroma(data, num_classes=self.num_classes)


def log_filtered_chain(frame_size, fps=None, hop_size=441.0, num_bands=24, fmin=65.0, fmax=2100.0,
                       unique_filters=True, mul=1.0, add=1.0, sample_rate=44100):
    return SequentialProcessor((
        SignalProcessor(num_channels=1, sample_rate=sample_rate),
        FramedSignalProcessor(frame_size=frame_size, fps=fps, hop_size=hop_size),
        ShortTimeFourierTransformProcessor(),
        LogarithmicFilteredSpectrogramProcessor(num_bands=num_bands, fmin=fmin, fmax=fmax,
                                                unique_filters=unique_filters, mul=mul, add=add),
    ))


def deep_chroma_frontend(fmin=65.0, fmax=2100.0, unique_filters=True):
    """DeepChromaProcessor() up to (not including) the context stacking and the DNN: (T, 105)."""
    return log_filtered_chain(8192, fps=10, num_bands=24, fmin=fmin, fmax=fmax, unique_filters=unique_filters)


def cnn_key_frontend():
    """CNNKeyRecognitionProcessor() front end: 8192 @ fps 5, 24 bands/octave 65-2100 Hz -> (T, 105)."""
    return log_filtered_chain(8192, fps=5, num_bands=24, fmin=65.0, fmax=2100.0, unique_filters=True)


def cnn_chord_frontend():
    """CNNChordFeatureProcessor() front end: 8192 @ fps 10, 24 bands/octave 60-2600 Hz -> (T, 113)."""
    return log_filtered_chain(8192, fps=10, num_bands=24, fmin=60.0, fmax=2600.0, unique_filters=True)


def chord_chroma_frontend(frame_size=4096, hop_size=441.0, fps=None, num_bands=24, fmin=65.0, fmax=2100.0):
    """BASELINE config 3: frame-4096 STFT -> log filterbank -> 12-bin folded chroma."""
    chain = log_filtered_chain(frame_size, fps=fps, hop_size=hop_size, num_bands=num_bands, fmin=fmin, fmax=fmax)
    return SequentialProcessor((chain, FoldedChromaProcessor(12)))


def context_stack_device(spec, context=15, frame_off=None):
    """DeepChromaProcessor's ``FramedSignal(frame_size=15, hop_size=1)`` + ``_dcp_flatten`` on the GPU:
    (T, B) CUDA tensor -> (T, context*B), zero padded at both ends of every clip (``frame_off``: (n+1,)
    int64 row offsets of the clips packed in ``spec``; default one clip)."""
    import ctypes as C
    import torch
    from .. import _ffi
    T, B = spec.shape
    if frame_off is None:
        frame_off = torch.tensor([0, T], dtype=torch.int64, device=spec.device)
    out = torch.empty((T, context * B), dtype=torch.float32, device=spec.device)
    _ffi.check(_ffi.lib().b200spec_context_stack(C.c_void_p(spec.data_ptr()), spec.stride(0), B,
                                                 C.c_void_p(frame_off.data_ptr()), frame_off.numel() - 1, T,
                                                 int(context), C.c_void_p(out.data_ptr()),
                                                 C.c_void_p(torch.cuda.current_stream(spec.device).cuda_stream)))
    return out


def context_stack(spec, context=15):
    """Host form of the same stacking (what madmom does): pure data movement, (T, B) -> (T, context*B)
    with zero padding at both ends.  CUDA tensors go through :func:`context_stack_device`."""
    try:
        import torch
        if isinstance(spec, torch.Tensor) and spec.is_cuda:
            return context_stack_device(spec, context)
    except ImportError:  # pragma: no cover
        pass
    spec = np.asarray(spec)
    T, B = spec.shape
    half = context // 2
    padded = np.zeros((T + 2 * half, B), dtype=spec.dtype)
    padded[half:half + T] = spec
    win = np.lib.stride_tricks.sliding_window_view(padded, (context, B))[:, 0]
    return np.ascontiguousarray(win.reshape(T, context * B))
