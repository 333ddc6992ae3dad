"""Ready-made front ends with the exact parameters madmom's feature processors use.

* ``rnn_beat_frontend()``   -- RNNBeatProcessor() pre-processor (reference: grid/beats.py:71-75):
  frames 1024/2048/4096, hop 441, 3/6/12 bands per octave, log10(1+x), positive diff at ratio 0.5,
  stacked -> (T, 314).
* ``rnn_onset_frontend()``  -- RNNOnsetProcessor() pre-processor -> (T, 266).
* ``beat_specs()`` / ``onset_specs()`` / ``log_filt_spec()`` build the ``ResolutionSpec`` lists the
  batch engine (``plan.FrontEnd``) consumes for the same chains.
"""
from __future__ import annotations

import numpy as np

from .audio.signal import FramedSignalProcessor, SignalProcessor
from .audio.spectrogram import (FilteredSpectrogramProcessor, LogarithmicSpectrogramProcessor,
                                SpectrogramDifferenceProcessor, _diff_frames)
from .audio.stft import ShortTimeFourierTransformProcessor, fft_frequencies
from .filters import LogarithmicFilterbank, fold_classes
from .plan import ResolutionSpec
from .processors import ParallelProcessor, Processor, SequentialProcessor

SAMPLE_RATE = 44100


class MultiResolutionFrontEnd(Processor):
    """All resolutions of a beat/onset front end in ONE stacked device buffer (no host hstack).

    Drop-in for the ``(ParallelProcessor(multi), np.hstack)`` tail of madmom's RNN pre-processors.
    """

    def __init__(self, specs, sample_rate=SAMPLE_RATE):
        self.specs = list(specs)
        self.sample_rate = sample_rate
        self._fe = {}

    def _front_end(self, device, dtype):
        from .plan import FrontEnd
        key = (device, dtype)
        if key not in self._fe:
            self._fe[key] = FrontEnd(self.specs, device=device, dtype=dtype, channels=1)
        return self._fe[key]

    def process(self, data, **kwargs):
        import torch
        from .engine import _device_index, _signal_tensor
        from .plan import Packed
        dev = torch.device("cuda", _device_index())
        sig, dtype = _signal_tensor(data, dev)
        if sig.ndim != 1:
            raise ValueError("frames must be a 2D array or iterable, got a %d-channel signal" % sig.shape[1])
        fe = self._front_end(dev.index, dtype)
        packed = Packed(sig, [sig.shape[0]], fe.hop_size)
        # Signal(norm=True) on a device tensor: fused as a per-clip gain, exactly as engine.run_chain does
        norm = bool(getattr(data, "norm", False)) and not isinstance(data, np.ndarray)
        scale = fe.peak_scales(packed, eps=0.0) if norm else None
        out = fe.run_packed(packed, clip_scale=scale)
        return out.cpu().numpy()


def log_filt_spec(frame_size, hop_size=441.0, num_bands=12, fmin=30.0, fmax=17000.0, norm_filters=True,
                  unique_filters=True, mul=1.0, add=1.0, diff_ratio=None, positive_diffs=True,
                  sample_rate=SAMPLE_RATE, int16=False, fold=False, origin=0, diff_max_bins=0):
    """ResolutionSpec of one madmom log-filtered-spectrogram chain (optionally with diff / chroma fold)."""
    window = np.hanning(frame_size)
    fft_window = window / 32767.0 if int16 else window
    fb = LogarithmicFilterbank(fft_frequencies(frame_size >> 1, sample_rate), num_bands=num_bands, fmin=fmin,
                               fmax=fmax, norm_filters=norm_filters, unique_filters=unique_filters)
    k = _diff_frames(diff_ratio, hop_size, frame_size, window) if diff_ratio is not None else 0
    extra = {}
    if fold:
        extra = dict(proj_classes=fold_classes(fb.center_frequencies, 12), num_classes=12)
    return ResolutionSpec(frame_size=frame_size, hop_size=hop_size, origin=origin, fft_window=fft_window,
                          filterbank=fb, log=True, mul=mul, add=add, diff_frames=k,
                          positive_diffs=positive_diffs and k > 0, diff_max_bins=diff_max_bins if k > 0 else 0, **extra)


def beat_specs(sample_rate=SAMPLE_RATE, int16=False):
    return [log_filt_spec(f, 441.0, nb, 30.0, 17000.0, mul=1.0, add=1.0, diff_ratio=0.5, sample_rate=sample_rate,
                          int16=int16) for f, nb in zip([1024, 2048, 4096], [3, 6, 12])]


def onset_specs(sample_rate=SAMPLE_RATE, int16=False):
    return [log_filt_spec(f, 441.0, 6, 30.0, 17000.0, mul=5.0, add=1.0, diff_ratio=0.25, sample_rate=sample_rate,
                          int16=int16) for f in [1024, 2048, 4096]]


def _multi_chain(frame_sizes, num_bands, mul, diff_ratio):
    multi = ParallelProcessor([])
    for frame_size, nb in zip(frame_sizes, num_bands):
        multi.append(SequentialProcessor((
            FramedSignalProcessor(frame_size=frame_size, fps=100),
            ShortTimeFourierTransformProcessor(),
            FilteredSpectrogramProcessor(num_bands=nb, fmin=30, fmax=17000, norm_filters=True),
            LogarithmicSpectrogramProcessor(mul=mul, add=1),
            SpectrogramDifferenceProcessor(diff_ratio=diff_ratio, positive_diffs=True, stack_diffs=np.hstack),
        )))
    return SequentialProcessor((SignalProcessor(num_channels=1, sample_rate=SAMPLE_RATE), multi, np.hstack))


def rnn_beat_frontend():
    """The madmom-shaped processor graph of RNNBeatProcessor's pre-processor (per-branch fusion)."""
    return _multi_chain([1024, 2048, 4096], [3, 6, 12], 1, 0.5)


def rnn_onset_frontend():
    return _multi_chain([1024, 2048, 4096], [6, 6, 6], 5, 0.25)


def rnn_beat_frontend_fused():
    """Same numbers as ``rnn_beat_frontend()`` but one stacked device buffer for all resolutions."""
    return SequentialProcessor((SignalProcessor(num_channels=1, sample_rate=SAMPLE_RATE),
                                MultiResolutionFrontEnd(beat_specs())))
