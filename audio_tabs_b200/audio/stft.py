"""ShortTimeFourierTransform equivalents (madmom 0.16.1 ``madmom/audio/stft.py``).

Replaces ``stft()`` -- the per-frame ``scipy.fftpack.fft`` Python loop that takes ~80 % of the
front-end CPU time, reached from
/root/reference/backend/app/services/grid/beats.py:74 (RNNBeatProcessor) -- with K1 of
libb200spec.so.  The result is lazy: followed by the filterbank / log / difference processors it
is never written to memory at all (the fused kernel keeps it in shared memory).
"""
from __future__ import annotations

import numpy as np

from ..processors import Processor
from .lazy import LazyArray
from .signal import FramedSignal

STFT_DTYPE = np.complex64


def fft_frequencies(num_fft_bins, sample_rate):
    return np.fft.fftfreq(num_fft_bins * 2, 1.0 / sample_rate)[:num_fft_bins]


def derive_fft_window(window, frame_size, signal_dtype):
    """(window, fft_window) exactly as madmom's ShortTimeFourierTransform.__new__ derives them."""
    if hasattr(window, "__call__"):
        window = window(frame_size)
    try:
        max_range = float(np.iinfo(signal_dtype).max)      # integer signals: scale the window, not the data
        try:
            fft_window = window / max_range
        except TypeError:
            fft_window = np.ones(frame_size) / max_range
    except ValueError:
        fft_window = window
    return window, fft_window


class ShortTimeFourierTransform(LazyArray):
    _result_dtype = np.dtype(STFT_DTYPE)

    def __init__(self, frames, window=np.hanning, fft_size=None, circular_shift=False, include_nyquist=False,
                 fft_window=None, fftw=None, **kwargs):
        if not isinstance(frames, FramedSignal):
            frames = FramedSignal(frames, **kwargs)
        if frames.ndim != 2:
            raise ValueError("frames must be a 2D array or iterable, got %s with shape %s."
                             % (type(frames), frames.shape))
        frame_size = frames.shape[1]
        if fft_window is None:
            window, fft_window = derive_fft_window(window, frame_size, frames.signal.dtype)
        elif hasattr(window, "__call__"):
            window = window(frame_size)
        if fft_window is None:
            fft_window = np.ones(frame_size)               # window=None: rectangular
        fft_size = frame_size if fft_size is None else int(fft_size)
        if fft_size not in (1024, 2048, 4096, 8192):
            raise ValueError("the CUDA path transforms 1024, 2048, 4096 or 8192 points (fft_size %d); give one of them "
                             "as fft_size for another frame_size" % fft_size)
        if circular_shift and fft_size != frame_size:
            raise ValueError("circular_shift=True with fft_size != frame_size is not supported by the CUDA path")
        self.frames = frames
        self.window = window
        self.fft_window = fft_window
        self.fft_size = fft_size
        self.circular_shift = bool(circular_shift)
        self.include_nyquist = bool(include_nyquist)
        self.fftw = None
        nb = (fft_size >> 1) + int(self.include_nyquist)
        if not frames.signal.sample_rate:
            self.bin_frequencies = np.arange(nb, dtype=float)
        elif self.include_nyquist:
            self.bin_frequencies = np.fft.rfftfreq(fft_size, 1.0 / frames.signal.sample_rate)
        else:
            self.bin_frequencies = fft_frequencies(fft_size >> 1, frames.signal.sample_rate)

    def kernel_window_and_origin(self):
        """(window of fft_size points, origin) that make the kernel's frame [int(n hop) - fft_size/2 - origin, + fft_size)
        start where madmom's frame starts: scipy.fftpack.fft(frame * window, fft_size) zero-pads (or cuts) the windowed
        frame at its END, so the window is padded / cut the same way and the origin absorbs the difference of the two
        half lengths."""
        win = np.asarray(self.fft_window, dtype=np.float64)
        frame_size = self.frames.frame_size
        if self.fft_size > frame_size:
            win = np.concatenate((win, np.zeros(self.fft_size - frame_size)))
        elif self.fft_size < frame_size:
            win = win[:self.fft_size]
        return win, self.frames.origin + frame_size // 2 - self.fft_size // 2

    def _result_shape(self):
        return (self.frames.num_frames, self.num_bins)

    @property
    def num_frames(self):
        return self.frames.num_frames

    @property
    def num_bins(self):
        return (self.fft_size >> 1) + int(self.include_nyquist)

    def _compute_tensor(self):
        from ..engine import run_chain
        return run_chain(self, kind="stft")

    def spec(self, **kwargs):
        from .spectrogram import Spectrogram
        return Spectrogram(self, **kwargs)

    def phase(self, **kwargs):
        return np.angle(np.asarray(self))


STFT = ShortTimeFourierTransform


class ShortTimeFourierTransformProcessor(Processor):
    def __init__(self, window=np.hanning, fft_size=None, circular_shift=False, include_nyquist=False, **kwargs):
        self.window = window
        self.fft_size = fft_size
        self.circular_shift = circular_shift
        self.include_nyquist = include_nyquist
        self.fft_window = None      # cached after the first call, like madmom
        self.fftw = None

    def process(self, data, **kwargs):
        args = dict(window=self.window, fft_size=self.fft_size, circular_shift=self.circular_shift,
                    include_nyquist=self.include_nyquist, fft_window=self.fft_window)
        args.update(kwargs)
        out = ShortTimeFourierTransform(data, **args)
        self.fft_window = out.fft_window
        return out


STFTProcessor = ShortTimeFourierTransformProcessor
