"""Spectrogram family (madmom 0.16.1 ``madmom/audio/spectrogram.py``) on the fused CUDA kernels.

Classes and processors keep madmom's constructor keywords, defaults, attributes and exception
types.  The reference reaches them through RNNBeatProcessor
(/root/reference/backend/app/services/grid/beats.py:74), DeepChromaProcessor
(chords/extract.py:54, chords/deep_chords.py:48) and CNNKeyRecognitionProcessor (theory/key.py:101).
All stages are lazy (see lazy.py): materialising the last stage of an intact chain runs
framing+window+FFT+magnitude+filterbank+log10+difference in one kernel.
"""
from __future__ import annotations

import inspect

import numpy as np

from ..filters import (A4, FMAX, FMIN, NORM_FILTERS, NUM_BANDS, UNIQUE_FILTERS, Filterbank,
                       LogarithmicFilterbank)
from ..processors import Processor
from .lazy import LazyArray
from .stft import ShortTimeFourierTransform

MUL, ADD, LOG = 1.0, 1.0, np.log10
DIFF_RATIO, DIFF_FRAMES, DIFF_MAX_BINS, POSITIVE_DIFFS = 0.5, None, None, False


class _Stage(LazyArray):
    """Common attribute plumbing: every stage exposes ``stft``, ``frames`` and ``bin_frequencies``."""
    stft = None
    source = None          # upstream stage (LazyArray) or a plain ndarray

    @property
    def frames(self):
        return self.stft.frames if self.stft is not None else None

    @property
    def num_frames(self):
        return self.shape[0]

    @property
    def num_bins(self):
        return self.shape[1]

    def _compute_tensor(self):
        from ..engine import run_chain
        return run_chain(self)


class Spectrogram(_Stage):
    """np.abs(stft) (madmom Spectrogram)."""

    def __init__(self, stft, **kwargs):
        if isinstance(stft, Spectrogram):
            self.__dict__.update(stft.__dict__)
            return
        if isinstance(stft, np.ndarray) and not np.iscomplexobj(stft) and stft.ndim == 2:
            # already a magnitude spectrogram held on the host
            self.stft, self.source = None, np.ascontiguousarray(stft, dtype=np.float32)
            self.bin_frequencies = kwargs.get("bin_frequencies", np.arange(stft.shape[1], dtype=float))
            return
        if not isinstance(stft, ShortTimeFourierTransform):
            stft = ShortTimeFourierTransform(stft, **kwargs)
        self.stft = stft
        self.source = stft
        self.bin_frequencies = stft.bin_frequencies

    def _result_shape(self):
        return self.source.shape

    def diff(self, **kwargs):
        return SpectrogramDifference(self, **kwargs)

    def filter(self, **kwargs):
        return FilteredSpectrogram(self, **kwargs)

    def log(self, **kwargs):
        return LogarithmicSpectrogram(self, **kwargs)


class SpectrogramProcessor(Processor):
    def __init__(self, **kwargs):
        pass

    def process(self, data, **kwargs):
        return Spectrogram(data, **kwargs)


def _as_spectrogram(data, **kwargs):
    if isinstance(data, _Stage):
        return data
    return Spectrogram(data, **kwargs)


class FilteredSpectrogram(_Stage):
    """np.dot(spec, filterbank) (madmom FilteredSpectrogram)."""

    def __init__(self, spectrogram, filterbank=LogarithmicFilterbank, num_bands=NUM_BANDS, fmin=FMIN,
                 fmax=FMAX, fref=A4, norm_filters=NORM_FILTERS, unique_filters=UNIQUE_FILTERS, **kwargs):
        spectrogram = _as_spectrogram(spectrogram, **kwargs)
        if inspect.isclass(filterbank) and issubclass(filterbank, Filterbank):
            filterbank = filterbank(spectrogram.bin_frequencies, num_bands=num_bands, fmin=fmin, fmax=fmax,
                                    fref=fref, norm_filters=norm_filters, unique_filters=unique_filters)
        if not isinstance(filterbank, Filterbank):
            raise TypeError("not a Filterbank type or instance: %s" % filterbank)
        if filterbank.shape[0] != spectrogram.shape[1]:
            raise ValueError("filterbank has %d bins, spectrogram %d" % (filterbank.shape[0], spectrogram.shape[1]))
        self.source = spectrogram
        self.stft = spectrogram.stft
        self.filterbank = filterbank
        self.bin_frequencies = filterbank.center_frequencies

    def _result_shape(self):
        return (self.source.shape[0], self.filterbank.shape[1])


class FilteredSpectrogramProcessor(Processor):
    def __init__(self, filterbank=LogarithmicFilterbank, num_bands=NUM_BANDS, fmin=FMIN, fmax=FMAX, fref=A4,
                 norm_filters=NORM_FILTERS, unique_filters=UNIQUE_FILTERS, **kwargs):
        self.filterbank = filterbank
        self.num_bands, self.fmin, self.fmax, self.fref = num_bands, fmin, fmax, fref
        self.norm_filters, self.unique_filters = norm_filters, unique_filters

    def process(self, data, **kwargs):
        args = dict(filterbank=self.filterbank, num_bands=self.num_bands, fmin=self.fmin, fmax=self.fmax,
                    fref=self.fref, norm_filters=self.norm_filters, unique_filters=self.unique_filters)
        args.update(kwargs)
        out = FilteredSpectrogram(data, **args)
        self.filterbank = out.filterbank        # cache the built filterbank, like madmom
        return out


# log functions the device evaluates: out = scale * log10(mul * x + add + shift)
_LOG_FORMS = {np.log10: (1.0, 0.0), np.log: (float(np.log(10.0)), 0.0), np.log2: (float(1.0 / np.log10(2.0)), 0.0),
              np.log1p: (float(np.log(10.0)), 1.0)}


class LogarithmicSpectrogram(_Stage):
    """log(mul * spec + add) (madmom LogarithmicSpectrogram); log is np.log10 (madmom's default), np.log,
    np.log2 or np.log1p -- all evaluated on the device as a scaled log10."""

    def __init__(self, spectrogram, log=LOG, mul=MUL, add=ADD, **kwargs):
        spectrogram = _as_spectrogram(spectrogram, **kwargs)
        if log not in _LOG_FORMS:
            raise ValueError("log must be np.log10, np.log, np.log2 or np.log1p on the device (got %r)" % (log,))
        self.log = log
        self.log_scale, self.log_shift = _LOG_FORMS[log]
        self.source = spectrogram
        self.stft = spectrogram.stft
        self.filterbank = getattr(spectrogram, "filterbank", None)
        self.bin_frequencies = spectrogram.bin_frequencies
        self.mul, self.add = mul, add

    def _result_shape(self):
        return self.source.shape


class LogarithmicSpectrogramProcessor(Processor):
    def __init__(self, log=LOG, mul=MUL, add=ADD, **kwargs):
        self.log, self.mul, self.add = log, mul, add

    def process(self, data, **kwargs):
        args = dict(log=self.log, mul=self.mul, add=self.add)
        args.update(kwargs)
        return LogarithmicSpectrogram(data, **args)


class LogarithmicFilteredSpectrogram(LogarithmicSpectrogram):
    def __init__(self, spectrogram, filterbank=LogarithmicFilterbank, num_bands=NUM_BANDS, fmin=FMIN, fmax=FMAX,
                 fref=A4, norm_filters=NORM_FILTERS, unique_filters=UNIQUE_FILTERS, mul=MUL, add=ADD, **kwargs):
        if not isinstance(spectrogram, FilteredSpectrogram):
            spectrogram = FilteredSpectrogram(spectrogram, filterbank=filterbank, num_bands=num_bands, fmin=fmin,
                                              fmax=fmax, fref=fref, norm_filters=norm_filters,
                                              unique_filters=unique_filters, **kwargs)
        LogarithmicSpectrogram.__init__(self, spectrogram, mul=mul, add=add)


class LogarithmicFilteredSpectrogramProcessor(Processor):
    def __init__(self, filterbank=LogarithmicFilterbank, num_bands=NUM_BANDS, fmin=FMIN, fmax=FMAX, fref=A4,
                 norm_filters=NORM_FILTERS, unique_filters=UNIQUE_FILTERS, mul=MUL, add=ADD, **kwargs):
        self.filterbank = filterbank
        self.num_bands, self.fmin, self.fmax, self.fref = num_bands, fmin, fmax, fref
        self.norm_filters, self.unique_filters = norm_filters, unique_filters
        self.mul, self.add = mul, add

    def process(self, data, **kwargs):
        args = dict(filterbank=self.filterbank, num_bands=self.num_bands, fmin=self.fmin, fmax=self.fmax,
                    fref=self.fref, norm_filters=self.norm_filters, unique_filters=self.unique_filters,
                    mul=self.mul, add=self.add)
        args.update(kwargs)
        out = LogarithmicFilteredSpectrogram(data, **args)
        self.filterbank = out.filterbank
        return out


def _diff_frames(diff_ratio, hop_size, frame_size, window=np.hanning):
    """madmom.audio.spectrogram._diff_frames (host-side integer geometry, bit-exact)."""
    if hasattr(window, "__call__"):
        window = window(frame_size)
    sample = np.argmax(window > float(diff_ratio) * max(window))
    diff_samples = len(window) / 2 - sample
    return int(max(1, round(diff_samples / hop_size)))


class SpectrogramDifference(_Stage):
    """Lagged (positive) first-order difference (madmom SpectrogramDifference)."""

    def __init__(self, spectrogram, diff_ratio=DIFF_RATIO, diff_frames=DIFF_FRAMES, diff_max_bins=DIFF_MAX_BINS,
                 positive_diffs=POSITIVE_DIFFS, keep_dims=True, **kwargs):
        spectrogram = _as_spectrogram(spectrogram, **kwargs)
        if diff_frames is None:
            diff_frames = _diff_frames(diff_ratio, frame_size=spectrogram.stft.frames.frame_size,
                                       hop_size=spectrogram.stft.frames.hop_size, window=spectrogram.stft.window)
        if diff_frames < 1:
            raise ValueError("number of `diff_frames` must be >= 1")
        if diff_max_bins is not None and not 0 <= int(diff_max_bins) <= 64:
            raise ValueError("diff_max_bins must be within [0, 64]")
        self.source = spectrogram
        self.spectrogram = spectrogram
        self.stft = spectrogram.stft
        self.filterbank = getattr(spectrogram, "filterbank", None)
        self.bin_frequencies = spectrogram.bin_frequencies
        self.diff_ratio, self.diff_frames, self.diff_max_bins = diff_ratio, int(diff_frames), diff_max_bins
        self.positive_diffs = positive_diffs

    def _result_shape(self):
        return self.source.shape

    def positive_diff(self):
        return np.maximum(np.asarray(self), 0)


class StackedDifference(_Stage):
    """``np.hstack((spec, diff))`` produced directly by the fused kernel."""

    def __init__(self, diff):
        self.source = diff
        self.stft = diff.stft
        self.diff = diff
        self.bin_frequencies = diff.bin_frequencies

    def _result_shape(self):
        t, b = self.source.shape
        return (t, 2 * b)


class BufferedDifference(_Stage):
    """Result of an online (``reset=False``) SpectrogramDifferenceProcessor call: rows already computed on the device."""

    def __init__(self, tensor, like, diff_frames, spectrogram=None):
        self._tensor = tensor
        self.source = like
        self.stft = like.stft
        self.spectrogram = spectrogram if spectrogram is not None else like
        self.filterbank = getattr(like, "filterbank", None)
        self.bin_frequencies = like.bin_frequencies
        self.diff_frames = diff_frames

    def _compute_tensor(self):
        return self._tensor

    def _result_shape(self):
        return tuple(self._tensor.shape)


class SpectrogramDifferenceProcessor(Processor):
    """madmom's SpectrogramDifferenceProcessor.  ``reset=True`` (the default, what every caller of the reference
    uses) is the offline behaviour: the first ``diff_frames`` rows of the difference are 0 and the whole chain is
    one fused launch.  ``reset=False`` continues from the rows of the previous calls like madmom's BufferProcessor
    does: the buffer is as long as the first call's rows plus ``diff_frames``, new rows are shifted in at its end
    and the result covers the whole buffer behind its first ``diff_frames`` rows (difference kernel on the
    device-resident buffer)."""

    def __init__(self, diff_ratio=DIFF_RATIO, diff_frames=DIFF_FRAMES, diff_max_bins=DIFF_MAX_BINS,
                 positive_diffs=POSITIVE_DIFFS, stack_diffs=None, **kwargs):
        self.diff_ratio, self.diff_frames, self.diff_max_bins = diff_ratio, diff_frames, diff_max_bins
        self.positive_diffs, self.stack_diffs = positive_diffs, stack_diffs
        self._buffer = None

    def __getstate__(self):
        state = self.__dict__.copy()
        state.pop("_buffer", None)
        return state

    def __setstate__(self, state):
        self.__dict__.update(state)
        self._buffer = None

    def _history(self, num_bands):
        """The buffer as a device tensor (diff_frames + T0, B): inf rows, then the rows seen so far."""
        import torch
        kind, held = self._buffer
        if kind == "rows":
            return held
        first = held.tensor()                       # the offline call's spectrogram rows
        if first.ndim != 2 or first.shape[1] != num_bands:
            raise ValueError("could not broadcast input array from shape (%d,) into shape (%d,)"
                             % (num_bands, first.shape[-1]))
        init = torch.full((self.diff_frames, num_bands), float("inf"), dtype=torch.float32, device=first.device)
        return torch.cat((init, first.to(torch.float32)))

    def _process_online(self, data, args):
        import torch
        from ..engine import buffered_difference
        x = data.tensor()
        buf = self._history(x.shape[1])
        n = x.shape[0]
        if n > buf.shape[0]:                        # madmom: buffer[-n:] = data does not fit
            raise ValueError("could not broadcast input array from shape (%d,%d) into shape (%d,%d)"
                             % (n, x.shape[1], buf.shape[0], buf.shape[1]))
        buf = torch.cat((buf[n:], x.to(device=buf.device, dtype=torch.float32)))
        self._buffer = ("rows", buf)
        kd = self.diff_frames
        stacked = self.stack_diffs is np.hstack
        out = buffered_difference(buf, kd, bool(args["positive_diffs"]), int(args["diff_max_bins"] or 0), stacked)
        if self.stack_diffs is None or stacked:
            return BufferedDifference(out, data, kd)
        B = buf.shape[1]
        return self.stack_diffs((buf[kd:].cpu().numpy(), out[:, -B:].cpu().numpy()))

    def process(self, data, reset=True, **kwargs):
        args = dict(diff_ratio=self.diff_ratio, diff_frames=self.diff_frames, diff_max_bins=self.diff_max_bins,
                    positive_diffs=self.positive_diffs)
        args.update(kwargs)
        data = _as_spectrogram(data)
        if self.diff_frames is None:
            self.diff_frames = _diff_frames(args["diff_ratio"], frame_size=data.stft.frames.frame_size,
                                            hop_size=data.stft.frames.hop_size, window=data.stft.window)
            args["diff_frames"] = self.diff_frames
        if not reset and self._buffer is not None:
            return self._process_online(data, args)
        self._buffer = ("stage", data)      # these rows are the history of a later reset=False call (nothing is computed now)
        diff = SpectrogramDifference(data, **args)
        if self.stack_diffs is None:
            return diff
        if self.stack_diffs is np.hstack:
            return StackedDifference(diff)
        return self.stack_diffs((np.asarray(data), np.asarray(diff)))
