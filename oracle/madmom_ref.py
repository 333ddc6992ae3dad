"""CPU oracle: numpy restatement of the madmom 0.16.1 spectral front end.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package may import this
module; only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs use it, and only as the checker or
as the CPU baseline being timed.

PARITY UNPINNED.  The algorithm lives in the third-party package
``madmom==0.16.1`` (pinned at /root/reference/backend/requirements.txt:15),
which is neither vendored under /root/reference nor installable here (no
network).  The reference's own tests hold no golden vectors for this path
(SURVEY.md §8c).  This file restates madmom's published algorithm
(madmom/audio/{signal,stft,spectrogram,filters,hpcp,chroma}.py and
madmom/processors.py); what pins it are the structural constants of madmom
0.16.1 checked in tests/test_oracle_pins.py (band counts 81/108/105/113,
21/45/91 -> 314, 39/45/49 -> 266, diff_frames 1/1/2 and 1/2/3, 281 frames for
a 123481-sample signal) and analytic known-answer tests.

The reference call sites that fix which parameters matter:
  /root/reference/backend/app/services/grid/beats.py:71-82      RNNBeatProcessor
  /root/reference/backend/app/services/chords/extract.py:50-55  DeepChromaProcessor
  /root/reference/backend/app/services/chords/deep_chords.py:44-49,74-81
  /root/reference/backend/app/services/theory/key.py:99-101     CNNKeyRecognitionProcessor

Arithmetic follows madmom exactly: float64 Hann window, per-frame
``scipy.fftpack.fft`` on the float64 product, complex64 store, ``np.abs`` ->
float32, float32 filterbank ``np.dot``, ``np.log10(mul * x + add)`` in float32,
lagged difference with ``np.maximum(.., 0)``, ``np.hstack``.
"""
from __future__ import annotations

import numpy as np
from scipy import fftpack

FILTER_DTYPE = np.float32
A4 = 440.0

# ----------------------------------------------------------------------------
# madmom/processors.py
# ----------------------------------------------------------------------------


class Processor:
    """madmom.processors.Processor: ``process(data, **kwargs)``; call == process."""

    def process(self, data, **kwargs):  # pragma: no cover - abstract
        raise NotImplementedError("Must be implemented by subclass.")

    def __call__(self, *args, **kwargs):
        return self.process(*args, **kwargs)


def _process(process_tuple):
    # madmom.processors._process: Processors get kwargs, plain callables do not
    proc, data, kwargs = process_tuple
    if isinstance(proc, Processor):
        return proc(data, **kwargs)
    return proc(data)


class SequentialProcessor(Processor):
    """madmom.processors.SequentialProcessor: left fold over the processors."""

    def __init__(self, processors):
        self.processors = list(processors)

    def process(self, data, **kwargs):
        for p in self.processors:
            data = _process((p, data, kwargs))
        return data


class ParallelProcessor(SequentialProcessor):
    """madmom.processors.ParallelProcessor with num_threads=1: serial map -> list."""

    def __init__(self, processors, num_threads=None):
        self.processors = list(processors)

    def process(self, data, **kwargs):
        return [_process((p, data, kwargs)) for p in self.processors]


# ----------------------------------------------------------------------------
# madmom/audio/signal.py
# ----------------------------------------------------------------------------


def remix(signal, num_channels):
    """madmom.audio.signal.remix (down-mix branch): mean over channels, same dtype."""
    if num_channels == signal.ndim or num_channels is None:
        return signal
    if num_channels == 1 and signal.ndim > 1:
        return np.mean(signal, axis=-1).astype(signal.dtype)
    if num_channels > 1 and signal.ndim == 1:
        return np.tile(signal[:, np.newaxis], num_channels)
    if num_channels > 1 and num_channels == signal.shape[1]:
        return signal
    raise NotImplementedError("only down-mixing to mono is restated")


class Signal:
    """Minimal stand-in for madmom.audio.signal.Signal (ndarray + sample_rate)."""

    def __init__(self, data, sample_rate=None, num_channels=None, norm=False, gain=0.0, dtype=None):
        data = np.asarray(data)
        if dtype is not None:
            data = data.astype(dtype)
        data = remix(data, num_channels)
        if norm:
            data = data.astype(np.float32) / np.max(np.abs(data))  # float result
        if gain:
            data = (data * np.power(np.sqrt(10.0), 0.1 * gain)).astype(data.dtype)
        self.data = data
        self.sample_rate = sample_rate

    def __len__(self):
        return len(self.data)

    @property
    def dtype(self):
        return self.data.dtype

    @property
    def num_samples(self):
        return len(self.data)

    @property
    def num_channels(self):
        return 1 if self.data.ndim == 1 else self.data.shape[1]


class SignalProcessor(Processor):
    def __init__(self, sample_rate=None, num_channels=None, start=None, stop=None,
                 norm=False, gain=0.0, dtype=None, **kwargs):
        self.sample_rate = sample_rate
        self.num_channels = num_channels
        self.norm = norm
        self.gain = gain
        self.dtype = dtype

    def process(self, data, **kwargs):
        if isinstance(data, Signal):
            sr = data.sample_rate if data.sample_rate is not None else self.sample_rate
            return Signal(data.data, sample_rate=sr, num_channels=self.num_channels,
                          norm=self.norm, gain=self.gain, dtype=self.dtype)
        return Signal(data, sample_rate=self.sample_rate, num_channels=self.num_channels,
                      norm=self.norm, gain=self.gain, dtype=self.dtype)


def frame_start(index, frame_size, hop_size, origin=0):
    """First sample of frame `index` (madmom.audio.signal.signal_frame geometry)."""
    ref_sample = int(index * hop_size)          # float64 product, truncated
    return ref_sample - frame_size // 2 - int(origin)


def signal_frame(signal, index, frame_size, hop_size, origin=0):
    """madmom.audio.signal.signal_frame: zero-padded frame around int(index*hop)."""
    num_samples = len(signal)
    start = frame_start(index, frame_size, hop_size, origin)
    stop = start + frame_size
    if start >= 0 and stop <= num_samples:
        return signal[start:stop]
    frame = np.zeros((frame_size,) + signal.shape[1:], dtype=signal.dtype)
    lo, hi = max(start, 0), min(stop, num_samples)
    if hi > lo:
        frame[lo - start:hi - start] = signal[lo:hi]
    return frame


def num_frames_for(num_samples, hop_size, end="normal"):
    """FramedSignal frame count: 'normal' = ceil(N/hop), 'extend' = floor(N/hop + 1)."""
    if end == "extend":
        return int(np.floor(num_samples / float(hop_size) + 1))
    if end == "normal":
        return int(np.ceil(num_samples / float(hop_size)))
    raise ValueError("end of signal handling '%s' unknown" % end)


def resolve_origin(origin, frame_size):
    if origin in ("center", "offline"):
        origin = 0
    elif origin in ("left", "past", "online"):
        origin = (frame_size - 1) / 2
    elif origin in ("right", "future", "stream"):
        origin = -(frame_size / 2)
    return int(origin)


class FramedSignal:
    """madmom.audio.signal.FramedSignal: lazy overlapping frames of a Signal."""

    def __init__(self, signal, frame_size=2048, hop_size=441.0, fps=None, origin=0,
                 end="normal", num_frames=None, **kwargs):
        if not isinstance(signal, Signal):
            signal = Signal(signal, **kwargs)
        self.signal = signal
        self.frame_size = int(frame_size)
        self.hop_size = float(hop_size)
        if fps:
            self.hop_size = self.signal.sample_rate / float(fps)
        self.origin = resolve_origin(origin, self.frame_size)
        if num_frames is None:
            num_frames = num_frames_for(len(self.signal), self.hop_size, end)
        self.num_frames = int(num_frames)

    def __len__(self):
        return self.num_frames

    def __getitem__(self, index):
        if isinstance(index, (int, np.integer)):
            if index < 0:
                index += self.num_frames
            if index < self.num_frames and index >= 0:
                return signal_frame(self.signal.data, index, self.frame_size,
                                    self.hop_size, self.origin)
            raise IndexError("end of signal reached")
        raise TypeError("only integer indexing is restated")

    def __iter__(self):
        for i in range(self.num_frames):
            yield self[i]

    @property
    def shape(self):
        shape = (self.num_frames, self.frame_size)
        if self.signal.num_channels != 1:
            shape += (self.signal.num_channels,)
        return shape

    @property
    def ndim(self):
        return len(self.shape)

    @property
    def fps(self):
        return self.signal.sample_rate / float(self.hop_size)


class FramedSignalProcessor(Processor):
    def __init__(self, frame_size=2048, hop_size=441.0, fps=None, origin=0, end="normal",
                 num_frames=None, **kwargs):
        self.frame_size = frame_size
        self.hop_size = hop_size
        self.fps = fps
        self.origin = origin
        self.end = end
        self.num_frames = num_frames

    def process(self, data, **kwargs):
        args = dict(frame_size=self.frame_size, hop_size=self.hop_size, fps=self.fps,
                    origin=self.origin, end=self.end, num_frames=self.num_frames)
        args.update(kwargs)
        return FramedSignal(data, **args)


# ----------------------------------------------------------------------------
# madmom/audio/stft.py
# ----------------------------------------------------------------------------


def fft_frequencies(num_fft_bins, sample_rate):
    return np.fft.fftfreq(num_fft_bins * 2, 1.0 / sample_rate)[:num_fft_bins]


def fft_window_for(window, frame_size, signal_dtype):
    """(window, fft_window) as ShortTimeFourierTransform.__new__ derives them."""
    if callable(window):
        window = window(frame_size)
    try:
        max_range = float(np.iinfo(signal_dtype).max)
        fft_window = window / max_range if window is not None else np.ones(frame_size) / max_range
    except ValueError:
        fft_window = window
    return window, fft_window


def stft(frames, window, fft_size=None, circular_shift=False, include_nyquist=False):
    """madmom.audio.stft.stft: per-frame float64 FFTPACK transform, complex64 store."""
    if frames.ndim != 2:
        raise ValueError("frames must be a 2D array or iterable, got %s with shape %s."
                         % (type(frames), frames.shape))
    num_frames, frame_size = frames.shape
    if fft_size is None:
        fft_size = frame_size
    num_fft_bins = fft_size >> 1
    if include_nyquist:
        num_fft_bins += 1
    if circular_shift:
        fft_shift = frame_size >> 1
    data = np.empty((num_frames, num_fft_bins), np.complex64)
    for f, frame in enumerate(frames):
        if circular_shift:
            fft_signal = np.zeros(fft_size)
            if window is not None:
                fft_signal[:fft_shift] = frame[fft_shift:] * window[fft_shift:]
                fft_signal[-fft_shift:] = frame[:fft_shift] * window[:fft_shift]
            else:
                fft_signal[:fft_shift] = frame[fft_shift:]
                fft_signal[-fft_shift:] = frame[:fft_shift]
        else:
            fft_signal = np.multiply(frame, window) if window is not None else frame
        data[f] = fftpack.fft(fft_signal, fft_size, axis=0)[:num_fft_bins]
    return data


class ShortTimeFourierTransform:
    def __init__(self, frames, window=np.hanning, fft_size=None, circular_shift=False,
                 include_nyquist=False, fft_window=None, **kwargs):
        if not isinstance(frames, FramedSignal):
            frames = FramedSignal(frames, **kwargs)
        self.frames = frames
        frame_size = frames.shape[1]
        if fft_window is None:
            window, fft_window = fft_window_for(window, frame_size, frames.signal.dtype)
        self.window = window
        self.fft_window = fft_window
        self.fft_size = fft_size if fft_size is not None else frame_size
        self.circular_shift = circular_shift
        self.include_nyquist = include_nyquist
        self.data = stft(frames, fft_window, fft_size=fft_size, circular_shift=circular_shift,
                         include_nyquist=include_nyquist)
        self.bin_frequencies = fft_frequencies(self.data.shape[1] - (1 if include_nyquist else 0),
                                               frames.signal.sample_rate)
        if include_nyquist:
            self.bin_frequencies = np.fft.rfftfreq(self.fft_size, 1.0 / frames.signal.sample_rate)

    @property
    def num_frames(self):
        return self.data.shape[0]

    @property
    def num_bins(self):
        return self.data.shape[1]

    def __array__(self, dtype=None, copy=None):
        return self.data if dtype is None else self.data.astype(dtype)


class ShortTimeFourierTransformProcessor(Processor):
    def __init__(self, window=np.hanning, fft_size=None, circular_shift=False,
                 include_nyquist=False, **kwargs):
        self.window = window
        self.fft_size = fft_size
        self.circular_shift = circular_shift
        self.include_nyquist = include_nyquist
        self.fft_window = None

    def process(self, data, **kwargs):
        out = ShortTimeFourierTransform(data, window=self.window, fft_size=self.fft_size,
                                        circular_shift=self.circular_shift,
                                        include_nyquist=self.include_nyquist,
                                        fft_window=self.fft_window, **kwargs)
        self.fft_window = out.fft_window      # cached like madmom does
        if self.fft_window is not out.window and callable(self.window):
            out.window = self.window(out.frames.frame_size)
        return out


# ----------------------------------------------------------------------------
# madmom/audio/filters.py
# ----------------------------------------------------------------------------


def hz2midi(f, fref=A4):
    return 12.0 * np.log2(np.asarray(f, dtype=float) / fref) + 69.0


def midi2hz(m, fref=A4):
    return 2.0 ** ((np.asarray(m, dtype=float) - 69.0) / 12.0) * fref


def log_frequencies(bands_per_octave, fmin, fmax, fref=A4):
    left = np.floor(np.log2(float(fmin) / fref) * bands_per_octave)
    right = np.ceil(np.log2(float(fmax) / fref) * bands_per_octave)
    frequencies = fref * 2.0 ** (np.arange(left, right) / float(bands_per_octave))
    frequencies = frequencies[np.searchsorted(frequencies, fmin):]
    frequencies = frequencies[:np.searchsorted(frequencies, fmax, "right")]
    return frequencies


def frequencies2bins(frequencies, bin_frequencies, unique_bins=False):
    frequencies = np.asarray(frequencies)
    bin_frequencies = np.asarray(bin_frequencies)
    indices = bin_frequencies.searchsorted(frequencies)
    indices = np.clip(indices, 1, len(bin_frequencies) - 1)
    left = bin_frequencies[indices - 1]
    right = bin_frequencies[indices]
    indices -= frequencies - left < right - frequencies
    if unique_bins:
        indices = np.unique(indices)
    return indices


def triangular_band_bins(bins, overlap=True):
    """TriangularFilter.band_bins: sliding (start, center, stop) triples."""
    if len(bins) < 3:
        raise ValueError("not enough bins to create a TriangularFilter")
    index = 0
    while index + 3 <= len(bins):
        start, center, stop = bins[index:index + 3]
        if not overlap:
            start = int(np.floor((center + start) / 2.0))
            stop = int(np.ceil((center + stop) / 2.0))
        if stop - start < 2:
            center = start
            stop = start + 1
        yield int(start), int(center), int(stop)
        index += 1


def triangular_filter(start, center, stop, norm=False):
    """TriangularFilter.__new__ + Filter.__new__: (data float32, start)."""
    if not start <= center < stop:
        raise ValueError("`center` must be between `start` and `stop`")
    center -= start
    stop -= start
    data = np.zeros(stop)
    data[:center] = np.linspace(0, 1, center, endpoint=False)
    data[center:] = np.linspace(1, 0, stop - center, endpoint=False)
    data = np.asarray(data, dtype=FILTER_DTYPE)
    if norm:
        data /= np.sum(data)
    return data, start


class Filterbank:
    """madmom.audio.filters.Filterbank: float32 (num_bins, num_bands) + frequencies."""

    def __init__(self, data, bin_frequencies):
        self.data = np.asarray(data, dtype=FILTER_DTYPE)
        if self.data.ndim != 2:
            raise TypeError("wrong input data for Filterbank, must be a 2D np.ndarray")
        if len(bin_frequencies) != self.data.shape[0]:
            raise ValueError("`bin_frequencies` must have the same length as the first "
                             "dimension of `data`.")
        self.bin_frequencies = np.asarray(bin_frequencies, dtype=float)

    @classmethod
    def from_filters(cls, filters, bin_frequencies):
        fb = np.zeros((len(bin_frequencies), len(filters)))
        for band_id, (filt, start) in enumerate(filters):
            band = fb[:, band_id]
            stop = start + len(filt)
            if start < 0:
                filt = filt[-start:]
                start = 0
            if stop > len(band):
                filt = filt[:-(stop - len(band))]
                stop = len(band)
            position = band[start:stop]
            np.maximum(filt, position, out=position)
        obj = cls.__new__(cls)
        Filterbank.__init__(obj, fb, bin_frequencies)
        return obj

    @property
    def shape(self):
        return self.data.shape

    @property
    def num_bins(self):
        return self.data.shape[0]

    @property
    def num_bands(self):
        return self.data.shape[1]

    @property
    def corner_frequencies(self):
        freqs = []
        for band in range(self.num_bands):
            bins = np.nonzero(self.data[:, band])[0]
            freqs.append([np.min(bins), np.max(bins)])
        return self.bin_frequencies[freqs].T

    @property
    def center_frequencies(self):
        freqs = []
        for band in range(self.num_bands):
            bins = np.nonzero(self.data[:, band])[0]
            min_bin, max_bin = np.min(bins), np.max(bins)
            if self.data[min_bin, band] == self.data[max_bin, band]:
                center = int(min_bin + (max_bin - min_bin) / 2.0)
            else:
                center = min_bin + np.argmax(self.data[min_bin:max_bin, band])
            freqs.append(center)
        return self.bin_frequencies[freqs]

    def __array__(self, dtype=None, copy=None):
        return self.data if dtype is None else self.data.astype(dtype)


class LogarithmicFilterbank(Filterbank):
    def __init__(self, bin_frequencies, num_bands=12, fmin=30.0, fmax=17000.0, fref=A4,
                 norm_filters=True, unique_filters=True, bands_per_octave=True):
        frequencies = log_frequencies(num_bands, fmin, fmax, fref)
        bins = frequencies2bins(frequencies, bin_frequencies, unique_bins=unique_filters)
        filters = [triangular_filter(s, c, e, norm=norm_filters)
                   for s, c, e in triangular_band_bins(bins, overlap=True)]
        fb = Filterbank.from_filters(filters, bin_frequencies)
        Filterbank.__init__(self, fb.data, bin_frequencies)
        self.fref = fref
        self.norm_filters = norm_filters
        self.unique_filters = unique_filters


class PitchClassProfileFilterbank(Filterbank):
    """madmom.audio.filters.PitchClassProfileFilterbank (class 0 = pitch class of fref)."""

    def __init__(self, bin_frequencies, num_classes=12, fmin=100.0, fmax=5000.0, fref=A4):
        bin_frequencies = np.asarray(bin_frequencies, dtype=float)
        fb = np.zeros((len(bin_frequencies), num_classes))
        with np.errstate(divide="ignore"):
            log_dev = np.log2(bin_frequencies / fref)
        min_bin = np.searchsorted(bin_frequencies, fmin)
        max_bin = np.searchsorted(bin_frequencies, fmax, "right")
        for b in range(min_bin, max_bin):
            cls_ = int(np.round(num_classes * log_dev[b])) % num_classes
            fb[b, cls_] = 1
        Filterbank.__init__(self, fb, bin_frequencies)
        self.fref = fref


# ----------------------------------------------------------------------------
# madmom/audio/spectrogram.py
# ----------------------------------------------------------------------------


class _Spec:
    """Carrier for the attributes the next madmom stage reads."""

    def __init__(self, data, stft=None, bin_frequencies=None, **extra):
        self.data = data
        self.stft = stft
        self.frames = stft.frames if stft is not None else None
        self.bin_frequencies = bin_frequencies
        for k, v in extra.items():
            setattr(self, k, v)

    @property
    def shape(self):
        return self.data.shape

    @property
    def num_frames(self):
        return self.data.shape[0]

    @property
    def num_bins(self):
        return self.data.shape[1]

    def __array__(self, dtype=None, copy=None):
        return self.data if dtype is None else self.data.astype(dtype)


def spectrogram(stft_obj):
    """Spectrogram.__new__: np.abs(stft) -> float32."""
    return _Spec(np.abs(stft_obj.data), stft=stft_obj, bin_frequencies=stft_obj.bin_frequencies)


def filtered_spectrogram(spec, filterbank=LogarithmicFilterbank, num_bands=12, fmin=30.0,
                         fmax=17000.0, fref=A4, norm_filters=True, unique_filters=True):
    """FilteredSpectrogram.__new__: np.dot(spec, filterbank) in float32."""
    if isinstance(filterbank, type) and issubclass(filterbank, Filterbank):
        filterbank = filterbank(spec.bin_frequencies, num_bands=num_bands, fmin=fmin, fmax=fmax,
                                fref=fref, norm_filters=norm_filters,
                                unique_filters=unique_filters)
    if not isinstance(filterbank, Filterbank):
        raise TypeError("not a Filterbank type or instance: %s" % filterbank)
    data = np.dot(spec.data, filterbank.data)
    return _Spec(data, stft=spec.stft, bin_frequencies=filterbank.center_frequencies,
                 filterbank=filterbank)


def logarithmic_spectrogram(spec, log=np.log10, mul=1.0, add=1.0):
    """LogarithmicSpectrogram.__new__: log(mul * spec + add), float32 throughout."""
    data = spec.data
    data = log(np.float32(mul) * data + np.float32(add)) if data.dtype == np.float32 \
        else log(mul * data + add)
    return _Spec(data, stft=spec.stft, bin_frequencies=spec.bin_frequencies,
                 filterbank=getattr(spec, "filterbank", None), mul=mul, add=add)


def diff_frames_for(diff_ratio, hop_size, frame_size, window=np.hanning):
    """madmom.audio.spectrogram._diff_frames."""
    if callable(window):
        window = window(frame_size)
    sample = np.argmax(window > float(diff_ratio) * max(window))
    diff_samples = len(window) / 2 - sample
    return int(max(1, round(diff_samples / hop_size)))


def spectrogram_difference(spec_data, diff_frames, diff_max_bins=None, positive_diffs=False):
    """SpectrogramDifference.__new__ on a plain (T, B) array."""
    if diff_frames < 1:
        raise ValueError("number of `diff_frames` must be >= 1")
    if diff_max_bins is not None and diff_max_bins > 1:
        from scipy.ndimage import maximum_filter
        diff_spec = maximum_filter(spec_data, size=(1, int(diff_max_bins)))
    else:
        diff_spec = spec_data
    diff = np.zeros_like(spec_data)
    diff[diff_frames:] = spec_data[diff_frames:] - diff_spec[:-diff_frames]
    if positive_diffs:
        np.maximum(diff, 0, out=diff)
    return diff


class SpectrogramProcessor(Processor):
    def process(self, data, **kwargs):
        if not isinstance(data, ShortTimeFourierTransform):
            data = ShortTimeFourierTransform(data, **kwargs)
        return spectrogram(data)


class FilteredSpectrogramProcessor(Processor):
    def __init__(self, filterbank=LogarithmicFilterbank, num_bands=12, fmin=30.0, fmax=17000.0,
                 fref=A4, norm_filters=True, unique_filters=True, **kwargs):
        self.filterbank = filterbank
        self.num_bands = num_bands
        self.fmin = fmin
        self.fmax = fmax
        self.fref = fref
        self.norm_filters = norm_filters
        self.unique_filters = unique_filters

    def process(self, data, **kwargs):
        if isinstance(data, ShortTimeFourierTransform):
            data = spectrogram(data)
        out = filtered_spectrogram(data, filterbank=self.filterbank, num_bands=self.num_bands,
                                   fmin=self.fmin, fmax=self.fmax, fref=self.fref,
                                   norm_filters=self.norm_filters,
                                   unique_filters=self.unique_filters)
        self.filterbank = out.filterbank      # cached like madmom does
        return out


class LogarithmicSpectrogramProcessor(Processor):
    def __init__(self, log=np.log10, mul=1.0, add=1.0, **kwargs):
        self.log = log
        self.mul = mul
        self.add = add

    def process(self, data, **kwargs):
        if isinstance(data, ShortTimeFourierTransform):
            data = spectrogram(data)
        return logarithmic_spectrogram(data, log=self.log, mul=self.mul, add=self.add)


class LogarithmicFilteredSpectrogramProcessor(Processor):
    def __init__(self, filterbank=LogarithmicFilterbank, num_bands=12, fmin=30.0, fmax=17000.0,
                 fref=A4, norm_filters=True, unique_filters=True, mul=1.0, add=1.0, **kwargs):
        self.filt = FilteredSpectrogramProcessor(filterbank, num_bands, fmin, fmax, fref,
                                                 norm_filters, unique_filters)
        self.logp = LogarithmicSpectrogramProcessor(np.log10, mul, add)

    def process(self, data, **kwargs):
        return self.logp(self.filt(data))


class BufferProcessor(Processor):
    """madmom.processors.BufferProcessor: a fixed-length buffer; new rows are shifted in at the end."""

    def __init__(self, buffer_size=None, init=None, init_value=0):
        if buffer_size is None and init is not None:
            buffer_size = init.shape
        elif isinstance(buffer_size, (int, np.integer)):
            buffer_size = (buffer_size,)
        if buffer_size is not None and init is None:
            init = np.ones(buffer_size) * init_value
        self.buffer_size = buffer_size
        self.init = init
        self.data = init

    def reset(self, init=None):
        self.data = init if init is not None else self.init

    def process(self, data, **kwargs):
        ndmin = len(self.buffer_size)
        if data.ndim < ndmin:
            data = np.array(data, subok=True, ndmin=ndmin)
        data_length = len(data)
        # remove `data_length` rows at the beginning, append the new data (a longer block does not fit: numpy raises)
        self.data = np.roll(self.data, -data_length, axis=0)
        self.data[-data_length:] = data
        return self.data


class SpectrogramDifferenceProcessor(Processor):
    """madmom.audio.spectrogram.SpectrogramDifferenceProcessor.process(data, reset=True).

    reset=True (offline): `diff_frames` rows of inf are put before the data, so rows n < diff_frames of the
    diff come out 0.  reset=False (online): the rows of the previous calls, kept by a BufferProcessor as long as
    the FIRST call's data plus diff_frames, are the history of the new rows; the result covers the whole buffer
    behind its first diff_frames rows."""

    def __init__(self, diff_ratio=0.5, diff_frames=None, diff_max_bins=None,
                 positive_diffs=False, stack_diffs=None, **kwargs):
        self.diff_ratio = diff_ratio
        self.diff_frames = diff_frames
        self.diff_max_bins = diff_max_bins
        self.positive_diffs = positive_diffs
        self.stack_diffs = stack_diffs
        self._buffer = None

    def process(self, data, reset=True, **kwargs):
        if self.diff_frames is None:
            self.diff_frames = diff_frames_for(self.diff_ratio,
                                               frame_size=data.stft.frames.frame_size,
                                               hop_size=data.stft.frames.hop_size,
                                               window=data.stft.window)
        k = self.diff_frames
        spec = data.data
        if self._buffer is None or reset:
            init = np.empty((k, spec.shape[1]), dtype=spec.dtype)
            init[:] = np.inf
            rows = np.insert(spec, 0, init, axis=0)
            self._buffer = BufferProcessor(init=rows)
        else:
            rows = self._buffer(spec)
        with np.errstate(invalid="ignore"):
            diff = spectrogram_difference(rows, k, self.diff_max_bins, self.positive_diffs)[k:]    # keep_dims=False
        diff[np.isinf(diff)] = 0
        if self.stack_diffs is None:
            return _Spec(diff, stft=data.stft, bin_frequencies=data.bin_frequencies,
                         diff_frames=k)
        return self.stack_diffs((rows[k:], diff))


# ----------------------------------------------------------------------------
# madmom/features/onsets.py (spectral_flux), madmom/audio/hpcp.py, chroma fold
# ----------------------------------------------------------------------------


def spectral_flux(diff):
    """madmom.features.onsets.spectral_flux on an already-positive diff: row sums."""
    return np.sum(diff, axis=1)


def pitch_class_profile(spec, num_classes=12, fmin=100.0, fmax=5000.0, fref=A4):
    """madmom.audio.hpcp.PitchClassProfile: np.dot(spec, PCP filterbank)."""
    fb = PitchClassProfileFilterbank(spec.bin_frequencies, num_classes, fmin, fmax, fref)
    return np.dot(spec.data, fb.data), fb


def fold_classes(center_frequencies, num_classes=12):
    """Pitch class (0 = C) of each band centre: round(hz2midi(f)) % 12."""
    midi = np.round(hz2midi(center_frequencies)).astype(int)
    return np.mod(midi, num_classes)


def fold_chroma(spec_data, center_frequencies, num_classes=12):
    """Octave fold as in madmom.audio.chroma.CLPChroma: chroma[:, pc(p)] += band[:, p]."""
    classes = fold_classes(center_frequencies, num_classes)
    chroma = np.zeros((spec_data.shape[0], num_classes), dtype=spec_data.dtype)
    for p, c in enumerate(classes):
        chroma[:, c] += spec_data[:, p]
    return chroma


def dcp_context(spec_data, context=15):
    """DeepChromaProcessor context stacking: FramedSignal(frame 15, hop 1) + flatten."""
    T, B = spec_data.shape
    half = context // 2
    out = np.zeros((T, context * B), dtype=spec_data.dtype)
    for t in range(T):
        start = t - half
        for c in range(context):
            s = start + c
            if 0 <= s < T:
                out[t, c * B:(c + 1) * B] = spec_data[s]
    return out


# ----------------------------------------------------------------------------
# feature-processor wiring (madmom/features/{beats,onsets,key,chords}.py,
# madmom/audio/chroma.py) -- only the pre-processing, never the networks
# ----------------------------------------------------------------------------


def rnn_beat_preprocessor():
    """RNNBeatProcessor() offline pre-processor (beats.py:71-75 of the reference reaches it)."""
    multi = []
    for frame_size, num_bands in zip([1024, 2048, 4096], [3, 6, 12]):
        multi.append(SequentialProcessor((
            FramedSignalProcessor(frame_size=frame_size, fps=100),
            ShortTimeFourierTransformProcessor(),
            FilteredSpectrogramProcessor(num_bands=num_bands, fmin=30, fmax=17000,
                                         norm_filters=True),
            LogarithmicSpectrogramProcessor(mul=1, add=1),
            SpectrogramDifferenceProcessor(diff_ratio=0.5, positive_diffs=True,
                                           stack_diffs=np.hstack),
        )))
    return SequentialProcessor((SignalProcessor(num_channels=1, sample_rate=44100),
                                ParallelProcessor(multi), np.hstack))


def rnn_onset_preprocessor():
    """RNNOnsetProcessor() offline pre-processor."""
    multi = []
    for frame_size in [1024, 2048, 4096]:
        multi.append(SequentialProcessor((
            FramedSignalProcessor(frame_size=frame_size, fps=100),
            ShortTimeFourierTransformProcessor(),
            FilteredSpectrogramProcessor(num_bands=6, fmin=30, fmax=17000, norm_filters=True),
            LogarithmicSpectrogramProcessor(mul=5, add=1),
            SpectrogramDifferenceProcessor(diff_ratio=0.25, positive_diffs=True,
                                           stack_diffs=np.hstack),
        )))
    return SequentialProcessor((SignalProcessor(num_channels=1, sample_rate=44100),
                                ParallelProcessor(multi), np.hstack))


def log_filt_chain(frame_size, fps=None, hop_size=441.0, num_bands=24, fmin=65.0, fmax=2100.0,
                   unique_filters=True, mul=1.0, add=1.0, sample_rate=44100):
    """DeepChroma / CNN key / CNN chord front end (extract.py:54, key.py:101)."""
    return SequentialProcessor((
        SignalProcessor(num_channels=1, sample_rate=sample_rate),
        FramedSignalProcessor(frame_size=frame_size, fps=fps, hop_size=hop_size),
        ShortTimeFourierTransformProcessor(),
        LogarithmicFilteredSpectrogramProcessor(num_bands=num_bands, fmin=fmin, fmax=fmax,
                                                unique_filters=unique_filters, mul=mul, add=add),
    ))


def log_filtered_spectrogram(x, sample_rate=44100, frame_size=2048, hop_size=441.0, fps=None,
                             num_bands=12, fmin=30.0, fmax=17000.0, norm_filters=True,
                             unique_filters=True, mul=1.0, add=1.0):
    """Config-1 shaped helper: one resolution, returns the (T, B) float32 array."""
    chain = SequentialProcessor((
        SignalProcessor(num_channels=1, sample_rate=sample_rate),
        FramedSignalProcessor(frame_size=frame_size, hop_size=hop_size, fps=fps),
        ShortTimeFourierTransformProcessor(),
        FilteredSpectrogramProcessor(num_bands=num_bands, fmin=fmin, fmax=fmax,
                                     norm_filters=norm_filters, unique_filters=unique_filters),
        LogarithmicSpectrogramProcessor(mul=mul, add=add),
    ))
    return chain(x).data
