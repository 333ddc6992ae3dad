"""A behavioural stand-in for the parts of madmom 0.16.1 that surround the hot path -- TEST INFRASTRUCTURE.

madmom itself (pinned at /root/reference/backend/requirements.txt:15) cannot be installed here, so
``install()`` is exercised against this package instead.  Unlike a bag of empty classes it reproduces the
behaviours of the real package that decide whether a swapped-in processor survives inside madmom's own
feature processors:

* every stage is an ``np.ndarray`` SUBCLASS created in ``__new__`` (Signal, ShortTimeFourierTransform,
  Spectrogram, FilteredSpectrogram, LogarithmicSpectrogram, SpectrogramDifference);
* ``Signal.__new__`` treats anything that is not an ``np.ndarray`` as a FILE NAME and tries to load it
  (madmom/audio/signal.py) -- the hazard for lazy, non-ndarray stages;
* ``processors._process`` forwards ``**kwargs`` only to instances of madmom's OWN ``Processor`` class and
  ``SequentialProcessor`` wraps nested lists / tuples instead of flattening them (madmom/processors.py);
* the feature processors import the audio processors lazily inside ``__init__`` and have the shapes of
  ``RNNBeatProcessor`` (ParallelProcessor over three resolutions + ``np.hstack``),
  ``DeepChromaProcessor`` (``SignalProcessor(sample_rate=10)`` re-wrap, ``FramedSignalProcessor(frame_size=15,
  hop_size=1, fps=10)``, ``_dcp_flatten``) and ``CNNKeyRecognitionProcessor`` (file name in, network layers that
  use ``data.ndim`` / ``data.shape`` / arithmetic / ``reshape`` on the spectrogram object);
* the networks are small deterministic stand-ins (the real weights are not available), loaded through a
  ``NeuralNetworkEnsemble.load`` with madmom's ``process(data)`` -> layers protocol.

The numerical kernels of the stock path are the oracle's (oracle/madmom_ref.py), so a run through the
UNPATCHED stub is the expected value of the same run after ``install()``.

Reference call sites modelled: /root/reference/backend/app/services/grid/beats.py:28-32,71-75,
chords/extract.py:37-41,54-57, theory/key.py:99-101,143-144.
"""
from __future__ import annotations

import sys
import types
from collections.abc import MutableSequence

import numpy as np

from oracle import madmom_ref as ref


def build():
    """Return {module name: module} of a fresh stub package (insert into sys.modules to 'install madmom')."""
    mods = {n: types.ModuleType(n) for n in (
        "madmom", "madmom.processors", "madmom.io", "madmom.io.audio", "madmom.audio", "madmom.audio.signal",
        "madmom.audio.stft", "madmom.audio.filters", "madmom.audio.spectrogram", "madmom.audio.chroma",
        "madmom.ml", "madmom.ml.nn", "madmom.models", "madmom.features", "madmom.features.beats",
        "madmom.features.key", "madmom.features.onsets")}
    mods["madmom"].__version__ = "0.16.1-stub"
    for name, mod in mods.items():
        if "." in name:
            parent, child = name.rsplit(".", 1)
            setattr(mods[parent], child, mod)

    # ---- madmom/processors.py ------------------------------------------------------------------
    P = mods["madmom.processors"]

    class Processor(object):
        def process(self, data, **kwargs):
            raise NotImplementedError("Must be implemented by subclass.")

        def __call__(self, *args, **kwargs):
            return self.process(*args, **kwargs)

    def _process(process_tuple):
        if isinstance(process_tuple[0], Processor):              # madmom's OWN Processor class
            return process_tuple[0](*process_tuple[1:-1], **process_tuple[-1])
        return process_tuple[0](*process_tuple[1:-1])

    class SequentialProcessor(MutableSequence, Processor):
        def __init__(self, processors):
            self.processors = []
            for processor in processors:
                if isinstance(processor, (list, tuple)):         # 0.16.1 wraps lists / tuples, it does not flatten
                    processor = SequentialProcessor(processor)
                self.processors.append(processor)

        def __getitem__(self, index):
            return self.processors[index]

        def __setitem__(self, index, processor):
            self.processors[index] = processor

        def __delitem__(self, index):
            del self.processors[index]

        def __len__(self):
            return len(self.processors)

        def insert(self, index, processor):
            self.processors.insert(index, processor)

        def process(self, data, **kwargs):
            for processor in self.processors:
                data = _process((processor, data, kwargs))
            return data

    class ParallelProcessor(SequentialProcessor):
        def __init__(self, processors, num_threads=None):
            super(ParallelProcessor, self).__init__(processors)
            self.map = map                                       # num_threads None -> 1 -> builtin map

        def process(self, data, **kwargs):
            import itertools as it
            return list(self.map(_process, zip(self.processors, it.repeat(data), it.repeat(kwargs))))

    P.Processor, P.SequentialProcessor, P.ParallelProcessor, P._process = Processor, SequentialProcessor, ParallelProcessor, _process

    # ---- madmom/io/audio.py --------------------------------------------------------------------
    IO = mods["madmom.io.audio"]

    class LoadAudioFileError(Exception):
        pass

    def load_audio_file(filename, sample_rate=None, num_channels=None, start=None, stop=None, dtype=None):
        if not isinstance(filename, (str, bytes)) and not hasattr(filename, "read"):
            raise LoadAudioFileError("cannot load %r as an audio file (madmom treats every non-ndarray as a file "
                                     "name)" % type(filename))
        from scipy.io import wavfile
        file_rate, signal = wavfile.read(filename, mmap=True)
        if dtype is not None:
            signal = signal.astype(dtype)
        if num_channels == 1 and signal.ndim > 1:
            signal = np.mean(signal, axis=-1).astype(signal.dtype)
        if sample_rate is not None and sample_rate != file_rate:
            raise LoadAudioFileError("resampling needs ffmpeg")
        return np.asarray(signal), file_rate

    IO.LoadAudioFileError, IO.load_audio_file = LoadAudioFileError, load_audio_file

    # ---- madmom/audio/signal.py ----------------------------------------------------------------
    S = mods["madmom.audio.signal"]

    class Signal(np.ndarray):
        def __new__(cls, data, sample_rate=None, num_channels=None, start=None, stop=None, norm=False, gain=0.,
                    dtype=None, **kwargs):
            if not isinstance(data, np.ndarray):                 # "try to load an audio file if the data is not a numpy array"
                data, sample_rate = load_audio_file(data, sample_rate=sample_rate, num_channels=num_channels,
                                                    start=start, stop=stop, dtype=dtype)
            if not isinstance(data, Signal):
                data = np.asarray(data).view(cls)
                data.sample_rate = sample_rate
            if num_channels:
                data = ref.remix(np.asarray(data), num_channels).view(cls)
                data.sample_rate = sample_rate
            if norm:
                data = (np.asarray(data).astype(np.float32) / np.max(np.abs(data))).view(cls)
                data.sample_rate = sample_rate
            return data

        def __array_finalize__(self, obj):
            if obj is None:
                return
            self.sample_rate = getattr(obj, "sample_rate", None)

        @property
        def num_samples(self):
            return len(self)

        @property
        def num_channels(self):
            return 1 if self.ndim == 1 else np.shape(self)[1]

    class SignalProcessor(Processor):
        def __init__(self, sample_rate=None, num_channels=None, start=None, stop=None, norm=False, gain=0.,
                     dtype=None, **kwargs):
            self.sample_rate, self.num_channels, self.norm, self.dtype = sample_rate, num_channels, norm, dtype

        def process(self, data, **kwargs):
            args = dict(sample_rate=self.sample_rate, num_channels=self.num_channels, norm=self.norm, dtype=self.dtype)
            args.update(kwargs)
            return Signal(data, **args)

    class FramedSignal(object):
        def __init__(self, signal, frame_size=2048, hop_size=441., fps=None, origin=0, end="normal", num_frames=None,
                     **kwargs):
            if not isinstance(signal, Signal):
                signal = Signal(signal, **kwargs)
            self.signal = signal
            self.frame_size = int(frame_size)
            self.hop_size = float(hop_size)
            if fps:
                self.hop_size = self.signal.sample_rate / float(fps)
            self.origin = ref.resolve_origin(origin, self.frame_size)
            if num_frames is None:
                num_frames = ref.num_frames_for(len(self.signal), self.hop_size, end)
            self.num_frames = int(num_frames)

        def __getitem__(self, index):
            if isinstance(index, (int, np.integer)):
                if index < 0:
                    index += self.num_frames
                if 0 <= index < self.num_frames:
                    return ref.signal_frame(self.signal, index, self.frame_size, self.hop_size, self.origin)
                raise IndexError("end of signal reached")
            raise TypeError("frame indices must be slices or integers")

        def __len__(self):
            return self.num_frames

        @property
        def shape(self):
            shape = self.num_frames, self.frame_size
            if self.signal.num_channels != 1:
                shape += (self.signal.num_channels,)
            return shape

        @property
        def ndim(self):
            return len(self.shape)

    class FramedSignalProcessor(Processor):
        def __init__(self, frame_size=2048, hop_size=441., fps=None, origin=0, end="normal", num_frames=None, **kwargs):
            self.frame_size, self.hop_size, self.fps = frame_size, hop_size, fps
            self.origin, self.end, self.num_frames = origin, end, num_frames

        def process(self, data, **kwargs):
            args = dict(frame_size=self.frame_size, hop_size=self.hop_size, fps=self.fps, origin=self.origin,
                        end=self.end, num_frames=self.num_frames)
            args.update(kwargs)
            return FramedSignal(data, **args)

    S.Signal, S.SignalProcessor, S.FramedSignal, S.FramedSignalProcessor = Signal, SignalProcessor, FramedSignal, FramedSignalProcessor

    # ---- madmom/audio/filters.py (numerics from the oracle) --------------------------------------
    F = mods["madmom.audio.filters"]

    class Filterbank(np.ndarray):
        def __new__(cls, data, bin_frequencies):
            obj = np.asarray(data, dtype=np.float32).view(cls)
            obj.bin_frequencies = np.asarray(bin_frequencies, dtype=float)
            return obj

        def __array_finalize__(self, obj):
            if obj is None:
                return
            self.bin_frequencies = getattr(obj, "bin_frequencies", None)

        @property
        def center_frequencies(self):
            return ref.Filterbank(np.asarray(self), self.bin_frequencies).center_frequencies

    class LogarithmicFilterbank(Filterbank):
        def __new__(cls, bin_frequencies, num_bands=12, fmin=30., fmax=17000., fref=440., norm_filters=True,
                    unique_filters=True, bands_per_octave=True):
            fb = ref.LogarithmicFilterbank(bin_frequencies, num_bands=num_bands, fmin=fmin, fmax=fmax, fref=fref,
                                           norm_filters=norm_filters, unique_filters=unique_filters)
            return Filterbank.__new__(cls, fb.data, bin_frequencies)

    F.Filterbank, F.LogarithmicFilterbank = Filterbank, LogarithmicFilterbank

    # ---- madmom/audio/stft.py ------------------------------------------------------------------
    T = mods["madmom.audio.stft"]

    class ShortTimeFourierTransform(np.ndarray):
        def __new__(cls, frames, window=np.hanning, fft_size=None, circular_shift=False, include_nyquist=False,
                    fft_window=None, fftw=None, **kwargs):
            if isinstance(frames, ShortTimeFourierTransform):
                frames = frames.frames
            if not isinstance(frames, FramedSignal):
                frames = FramedSignal(frames, **kwargs)
            frame_size = frames.shape[1]
            if fft_window is None:
                window, fft_window = ref.fft_window_for(window, frame_size, frames.signal.dtype)
            data = ref.stft(frames, fft_window, fft_size=fft_size, circular_shift=circular_shift,
                            include_nyquist=include_nyquist)
            obj = np.asarray(data).view(cls)
            obj.frames, obj.window, obj.fft_window = frames, window, fft_window
            obj.bin_frequencies = ref.fft_frequencies(obj.shape[1], frames.signal.sample_rate)
            return obj

        def __array_finalize__(self, obj):
            if obj is None:
                return
            for k in ("frames", "window", "fft_window", "bin_frequencies"):
                setattr(self, k, getattr(obj, k, None))

    class ShortTimeFourierTransformProcessor(Processor):
        def __init__(self, window=np.hanning, fft_size=None, circular_shift=False, include_nyquist=False, **kwargs):
            self.window, self.fft_size, self.circular_shift, self.include_nyquist = window, fft_size, circular_shift, include_nyquist
            self.fft_window = None

        def process(self, data, **kwargs):
            data = ShortTimeFourierTransform(data, window=self.window, fft_size=self.fft_size,
                                             circular_shift=self.circular_shift, include_nyquist=self.include_nyquist,
                                             fft_window=self.fft_window, **kwargs)
            self.fft_window = data.fft_window
            return data

    T.ShortTimeFourierTransform, T.ShortTimeFourierTransformProcessor = ShortTimeFourierTransform, ShortTimeFourierTransformProcessor
    T.STFT, T.STFTProcessor = ShortTimeFourierTransform, ShortTimeFourierTransformProcessor
    T.fft_frequencies = ref.fft_frequencies

    # ---- madmom/audio/spectrogram.py -------------------------------------------------------------
    SP = mods["madmom.audio.spectrogram"]

    class Spectrogram(np.ndarray):
        _attrs = ("stft", "bin_frequencies", "filterbank", "mul", "add", "diff_frames")

        def __new__(cls, stft, **kwargs):
            if isinstance(stft, Spectrogram):
                return stft
            if not isinstance(stft, ShortTimeFourierTransform):
                stft = ShortTimeFourierTransform(stft, **kwargs)
            obj = np.abs(np.asarray(stft)).view(cls)
            obj.stft, obj.bin_frequencies = stft, stft.bin_frequencies
            return obj

        def __array_finalize__(self, obj):
            if obj is None:
                return
            for k in self._attrs:
                setattr(self, k, getattr(obj, k, None))

        @property
        def frames(self):
            return self.stft.frames

    def _carry(data, cls, src, **extra):
        obj = np.asarray(data).view(cls)
        obj.stft, obj.bin_frequencies = src.stft, src.bin_frequencies
        obj.filterbank = getattr(src, "filterbank", None)
        for k, v in extra.items():
            setattr(obj, k, v)
        return obj

    class FilteredSpectrogram(Spectrogram):
        def __new__(cls, spectrogram, filterbank=LogarithmicFilterbank, num_bands=12, fmin=30., fmax=17000.,
                    fref=440., norm_filters=True, unique_filters=True, **kwargs):
            import inspect
            if not isinstance(spectrogram, Spectrogram):
                spectrogram = Spectrogram(spectrogram, **kwargs)
            if inspect.isclass(filterbank) and issubclass(filterbank, Filterbank):
                filterbank = filterbank(spectrogram.bin_frequencies, num_bands=num_bands, fmin=fmin, fmax=fmax,
                                        fref=fref, norm_filters=norm_filters, unique_filters=unique_filters)
            if not isinstance(filterbank, Filterbank):
                raise TypeError("not a Filterbank type or instance: %s" % filterbank)
            data = np.dot(np.asarray(spectrogram), np.asarray(filterbank))
            return _carry(data, cls, spectrogram, filterbank=filterbank, bin_frequencies=filterbank.center_frequencies)

    class LogarithmicSpectrogram(Spectrogram):
        def __new__(cls, spectrogram, log=np.log10, mul=1., add=1., **kwargs):
            if not isinstance(spectrogram, Spectrogram):
                spectrogram = Spectrogram(spectrogram, **kwargs)
            data = ref.logarithmic_spectrogram(ref._Spec(np.asarray(spectrogram)), log=log, mul=mul, add=add).data
            return _carry(data, cls, spectrogram, mul=mul, add=add)

    class LogarithmicFilteredSpectrogram(LogarithmicSpectrogram):
        def __new__(cls, spectrogram, filterbank=LogarithmicFilterbank, num_bands=12, fmin=30., fmax=17000.,
                    fref=440., norm_filters=True, unique_filters=True, mul=1., add=1., **kwargs):
            if not isinstance(spectrogram, FilteredSpectrogram):
                spectrogram = FilteredSpectrogram(spectrogram, filterbank=filterbank, num_bands=num_bands, fmin=fmin,
                                                  fmax=fmax, fref=fref, norm_filters=norm_filters,
                                                  unique_filters=unique_filters, **kwargs)
            return LogarithmicSpectrogram.__new__(cls, spectrogram, mul=mul, add=add)

    class SpectrogramDifference(Spectrogram):
        def __new__(cls, spectrogram, diff_ratio=0.5, diff_frames=None, diff_max_bins=None, positive_diffs=False,
                    keep_dims=True, **kwargs):
            if not isinstance(spectrogram, Spectrogram):
                spectrogram = Spectrogram(spectrogram, **kwargs)
            if diff_frames is None:
                diff_frames = ref.diff_frames_for(diff_ratio, frame_size=spectrogram.stft.frames.frame_size,
                                                  hop_size=spectrogram.stft.frames.hop_size,
                                                  window=spectrogram.stft.window)
            with np.errstate(invalid="ignore"):
                data = ref.spectrogram_difference(np.asarray(spectrogram), diff_frames, diff_max_bins, positive_diffs)
            return _carry(data, cls, spectrogram, diff_frames=diff_frames)

    class _SimpleProc(Processor):
        cls = None

        def __init__(self, **kwargs):
            self.kwargs = kwargs

        def process(self, data, **kwargs):
            args = dict(self.kwargs)
            args.update(kwargs)
            return self.cls(data, **args)

    class SpectrogramProcessor(_SimpleProc):
        cls = Spectrogram

    class FilteredSpectrogramProcessor(_SimpleProc):
        cls = FilteredSpectrogram

    class LogarithmicSpectrogramProcessor(_SimpleProc):
        cls = LogarithmicSpectrogram

    class LogarithmicFilteredSpectrogramProcessor(_SimpleProc):
        cls = LogarithmicFilteredSpectrogram

    class SpectrogramDifferenceProcessor(Processor):
        def __init__(self, diff_ratio=0.5, diff_frames=None, diff_max_bins=None, positive_diffs=False,
                     stack_diffs=None, **kwargs):
            self.diff_ratio, self.diff_frames, self.diff_max_bins = diff_ratio, diff_frames, diff_max_bins
            self.positive_diffs, self.stack_diffs = positive_diffs, stack_diffs

        def process(self, data, reset=True, **kwargs):
            if self.diff_frames is None:
                self.diff_frames = ref.diff_frames_for(self.diff_ratio, frame_size=data.stft.frames.frame_size,
                                                       hop_size=data.stft.frames.hop_size, window=data.stft.window)
            k = self.diff_frames
            init = np.empty((k, data.shape[1]))
            init[:] = np.nan
            padded = np.insert(np.asarray(data), 0, init, axis=0)          # NaN rows in front (reset=True)
            with np.errstate(invalid="ignore"):
                diff = ref.spectrogram_difference(padded, k, self.diff_max_bins, self.positive_diffs)
            diff[np.isnan(diff)] = 0
            diff = diff[k:]
            if self.stack_diffs is None:
                return _carry(diff, SpectrogramDifference, data, diff_frames=k)
            return self.stack_diffs((data, diff))

    for c in (Spectrogram, FilteredSpectrogram, LogarithmicSpectrogram, LogarithmicFilteredSpectrogram,
              SpectrogramDifference, SpectrogramProcessor, FilteredSpectrogramProcessor,
              LogarithmicSpectrogramProcessor, LogarithmicFilteredSpectrogramProcessor,
              SpectrogramDifferenceProcessor):
        setattr(SP, c.__name__, c)
    SP._diff_frames = ref.diff_frames_for

    # ---- madmom/ml/nn: deterministic stand-ins for the pickled networks ---------------------------
    NN = mods["madmom.ml.nn"]

    class FeedForwardLayer(object):
        def __init__(self, n_in, n_out, seed):
            rng = np.random.default_rng(seed)
            self.weights = (rng.standard_normal((n_in, n_out)) / np.sqrt(n_in)).astype(np.float32)
            self.bias = rng.standard_normal(n_out).astype(np.float32) * 0.1

        def activate(self, data, **kwargs):
            return 1.0 / (1.0 + np.exp(-(np.dot(data, self.weights) + self.bias)))

    class BatchNormLayer(object):
        def __init__(self, mean, inv_std):
            self.mean, self.inv_std = np.float32(mean), np.float32(inv_std)

        def activate(self, data, **kwargs):
            return (data - self.mean) * self.inv_std             # arithmetic directly on the spectrogram object

    class PoolFramesLayer(object):
        """Stands in for the convolution stack of the key CNN: uses data.shape / reshape like ConvolutionalLayer."""

        def activate(self, data, **kwargs):
            if len(data.shape) == 2:
                data = data.reshape(data.shape + (1,))
            return np.asarray(data).mean(axis=0)[:, 0][np.newaxis, :]

    class NeuralNetwork(Processor):
        def __init__(self, layers):
            self.layers = layers

        def process(self, data, reset=True, **kwargs):
            if data.ndim < 2:
                data = np.array(data, subok=True, copy=False, ndmin=2) if np.lib.NumpyVersion(np.__version__) < "2.0.0" \
                    else np.asarray(data).reshape(1, -1)
            for layer in self.layers:
                data = layer.activate(data)
            if data.ndim == 2 and data.shape[1] == 1:
                data = data.ravel()
            return data

    class NeuralNetworkEnsemble(SequentialProcessor):
        def __init__(self, networks, ensemble_fn=None, **kwargs):
            super(NeuralNetworkEnsemble, self).__init__((ParallelProcessor(networks), ensemble_fn or
                                                         (lambda preds: sum(preds) / len(preds))))

        @classmethod
        def load(cls, nn_files, **kwargs):
            return cls([NeuralNetwork(layers) for layers in nn_files], **kwargs)

    NN.NeuralNetwork, NN.NeuralNetworkEnsemble = NeuralNetwork, NeuralNetworkEnsemble
    M = mods["madmom.models"]
    M.BEATS_BLSTM = [[FeedForwardLayer(314, 1, 10 + i)] for i in range(2)]
    M.ONSETS_BRNN = [[FeedForwardLayer(266, 1, 40)]]
    M.CHROMA_DNN = [[FeedForwardLayer(1575, 12, 20)]]
    M.KEY_CNN = [[BatchNormLayer(0.3, 2.0), PoolFramesLayer(), FeedForwardLayer(105, 24, 30)]]

    # ---- feature processors: lazy imports inside __init__, as in madmom ---------------------------
    def _dcp_flatten(fs):
        return np.concatenate(fs).reshape(len(fs), -1)

    class DeepChromaProcessor(SequentialProcessor):
        def __init__(self, fmin=65, fmax=2100, unique_filters=True, models=None, **kwargs):
            from madmom.models import CHROMA_DNN
            from madmom.audio.signal import SignalProcessor, FramedSignalProcessor
            from madmom.audio.stft import ShortTimeFourierTransformProcessor
            from madmom.audio.spectrogram import LogarithmicFilteredSpectrogramProcessor
            from madmom.ml.nn import NeuralNetworkEnsemble
            sig = SignalProcessor(num_channels=1, sample_rate=44100)
            frames = FramedSignalProcessor(frame_size=8192, fps=10)
            stft = ShortTimeFourierTransformProcessor()
            spec = LogarithmicFilteredSpectrogramProcessor(num_bands=24, fmin=fmin, fmax=fmax,
                                                           unique_filters=unique_filters)
            spec_signal = SignalProcessor(sample_rate=10)
            spec_frames = FramedSignalProcessor(frame_size=15, hop_size=1, fps=10)
            nn = NeuralNetworkEnsemble.load(models or CHROMA_DNN, **kwargs)
            super(DeepChromaProcessor, self).__init__([sig, frames, stft, spec, spec_signal, spec_frames,
                                                       _dcp_flatten, nn])

    mods["madmom.audio.chroma"].DeepChromaProcessor = DeepChromaProcessor
    mods["madmom.audio.chroma"]._dcp_flatten = _dcp_flatten

    class _RNNFrontEndProcessor(SequentialProcessor):
        NUM_BANDS, MUL, DIFF_RATIO, MODELS = [3, 6, 12], 1, 0.5, "BEATS_BLSTM"

        def __init__(self, post_processor=None, online=False, nn_files=None, **kwargs):
            from madmom.audio.signal import SignalProcessor, FramedSignalProcessor
            from madmom.audio.stft import ShortTimeFourierTransformProcessor
            from madmom.audio.spectrogram import (FilteredSpectrogramProcessor, LogarithmicSpectrogramProcessor,
                                                  SpectrogramDifferenceProcessor)
            from madmom.ml.nn import NeuralNetworkEnsemble
            import madmom.models
            sig = SignalProcessor(num_channels=1, sample_rate=44100)
            multi = ParallelProcessor([])
            for frame_size, num_bands in zip([1024, 2048, 4096], self.NUM_BANDS):
                frames = FramedSignalProcessor(frame_size=frame_size, **kwargs)
                stft = ShortTimeFourierTransformProcessor()
                filt = FilteredSpectrogramProcessor(num_bands=num_bands, fmin=30, fmax=17000, norm_filters=True)
                spec = LogarithmicSpectrogramProcessor(mul=self.MUL, add=1)
                diff = SpectrogramDifferenceProcessor(diff_ratio=self.DIFF_RATIO, positive_diffs=True,
                                                      stack_diffs=np.hstack)
                multi.append(SequentialProcessor((frames, stft, filt, spec, diff)))
            pre_processor = SequentialProcessor((sig, multi, np.hstack))
            nn = NeuralNetworkEnsemble.load(nn_files or getattr(madmom.models, self.MODELS), **kwargs)
            super(_RNNFrontEndProcessor, self).__init__((pre_processor, nn))

    class RNNBeatProcessor(_RNNFrontEndProcessor):
        pass

    class RNNOnsetProcessor(_RNNFrontEndProcessor):
        NUM_BANDS, MUL, DIFF_RATIO, MODELS = [6, 6, 6], 5, 0.25, "ONSETS_BRNN"

    mods["madmom.features.beats"].RNNBeatProcessor = RNNBeatProcessor
    mods["madmom.features.onsets"].RNNOnsetProcessor = RNNOnsetProcessor

    class CNNKeyRecognitionProcessor(SequentialProcessor):
        def __init__(self, nn_files=None, **kwargs):
            from madmom.audio.signal import SignalProcessor, FramedSignalProcessor
            from madmom.audio.stft import ShortTimeFourierTransformProcessor
            from madmom.audio.spectrogram import LogarithmicFilteredSpectrogramProcessor
            from madmom.ml.nn import NeuralNetworkEnsemble
            from madmom.models import KEY_CNN
            sig = SignalProcessor(num_channels=1, sample_rate=44100)
            frames = FramedSignalProcessor(frame_size=8192, fps=5)
            stft = ShortTimeFourierTransformProcessor()
            spec = LogarithmicFilteredSpectrogramProcessor(num_bands=24, fmin=65, fmax=2100, unique_filters=True)
            nn = NeuralNetworkEnsemble.load(nn_files or KEY_CNN)
            super(CNNKeyRecognitionProcessor, self).__init__([sig, frames, stft, spec, nn])

    mods["madmom.features.key"].CNNKeyRecognitionProcessor = CNNKeyRecognitionProcessor
    return mods


def activate(monkeypatch):
    """Put a fresh stub into sys.modules for the duration of a test."""
    mods = build()
    for k in [k for k in sys.modules if k == "madmom" or k.startswith("madmom.")]:
        monkeypatch.delitem(sys.modules, k)
    for k, v in mods.items():
        monkeypatch.setitem(sys.modules, k, v)
    return mods
