"""Signal / FramedSignal equivalents (madmom 0.16.1 ``madmom/audio/signal.py``).

The reference wraps audio as ``Signal(arr, sample_rate=sr, num_channels=1)``
(/root/reference/backend/app/services/grid/beats.py:28-32, chords/extract.py:37-41) and madmom's
feature processors then apply ``SignalProcessor`` and ``FramedSignalProcessor``.  Here framing is
lazy: a ``FramedSignal`` only records the geometry (bit-exact with madmom, evaluated by the C
library in float64); the overlapping windows are read by the CUDA kernel straight from the signal.
"""
from __future__ import annotations

import numpy as np

from .. import _ffi
from ..processors import Processor

FRAME_SIZE, HOP_SIZE, FPS, ORIGIN, END_OF_SIGNAL, NUM_FRAMES = 2048, 441.0, None, 0, "normal", None
SAMPLE_RATE, NUM_CHANNELS = None, None

try:  # torch is only needed for device-resident signals
    import torch
except Exception:  # pragma: no cover
    torch = None


def _is_tensor(x):
    return torch is not None and isinstance(x, torch.Tensor)


def remix(signal, num_channels):
    """madmom.audio.signal.remix: down-mix by mean (cast back to the signal dtype) or up-mix by tiling."""
    if num_channels is None or num_channels == (1 if signal.ndim == 1 else signal.shape[1]):
        return signal
    if num_channels == 1 and signal.ndim > 1:
        if _is_tensor(signal):
            if signal.dtype.is_floating_point:
                return signal.mean(dim=-1)
            return (signal.to(torch.int32).sum(dim=-1).to(torch.float64) / signal.shape[-1]).to(signal.dtype)
        return np.mean(signal, axis=-1).astype(signal.dtype)
    if num_channels > 1 and signal.ndim == 1:
        if _is_tensor(signal):
            return signal[:, None].repeat(1, num_channels)
        return np.tile(np.asarray(signal)[:, np.newaxis], num_channels)
    raise NotImplementedError("Requested %d channels, but got %d channels and channel conversion is not "
                              "implemented." % (num_channels, signal.shape[1]))


class Signal(np.ndarray):
    """ndarray subclass carrying ``sample_rate`` like madmom.audio.signal.Signal (host data)."""

    def __new__(cls, data, sample_rate=None, num_channels=None, start=None, stop=None, norm=False,
                gain=0.0, dtype=None, **kwargs):
        if isinstance(data, (str, bytes)) or hasattr(data, "read"):
            data, file_rate = _load_wave(data, dtype=dtype)
            if start is not None or stop is not None:       # seconds, as madmom.io.audio.load_wave_file cuts them
                lo = 0 if start is None else int(start * file_rate)
                hi = len(data) if stop is None else min(len(data), int(stop * file_rate))
                data = data[lo:hi]
            if sample_rate is not None and sample_rate != file_rate:
                raise NotImplementedError("resampling needs ffmpeg, which the reference does before this "
                                          "path (services/audio.py:7-16)")
            sample_rate = file_rate
        src_rate = getattr(data, "sample_rate", None)
        arr = np.asarray(data)
        if dtype is not None and arr.dtype != np.dtype(dtype):
            arr = arr.astype(dtype)
        arr = remix(arr, num_channels)
        if norm:
            arr = arr.astype(np.float32) / np.max(np.abs(arr)) if np.issubdtype(arr.dtype, np.integer) \
                else arr / np.max(np.abs(arr))
        if gain is not None and gain != 0:
            scaled = np.asarray(arr, dtype=float) * np.power(np.sqrt(10.0), 0.1 * gain)
            arr = scaled.astype(arr.dtype)
        obj = np.asarray(arr).view(cls)
        obj.sample_rate = sample_rate if sample_rate is not None else src_rate
        obj.start, obj.stop = start, stop
        return obj

    def __array_finalize__(self, obj):
        if obj is None:
            return
        self.sample_rate = getattr(obj, "sample_rate", None)
        self.start = getattr(obj, "start", None)
        self.stop = getattr(obj, "stop", None)

    @property
    def num_samples(self):
        return len(self)

    @property
    def num_channels(self):
        return 1 if self.ndim == 1 else np.shape(self)[1]

    @property
    def length(self):
        return None if self.sample_rate is None else float(self.num_samples) / self.sample_rate


def _load_wave(path, dtype=None):
    """PCM WAV loader (what madmom.io.audio.load_wave_file does through scipy)."""
    from scipy.io import wavfile
    rate, data = wavfile.read(path)
    if dtype is not None:
        data = data.astype(dtype)
    return data, rate


class SignalProcessor(Processor):
    def __init__(self, sample_rate=SAMPLE_RATE, num_channels=NUM_CHANNELS, start=None, stop=None, norm=False,
                 gain=0.0, dtype=None, **kwargs):
        self.sample_rate, self.num_channels = sample_rate, num_channels
        self.start, self.stop, self.norm, self.gain, self.dtype = start, stop, norm, gain, dtype

    def process(self, data, **kwargs):
        args = dict(sample_rate=self.sample_rate, num_channels=self.num_channels, start=self.start,
                    stop=self.stop, norm=self.norm, gain=self.gain, dtype=self.dtype)
        args.update(kwargs)
        if _is_tensor(data) or isinstance(data, DeviceSignal):
            # device-resident samples are never rewritten: the options that would need a pass over them raise
            # instead of being dropped (norm is the exception: it is fused as a per-clip gain)
            if args["gain"] not in (None, 0, 0.0):
                raise NotImplementedError("gain != 0 is not implemented for device-resident signals")
            if args["start"] is not None or args["stop"] is not None:
                raise NotImplementedError("start / stop are not implemented for device-resident signals; slice the tensor")
            if args["dtype"] is not None and np.dtype(args["dtype"]) != (data.dtype if isinstance(data, DeviceSignal)
                                                                         else np.dtype(str(data.dtype).replace("torch.", ""))):
                raise NotImplementedError("dtype conversion is not implemented for device-resident signals")
        if _is_tensor(data):
            return DeviceSignal(data, sample_rate=args["sample_rate"], num_channels=args["num_channels"],
                                norm=args["norm"])
        if isinstance(data, DeviceSignal):
            return DeviceSignal(data.data, sample_rate=data.sample_rate or args["sample_rate"],
                                num_channels=args["num_channels"], norm=args["norm"] or data.norm)
        src_rate = getattr(data, "sample_rate", None)
        if src_rate is not None and args["sample_rate"] is not None and src_rate != args["sample_rate"]:
            raise NotImplementedError("resampling is done by ffmpeg upstream of this path "
                                      "(/root/reference/backend/app/services/audio.py:7-16)")
        if args["sample_rate"] is None:
            args["sample_rate"] = src_rate
        return Signal(data, **args)


class DeviceSignal:
    """A signal that already lives on the GPU (torch tensor); same attributes as ``Signal``."""

    def __init__(self, data, sample_rate=None, num_channels=None, norm=False):
        self.data = remix(data, num_channels)
        self.sample_rate = sample_rate
        # madmom Signal(norm=True) divides by max|x|.  The samples are not rewritten: the fused chains
        # apply g = 1 / max|x| (b200spec_clip_peak) as a gain on the band sums -- g for a magnitude
        # spectrogram, g^2 for a power spectrogram (the kernel squares it) -- which is the same thing
        # because the path is linear in the samples up to the magnitude.
        self.norm = bool(norm)

    def __len__(self):
        return int(self.data.shape[0])

    @property
    def dtype(self):
        return np.dtype(np.float32) if self.data.dtype == torch.float32 else np.dtype(np.int16) \
            if self.data.dtype == torch.int16 else np.dtype(str(self.data.dtype).replace("torch.", ""))

    @property
    def ndim(self):
        return self.data.ndim

    @property
    def shape(self):
        return tuple(self.data.shape)

    @property
    def num_samples(self):
        return len(self)

    @property
    def num_channels(self):
        return 1 if self.data.ndim == 1 else int(self.data.shape[1])


def resolve_origin(origin, frame_size):
    if origin in ("center", "offline"):
        origin = 0
    elif origin in ("left", "past", "online"):
        origin = (frame_size - 1) / 2
    elif origin in ("right", "future", "stream"):
        origin = -(frame_size / 2)
    return int(origin)


def signal_frame(signal, index, frame_size, hop_size, origin=0):
    """One zero-padded frame on the host (indexing convenience; the kernels never call this)."""
    num_samples = len(signal)
    start = _ffi.frame_start(index, hop_size, frame_size, int(origin))
    stop = start + frame_size
    if start >= 0 and stop <= num_samples:
        return signal[start:stop]
    frame = np.zeros((frame_size,) + tuple(np.shape(signal)[1:]), dtype=signal.dtype)
    lo, hi = max(start, 0), min(stop, num_samples)
    if hi > lo:
        frame[lo - start:hi - start] = np.asarray(signal[lo:hi])
    return frame


class FramedSignal(object):
    """Lazy framing: geometry only (madmom.audio.signal.FramedSignal)."""

    def __init__(self, signal, frame_size=FRAME_SIZE, hop_size=HOP_SIZE, fps=FPS, origin=ORIGIN,
                 end=END_OF_SIGNAL, num_frames=NUM_FRAMES, **kwargs):
        if _is_tensor(signal):
            signal = DeviceSignal(signal, **{k: v for k, v in kwargs.items() if k in ("sample_rate", "num_channels")})
        elif not isinstance(signal, (Signal, DeviceSignal)):
            signal = Signal(signal, **kwargs)
        self.signal = signal
        if frame_size:
            self.frame_size = int(frame_size)
        if hop_size:
            self.hop_size = float(hop_size)
        if fps:
            self.hop_size = self.signal.sample_rate / float(fps)
        self.origin = resolve_origin(origin, self.frame_size)
        if num_frames is None:
            num_frames = _ffi.num_frames(len(self.signal), self.hop_size, end)   # ValueError on a bad `end`
        self.num_frames = int(num_frames)
        self.end = end

    def __len__(self):
        return self.num_frames

    def __getitem__(self, index):
        if isinstance(index, (int, np.integer)):
            if index < 0:
                index += self.num_frames
            if 0 <= index < self.num_frames:
                sig = self.signal
                if isinstance(sig, DeviceSignal):
                    sig = sig.data.cpu().numpy()
                return signal_frame(sig, index, self.frame_size, self.hop_size, self.origin)
            raise IndexError("end of signal reached")
        if isinstance(index, slice):
            start, stop, step = index.indices(self.num_frames)
            if step != 1:
                raise ValueError("only slices with a step size of 1 supported")
            num = max(stop - start, 0)
            origin = self.origin - self.hop_size * start
            return FramedSignal(self.signal, frame_size=self.frame_size, hop_size=self.hop_size,
                                origin=origin, num_frames=num)
        raise TypeError("frame indices must be slices or integers")

    def __iter__(self):
        for i in range(self.num_frames):
            yield self[i]

    @property
    def frame_rate(self):
        if self.signal.sample_rate is None:
            return None
        return float(self.signal.sample_rate) / self.hop_size

    @property
    def fps(self):
        return self.frame_rate

    @property
    def overlap_factor(self):
        return 1.0 - self.hop_size / self.frame_size

    @property
    def shape(self):
        shape = (self.num_frames, self.frame_size)
        if self.signal.num_channels != 1:
            shape += (self.signal.num_channels,)
        return shape

    @property
    def ndim(self):
        return len(self.shape)


class FramedSignalProcessor(Processor):
    def __init__(self, frame_size=FRAME_SIZE, hop_size=HOP_SIZE, fps=FPS, origin=ORIGIN, end=END_OF_SIGNAL,
                 num_frames=NUM_FRAMES, **kwargs):
        self.frame_size, self.hop_size, self.fps = frame_size, hop_size, fps
        self.origin, self.end, self.num_frames = origin, end, num_frames

    def process(self, data, **kwargs):
        args = dict(frame_size=self.frame_size, hop_size=self.hop_size, fps=self.fps, origin=self.origin,
                    end=self.end, num_frames=self.num_frames)
        args.update(kwargs)
        if self.origin == "stream":
            data = data[-self.frame_size:]
        return FramedSignal(data, **args)
