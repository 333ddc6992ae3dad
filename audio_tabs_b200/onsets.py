"""librosa-style onset strength on the fused kernels.

The reference computes onset envelopes with ``librosa.onset.onset_strength``:
  /root/reference/backend/app/services/accompaniment/strum.py:114      (aggregate=np.median)
  /root/reference/backend/app/services/analysis/content_classifier.py:48,92   (default mean)
librosa 0.10.2 does: STFT (n_fft 2048, hop 512, periodic Hann, centre-padded with zeros) -> |X|^2 -> 128-band
Slaney mel filterbank -> ``power_to_db`` (10 log10(max(1e-10, S)), clipped 80 dB below the clip's maximum)
-> lag-1 positive difference -> mean / median over bands -> shifted by ``n_fft // (2 hop)`` frames.

That is the same skeleton as the madmom front end, so it runs on the same code: one ``k_front`` launch
(framing + window + FFT + power mel filterbank + dB) and one ``b200spec_onset_envelope`` launch
(per-clip maximum, top_db clip, difference, aggregate, shift).  ``onset_strength`` keeps librosa's
signature for the arguments the reference uses; unsupported options raise instead of being ignored.
"""
from __future__ import annotations

from typing import Optional, Sequence

import numpy as np
import torch

from . import _ffi
from .filters import SlaneyMelFilterbank
from .plan import FrontEnd, Packed, ResolutionSpec, _ptr, _stream_ptr

AGG_MEAN, AGG_MEDIAN = 0, 1


def hann_periodic(n: int) -> np.ndarray:
    """scipy.signal.get_window('hann', n, fftbins=True), the window librosa.stft uses."""
    return 0.5 - 0.5 * np.cos(2.0 * np.pi * np.arange(n) / n)


def mel_db_spec(sr: float, n_fft: int = 2048, hop_length: int = 512, n_mels: int = 128, fmin: float = 0.0,
                fmax: Optional[float] = None, amin: float = 1e-10) -> ResolutionSpec:
    """ResolutionSpec of ``power_to_db(melspectrogram(y, sr, n_fft, hop_length, n_mels), top_db=None)``."""
    fb = SlaneyMelFilterbank(sr, n_fft, n_mels=n_mels, fmin=fmin, fmax=fmax)
    return ResolutionSpec(frame_size=n_fft, hop_size=float(hop_length), origin=0, fft_window=hann_periodic(n_fft),
                          filterbank=fb, log=True, mul=1.0, add=0.0, power=True, log_scale=10.0, log_floor=amin)


def _aggregate_code(aggregate) -> int:
    if aggregate is None or aggregate is np.mean:
        return AGG_MEAN
    if aggregate is np.median:
        return AGG_MEDIAN
    raise ValueError("aggregate must be None, np.mean or np.median on the device")


class OnsetStrength:
    """Batch engine: many clips -> one packed (total_frames,) envelope, all on one GPU."""

    def __init__(self, sr: float = 22050, n_fft: int = 2048, hop_length: int = 512, n_mels: int = 128,
                 fmin: float = 0.0, fmax: Optional[float] = None, lag: int = 1, center: bool = True,
                 top_db: Optional[float] = 80.0, aggregate=None, device: int = 0, dtype: str = "f32",
                 channels: int = 1):
        if not center:
            raise ValueError("center=False is not implemented on the device (librosa's default is True)")
        if lag < 1:
            raise ValueError("lag must be a positive integer")
        self.spec = mel_db_spec(sr, n_fft, hop_length, n_mels, fmin, fmax)
        # librosa frames the centre-padded signal: 1 + N // hop frames == madmom's end='extend'
        self.fe = FrontEnd([self.spec], device=device, dtype=dtype, channels=channels, end="extend")
        self.lag, self.shift = int(lag), int(n_fft // (2 * hop_length))
        self.top_db = -1.0 if top_db is None else float(top_db)
        self.aggregate = _aggregate_code(aggregate)
        self.n_mels = int(n_mels)

    def mel_db(self, packed: Packed) -> torch.Tensor:
        """(total_frames, n_mels) dB mel spectrogram before the top_db clip."""
        return self.fe.run_packed(packed)

    def envelope(self, packed: Packed, mel_db: Optional[torch.Tensor] = None) -> torch.Tensor:
        L = self.mel_db(packed) if mel_db is None else mel_db
        dev = self.fe.device
        env = torch.empty(packed.total_frames, dtype=torch.float32, device=dev)
        scratch = torch.empty(max(packed.n_clips, 1), dtype=torch.float32, device=dev)
        _ffi.check(self.fe._lib.b200spec_onset_envelope(
            _ptr(L), L.shape[1], self.n_mels, _ptr(packed.frame_off), packed.n_clips, packed.total_frames,
            self.lag, self.top_db, self.aggregate, self.shift, _ptr(scratch), _ptr(env), _stream_ptr(None, dev)))
        return env

    def process_batch(self, signals: Sequence, return_tensors: bool = False):
        """signals: host arrays (or CUDA tensors) -> one (frames_i,) envelope per clip."""
        packed = self.fe.pack(signals)
        env = self.envelope(packed)
        off = packed.frame_off_host
        if return_tensors:
            return [env[off[i]:off[i + 1]] for i in range(packed.n_clips)]
        host = env.cpu().numpy()
        return [host[off[i]:off[i + 1]] for i in range(packed.n_clips)]


_ENGINES = {}


def onset_strength(*, y=None, sr=22050, S=None, lag=1, max_size=1, ref=None, detrend=False, center=True,
                   feature=None, aggregate=None, n_fft=2048, hop_length=512, n_mels=128, fmin=0.0, fmax=None,
                   **kwargs):
    """``librosa.onset.onset_strength(y=..., sr=...)`` for the argument combinations the reference uses."""
    if y is None or S is not None:
        raise ValueError("pass the time series y; a pre-computed S is not supported on the device")
    if max_size != 1 or ref is not None or detrend or feature is not None or kwargs:
        raise ValueError("only max_size=1, ref=None, detrend=False and the default mel feature run on the device")
    dtype = "f32"                                   # librosa only ever sees floating-point audio
    if isinstance(y, torch.Tensor):
        y = y.to(torch.float32)
    else:
        y = np.ascontiguousarray(y, dtype=np.float32)
    if y.ndim != 1:
        raise ValueError("y must be mono (librosa averages channels before this call)")
    key = (float(sr), int(n_fft), int(hop_length), int(n_mels), float(fmin), fmax, int(lag), bool(center),
           _aggregate_code(aggregate), dtype, torch.cuda.current_device())
    eng = _ENGINES.get(key)
    if eng is None:
        eng = _ENGINES[key] = OnsetStrength(sr, n_fft, hop_length, n_mels, fmin, fmax, lag, center, 80.0, aggregate,
                                            device=torch.cuda.current_device(), dtype=dtype)
    return eng.process_batch([y])[0]
