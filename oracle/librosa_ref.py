"""CPU oracle, part 2: a numpy restatement of librosa 0.10.2's onset-strength front end.

TEST INFRASTRUCTURE ONLY -- imported by tests/ (and nothing else).  The product path is the CUDA
library; it never imports this module.

What it restates and why: the reference calls ``librosa.onset.onset_strength(y=y, sr=sr)``
(/root/reference/backend/app/services/analysis/content_classifier.py:48,92, mean over bands) and
``librosa.onset.onset_strength(y=y, sr=int(sr), aggregate=np.median)``
(/root/reference/backend/app/services/accompaniment/strum.py:114).  librosa is pinned at
``librosa==0.10.2.post1`` (/root/reference/backend/requirements.txt:14), is NOT vendored under
/root/reference and is not installed here (no network), so -- exactly as for madmom -- this is a
restatement of its published algorithm and **parity is unpinned**: no golden vector of the reference
covers this path.  (Partial external pin, round 2: the mel filterbank, the power mel spectrogram in dB and the
top_db clip agree with ``transformers.audio_utils``, a third-party numpy port of the same librosa functions
that is installed in this image -- tests/test_oracle_crosscheck.py; ``onset_strength`` itself has no such
counterpart.)  Functions cite the librosa function they follow:

  librosa.core.convert.{hz_to_mel, mel_to_hz, mel_frequencies, fft_frequencies}
  librosa.filters.mel                       (Slaney scale, norm='slaney', float32)
  librosa.core.spectrum.stft                (center=True, pad_mode='constant', periodic Hann, complex64)
  librosa.feature.melspectrogram            (power=2.0, n_mels=128)
  librosa.core.spectrum.power_to_db         (ref=1.0, amin=1e-10, top_db=80.0)
  librosa.onset.onset_strength[_multi]      (lag=1, max_size=1, center=True, detrend=False)

Arrays here are (frames, bands) -- the transpose of librosa's (bands, frames) -- because that is the
layout of the rest of this repository; the arithmetic is the same.
"""
import numpy as np
from scipy import fftpack


def hz_to_mel(frequencies, htk=False):
    """librosa.hz_to_mel (Slaney by default)."""
    frequencies = np.asanyarray(frequencies, dtype=float)
    if htk:
        return 2595.0 * np.log10(1.0 + frequencies / 700.0)
    f_min, f_sp = 0.0, 200.0 / 3
    mels = (frequencies - f_min) / f_sp
    min_log_hz = 1000.0
    min_log_mel = (min_log_hz - f_min) / f_sp
    logstep = np.log(6.4) / 27.0
    if frequencies.ndim:
        log_t = frequencies >= min_log_hz
        mels[log_t] = min_log_mel + np.log(frequencies[log_t] / min_log_hz) / logstep
    elif frequencies >= min_log_hz:
        mels = min_log_mel + np.log(frequencies / min_log_hz) / logstep
    return mels


def mel_to_hz(mels, htk=False):
    """librosa.mel_to_hz."""
    mels = np.asanyarray(mels, dtype=float)
    if htk:
        return 700.0 * (10.0 ** (mels / 2595.0) - 1.0)
    f_min, f_sp = 0.0, 200.0 / 3
    freqs = f_min + f_sp * mels
    min_log_hz = 1000.0
    min_log_mel = (min_log_hz - f_min) / f_sp
    logstep = np.log(6.4) / 27.0
    if mels.ndim:
        log_t = mels >= min_log_mel
        freqs[log_t] = min_log_hz * np.exp(logstep * (mels[log_t] - min_log_mel))
    elif mels >= min_log_mel:
        freqs = min_log_hz * np.exp(logstep * (mels - min_log_mel))
    return freqs


def mel_frequencies(n_mels=128, fmin=0.0, fmax=11025.0, htk=False):
    """librosa.mel_frequencies: n_mels points uniformly spaced on the mel axis."""
    mels = np.linspace(hz_to_mel(fmin, htk=htk), hz_to_mel(fmax, htk=htk), n_mels)
    return mel_to_hz(mels, htk=htk)


def fft_frequencies(sr=22050, n_fft=2048):
    """librosa.fft_frequencies = np.fft.rfftfreq (includes the Nyquist bin)."""
    return np.fft.rfftfreq(n=n_fft, d=1.0 / sr)


def mel(sr, n_fft, n_mels=128, fmin=0.0, fmax=None, htk=False, norm="slaney", dtype=np.float32):
    """librosa.filters.mel -> (n_mels, 1 + n_fft // 2) float32."""
    if fmax is None:
        fmax = float(sr) / 2
    n_mels = int(n_mels)
    weights = np.zeros((n_mels, int(1 + n_fft // 2)), dtype=dtype)
    fftfreqs = fft_frequencies(sr=sr, n_fft=n_fft)
    mel_f = mel_frequencies(n_mels + 2, fmin=fmin, fmax=fmax, htk=htk)
    fdiff = np.diff(mel_f)
    ramps = np.subtract.outer(mel_f, fftfreqs)
    for i in range(n_mels):
        lower = -ramps[i] / fdiff[i]
        upper = ramps[i + 2] / fdiff[i + 1]
        weights[i] = np.maximum(0, np.minimum(lower, upper))
    if norm == "slaney":
        enorm = 2.0 / (mel_f[2:n_mels + 2] - mel_f[:n_mels])
        weights *= enorm[:, np.newaxis]
    elif norm is not None:
        raise NotImplementedError("only norm='slaney' / None are restated")
    return weights


def hann_periodic(n):
    """scipy.signal.get_window('hann', n, fftbins=True)."""
    return 0.5 - 0.5 * np.cos(2.0 * np.pi * np.arange(n) / n)


def stft(y, n_fft=2048, hop_length=512, center=True):
    """librosa.stft(window='hann', pad_mode='constant') -> (frames, 1 + n_fft // 2) complex64."""
    y = np.asarray(y)
    if center:
        y = np.pad(y, n_fft // 2, mode="constant")
    n_frames = 1 + (len(y) - n_fft) // hop_length
    win = hann_periodic(n_fft)
    out = np.empty((max(n_frames, 0), 1 + n_fft // 2), dtype=np.complex64)
    for t in range(n_frames):
        frame = y[t * hop_length:t * hop_length + n_fft]
        out[t] = fftpack.fft(win * frame)[:1 + n_fft // 2]
    return out


def melspectrogram(y, sr=22050, n_fft=2048, hop_length=512, n_mels=128, power=2.0, fmin=0.0, fmax=None):
    """librosa.feature.melspectrogram -> (frames, n_mels) float32."""
    S = np.abs(stft(y, n_fft=n_fft, hop_length=hop_length)) ** power
    basis = mel(sr=sr, n_fft=n_fft, n_mels=n_mels, fmin=fmin, fmax=fmax)
    return np.einsum("tf,mf->tm", S.astype(np.float32), basis, optimize=True).astype(np.float32)


def power_to_db(S, ref=1.0, amin=1e-10, top_db=80.0):
    """librosa.power_to_db."""
    magnitude = np.asarray(S)
    log_spec = 10.0 * np.log10(np.maximum(amin, magnitude))
    log_spec -= 10.0 * np.log10(np.maximum(amin, ref))
    if top_db is not None:
        log_spec = np.maximum(log_spec, log_spec.max() - top_db)
    return log_spec


def onset_strength(y, sr=22050, lag=1, center=True, aggregate=None, n_fft=2048, hop_length=512, n_mels=128,
                   return_db=False):
    """librosa.onset.onset_strength -> (frames,) float32; aggregate: None / np.mean / np.median."""
    if aggregate is None:
        aggregate = np.mean
    S = power_to_db(np.abs(melspectrogram(y, sr=sr, n_fft=n_fft, hop_length=hop_length, n_mels=n_mels)))
    if return_db:
        return S
    T = S.shape[0]
    if T <= lag:
        env = np.zeros((0,), S.dtype)
    else:
        env = aggregate(np.maximum(0.0, S[lag:] - S[:-lag]), axis=1)
    pad_width = lag
    if center:
        pad_width += n_fft // (2 * hop_length)
    env = np.pad(env, (int(pad_width), 0), mode="constant")
    if center:
        env = env[:T]
    return env.astype(np.float32)
