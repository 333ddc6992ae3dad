"""Host-side filterbank construction with madmom.audio.filters' API and bit-identical results.

Built once per plan; the float32 weights are uploaded unchanged in banded
form.  Follows madmom 0.16.1 ``madmom/audio/filters.py`` (log_frequencies, frequencies2bins,
TriangularFilter, Filterbank.from_filters, LogarithmicFilterbank, PitchClassProfileFilterbank),
reached from /root/reference/backend/app/services/grid/beats.py:74 through
FilteredSpectrogramProcessor's constructor.
"""
from __future__ import annotations

import numpy as np

FILTER_DTYPE = np.float32
A4 = 440.0
FMIN, FMAX, NUM_BANDS = 30.0, 17000.0, 12
NORM_FILTERS, UNIQUE_FILTERS = True, True


def hz2midi(f, fref=A4):
    return 12.0 * np.log2(np.asarray(f, dtype=float) / fref) + 69.0


def midi2hz(m, fref=A4):
    return 2.0 ** ((np.asarray(m, dtype=float) - 69.0) / 12.0) * fref


def fft_frequencies(num_fft_bins, sample_rate):
    """madmom.audio.stft.fft_frequencies."""
    return np.fft.fftfreq(num_fft_bins * 2, 1.0 / sample_rate)[:num_fft_bins]


def log_frequencies(bands_per_octave, fmin, fmax, fref=A4):
    lo = np.floor(np.log2(float(fmin) / fref) * bands_per_octave)
    hi = np.ceil(np.log2(float(fmax) / fref) * bands_per_octave)
    freqs = fref * 2.0 ** (np.arange(lo, hi) / float(bands_per_octave))
    freqs = freqs[np.searchsorted(freqs, fmin):]
    return freqs[:np.searchsorted(freqs, fmax, "right")]


def frequencies2bins(frequencies, bin_frequencies, unique_bins=False):
    frequencies = np.asarray(frequencies)
    bin_frequencies = np.asarray(bin_frequencies)
    idx = bin_frequencies.searchsorted(frequencies)
    idx = np.clip(idx, 1, len(bin_frequencies) - 1)
    below, above = bin_frequencies[idx - 1], bin_frequencies[idx]
    idx -= frequencies - below < above - frequencies
    return np.unique(idx) if unique_bins else idx


class Filter(np.ndarray):
    """1-D float32 filter positioned at bin `start` (madmom.audio.filters.Filter)."""

    def __new__(cls, data, start=0, norm=False):
        if not (isinstance(data, np.ndarray) and data.ndim == 1) and not isinstance(data, (list, tuple)):
            raise TypeError("wrong input data for Filter, must be np.ndarray")
        obj = np.asarray(data, dtype=FILTER_DTYPE).view(cls)
        if obj.ndim != 1:
            raise NotImplementedError("please add multi-dimension support")
        if norm:
            obj /= np.sum(obj)
        obj.start = int(start)
        obj.stop = int(start + len(data))
        return obj

    def __array_finalize__(self, obj):
        if obj is None:
            return
        self.start = getattr(obj, "start", 0)
        self.stop = getattr(obj, "stop", 0)


class TriangularFilter(Filter):
    def __new__(cls, start, center, stop, norm=False):
        if not start <= center < stop:
            raise ValueError("`center` must be between `start` and `stop`")
        start, center, stop = int(start), int(center), int(stop)
        rel_center, rel_stop = center - start, stop - start
        data = np.zeros(rel_stop)
        data[:rel_center] = np.linspace(0, 1, rel_center, endpoint=False)
        data[rel_center:] = np.linspace(1, 0, rel_stop - rel_center, endpoint=False)
        obj = Filter.__new__(cls, data, start, norm)
        obj.center = center
        return obj

    @classmethod
    def band_bins(cls, bins, overlap=True):
        if len(bins) < 3:
            raise ValueError("not enough bins to create a TriangularFilter")
        for i in range(len(bins) - 2):
            start, center, stop = bins[i:i + 3]
            if not overlap:
                start = int(np.floor((center + start) / 2.0))
                stop = int(np.ceil((center + stop) / 2.0))
            if stop - start < 2:          # too-small filter: a single bin of weight 1
                center = start
                stop = start + 1
            yield start, center, stop

    @classmethod
    def filters(cls, bins, norm, overlap=True):
        return [cls(s, c, e, norm) for s, c, e in cls.band_bins(bins, overlap)]


class Filterbank(np.ndarray):
    """float32 (num_bins, num_bands) matrix + bin_frequencies (madmom.audio.filters.Filterbank)."""

    def __new__(cls, data, bin_frequencies):
        if not (isinstance(data, np.ndarray) and data.ndim == 2):
            raise TypeError("wrong input data for Filterbank, must be a 2D np.ndarray")
        obj = np.asarray(data, dtype=FILTER_DTYPE).view(cls)
        if len(bin_frequencies) != obj.shape[0]:
            raise ValueError("`bin_frequencies` must have the same length as the first dimension of `data`.")
        obj.bin_frequencies = np.asarray(bin_frequencies, dtype=float)
        return obj

    def __array_finalize__(self, obj):
        if obj is None:
            return
        self.bin_frequencies = getattr(obj, "bin_frequencies", None)

    @classmethod
    def _put_filter(cls, filt, band):
        start, stop = filt.start, filt.start + len(filt)
        data = np.asarray(filt)
        if start < 0:
            data = data[-start:]
            start = 0
        if stop > len(band):
            data = data[:-(stop - len(band))]
            stop = len(band)
        region = band[start:stop]
        np.maximum(data, region, out=region)

    @classmethod
    def from_filters(cls, filters, bin_frequencies):
        fb = np.zeros((len(bin_frequencies), len(filters)))
        for band_id, band_filter in enumerate(filters):
            if isinstance(band_filter, Filter):
                band_filter = [band_filter]
            for filt in band_filter:
                cls._put_filter(filt, fb[:, band_id])
        return Filterbank.__new__(cls, fb, bin_frequencies)

    @property
    def num_bins(self):
        return self.shape[0]

    @property
    def num_bands(self):
        return self.shape[1]

    @property
    def corner_frequencies(self):
        out = []
        for band in range(self.num_bands):
            nz = np.nonzero(np.asarray(self)[:, band])[0]
            out.append([np.min(nz), np.max(nz)])
        return self.bin_frequencies[out].T

    @property
    def center_frequencies(self):
        data = np.asarray(self)
        out = []
        for band in range(self.num_bands):
            nz = np.nonzero(data[:, band])[0]
            lo, hi = np.min(nz), np.max(nz)
            if data[lo, band] == data[hi, band]:
                out.append(int(lo + (hi - lo) / 2.0))
            else:
                out.append(lo + np.argmax(data[lo:hi, band]))
        return self.bin_frequencies[out]

    @property
    def fmin(self):
        return self.bin_frequencies[np.nonzero(np.asarray(self))[0][0]]

    @property
    def fmax(self):
        return self.bin_frequencies[np.nonzero(np.asarray(self))[0][-1]]

    # ---- banded form for the device -----------------------------------------------------------
    def banded(self):
        """(band_start, band_len, band_woff, weights): contiguous support of every band.

        Zeros inside a band's support are kept (weight 0.0), so any Filterbank is representable;
        madmom's triangular banks have none.
        """
        data = np.asarray(self)
        B = data.shape[1]
        start = np.zeros(B, np.int32)
        length = np.zeros(B, np.int32)
        weights = []
        for j in range(B):
            nz = np.nonzero(data[:, j])[0]
            if nz.size:
                start[j], length[j] = nz[0], nz[-1] - nz[0] + 1
                weights.append(data[nz[0]:nz[-1] + 1, j])
        woff = np.zeros(B, np.int32)
        woff[1:] = np.cumsum(length)[:-1]
        w = np.concatenate(weights).astype(np.float32) if weights else np.zeros(0, np.float32)
        return start, length, woff, np.ascontiguousarray(w)


class LogarithmicFilterbank(Filterbank):
    NUM_BANDS_PER_OCTAVE = 12

    def __new__(cls, bin_frequencies, num_bands=NUM_BANDS, fmin=FMIN, fmax=FMAX, fref=A4,
                norm_filters=NORM_FILTERS, unique_filters=UNIQUE_FILTERS, bands_per_octave=True):
        if not bands_per_octave:
            raise NotImplementedError("please implement `num_bands` with `bands_per_octave` set to 'False'")
        freqs = log_frequencies(num_bands, fmin, fmax, fref)
        bins = frequencies2bins(freqs, bin_frequencies, unique_bins=unique_filters)
        filters = TriangularFilter.filters(bins, norm=norm_filters, overlap=True)
        obj = cls.from_filters(filters, bin_frequencies)
        obj.fref = fref
        obj.num_bands_per_octave = num_bands
        obj.norm_filters = norm_filters
        obj.unique_filters = unique_filters
        return obj

    def __array_finalize__(self, obj):
        Filterbank.__array_finalize__(self, obj)
        if obj is None:
            return
        for name in ("fref", "num_bands_per_octave", "norm_filters", "unique_filters"):
            setattr(self, name, getattr(obj, name, None))


LogFilterbank = LogarithmicFilterbank


class PitchClassProfileFilterbank(Filterbank):
    """madmom.audio.filters.PitchClassProfileFilterbank: bins -> 12 classes, class 0 = fref's class."""

    CLASSES, FMIN, FMAX = 12, 100.0, 5000.0

    def __new__(cls, bin_frequencies, num_classes=12, fmin=100.0, fmax=5000.0, fref=A4):
        bin_frequencies = np.asarray(bin_frequencies, dtype=float)
        fb = np.zeros((len(bin_frequencies), num_classes))
        with np.errstate(divide="ignore", invalid="ignore"):
            log_dev = np.log2(bin_frequencies / fref)
            classes = np.round(num_classes * log_dev) % num_classes
        rows = np.arange(len(fb))[1:]                    # bin 0 (0 Hz) has no pitch
        fb[rows, classes[1:].astype(int)] = 1
        fb[np.searchsorted(bin_frequencies, fmax, "right"):] = 0
        fb[:np.searchsorted(bin_frequencies, fmin)] = 0
        obj = Filterbank.__new__(cls, fb, bin_frequencies)
        obj.fref = fref
        return obj

    def __array_finalize__(self, obj):
        Filterbank.__array_finalize__(self, obj)
        if obj is None:
            return
        self.fref = getattr(obj, "fref", None)


def fold_classes(center_frequencies, num_classes=12):
    """Pitch class (0 = C) of each band centre, as madmom.audio.chroma.CLPChroma folds bands."""
    midi = np.round(hz2midi(center_frequencies)).astype(int)
    return np.mod(midi, num_classes)


# ---- librosa-style mel filterbank (onset strength) -------------------------------------
def _hz_to_mel(f):
    """librosa.hz_to_mel (Slaney scale: linear below 1 kHz, logarithmic above)."""
    f = np.asanyarray(f, dtype=float)
    f_sp = 200.0 / 3
    mels = f / f_sp
    min_log_hz, logstep = 1000.0, np.log(6.4) / 27.0
    min_log_mel = min_log_hz / f_sp
    if f.ndim:
        log_t = f >= min_log_hz
        mels[log_t] = min_log_mel + np.log(f[log_t] / min_log_hz) / logstep
    elif f >= min_log_hz:
        mels = min_log_mel + np.log(f / min_log_hz) / logstep
    return mels


def _mel_to_hz(mels):
    """librosa.mel_to_hz (Slaney scale)."""
    mels = np.asanyarray(mels, dtype=float)
    f_sp = 200.0 / 3
    freqs = f_sp * mels
    min_log_hz, logstep = 1000.0, np.log(6.4) / 27.0
    min_log_mel = min_log_hz / f_sp
    if mels.ndim:
        log_t = mels >= min_log_mel
        freqs[log_t] = min_log_hz * np.exp(logstep * (mels[log_t] - min_log_mel))
    elif mels >= min_log_mel:
        freqs = min_log_hz * np.exp(logstep * (mels - min_log_mel))
    return freqs


class SlaneyMelFilterbank(Filterbank):
    """``librosa.filters.mel(sr, n_fft, n_mels, fmin, fmax, htk=False, norm='slaney')`` as a Filterbank.

    librosa 0.10.2 (pinned at /root/reference/backend/requirements.txt:14) builds ``(n_mels, 1 + n_fft/2)``
    float32 weights; the reference reaches them through ``librosa.onset.onset_strength``
    (services/accompaniment/strum.py:114, services/analysis/content_classifier.py:48,92).  This class is
    the transpose without the Nyquist row: the FFT kernels produce bins ``0 .. n_fft/2 - 1``, and with the
    default ``fmax = sr/2`` the last triangle reaches zero exactly at the Nyquist bin (its weight there is
    ~1e-17 from rounding), so nothing measurable is dropped.  A smaller ``fmax`` is exact.
    """

    def __new__(cls, sample_rate, n_fft, n_mels=128, fmin=0.0, fmax=None):
        if fmax is None:
            fmax = float(sample_rate) / 2
        fftfreqs = np.fft.rfftfreq(n=int(n_fft), d=1.0 / sample_rate)
        mel_f = _mel_to_hz(np.linspace(_hz_to_mel(fmin), _hz_to_mel(fmax), int(n_mels) + 2))
        fdiff = np.diff(mel_f)
        ramps = np.subtract.outer(mel_f, fftfreqs)
        weights = np.zeros((int(n_mels), len(fftfreqs)), dtype=FILTER_DTYPE)
        for i in range(int(n_mels)):
            lower = -ramps[i] / fdiff[i]
            upper = ramps[i + 2] / fdiff[i + 1]
            weights[i] = np.maximum(0, np.minimum(lower, upper))
        enorm = 2.0 / (mel_f[2:int(n_mels) + 2] - mel_f[:int(n_mels)])
        weights *= enorm[:, np.newaxis]
        obj = Filterbank.__new__(cls, np.ascontiguousarray(weights[:, :-1].T), fftfreqs[:-1])
        obj.mel_frequencies = mel_f
        return obj

    def __array_finalize__(self, obj):
        Filterbank.__array_finalize__(self, obj)
        if obj is None:
            return
        self.mel_frequencies = getattr(obj, "mel_frequencies", None)
