"""Independent float64 cross-check of the oracle (oracle/madmom_ref.py).

madmom itself cannot run here (parity unpinned), so the oracle's arithmetic is checked
against a SECOND statement of the same mathematics that shares no code with it and is built differently:

* framing by explicit index arithmetic on a zero-padded copy (no ``signal_frame``),
* the STFT as one vectorised ``np.fft.rfft`` over the gathered frames (pocketfft, not ``scipy.fftpack``),
  and as a direct O(N^2) DFT (matrix of complex exponentials evaluated in float64) on a few frames,
* the logarithmic filterbank from the closed form of the triangles (slopes, no ``linspace`` / ``np.maximum``
  placement) with nearest-bin rounding written as ``floor(x + 1/2)``,
* log / lagged difference / stacking written out with plain slicing.

All four frame sizes of the path (1024 / 2048 / 4096 of RNNBeatProcessor, grid/beats.py:74; 8192 of
DeepChromaProcessor / CNNKeyRecognitionProcessor, chords/extract.py:54, theory/key.py:101), float32 and
int16 input, integer and non-integer hops.
"""
import numpy as np
import pytest

from audio_tabs_b200.synth import synth_guitar
from oracle import madmom_ref as ref

SR = 44100


# ---- the second statement ---------------------------------------------------------------------------
def gather_frames(x, frame_size, hop, origin=0):
    n = len(x)
    T = int(np.ceil(n / hop))
    pad = frame_size + int(abs(origin)) + 1
    padded = np.concatenate([np.zeros(pad, x.dtype), x, np.zeros(pad + int(hop) + frame_size, x.dtype)])
    starts = np.array([int(t * hop) for t in range(T)], dtype=np.int64) - frame_size // 2 - origin
    idx = starts[:, None] + np.arange(frame_size)[None, :] + pad
    return padded[idx]


def stft64(x, frame_size, hop):
    frames = gather_frames(x, frame_size, hop).astype(np.float64)
    win = 0.5 - 0.5 * np.cos(2.0 * np.pi * np.arange(frame_size) / (frame_size - 1))     # np.hanning, written out
    if np.issubdtype(x.dtype, np.integer):
        win = win / float(np.iinfo(x.dtype).max)
    return np.fft.rfft(frames * win, axis=1)[:, :frame_size // 2]


def dft_direct(frame64, bins):
    n = np.arange(len(frame64))
    return np.array([np.sum(frame64 * np.exp(-2j * np.pi * k * n / len(frame64))) for k in bins])


def log_filterbank64(frame_size, bpo, fmin, fmax, unique=True, norm=True, fref=440.0):
    nbins = frame_size // 2
    df = SR / frame_size
    lo, hi = int(np.floor(np.log2(fmin / fref) * bpo)), int(np.ceil(np.log2(fmax / fref) * bpo))
    freqs = [fref * 2.0 ** (i / bpo) for i in range(lo, hi)]
    freqs = [f for f in freqs if fmin <= f <= fmax]
    bins = [min(max(int(np.floor(f / df + 0.5)), 0), nbins - 1) for f in freqs]            # nearest bin, ties up
    if unique:
        bins = sorted(set(bins))
    fb = np.zeros((nbins, len(bins) - 2))
    for j in range(len(bins) - 2):
        start, center, stop = bins[j], bins[j + 1], bins[j + 2]
        if stop - start < 2:
            center, stop = start, start + 1
        tri = np.zeros(nbins)
        for b in range(start, min(stop, nbins)):
            tri[b] = (b - start) / (center - start) if b < center else (stop - b) / (stop - center)
        if norm:
            tri = tri / tri.sum()
        fb[:, j] = tri
    return fb


def front_end64(x, frame_size, hop, bpo, fmin, fmax, mul=1.0, add=1.0, diff_ratio=None):
    mag = np.abs(stft64(x, frame_size, hop).astype(np.complex64)).astype(np.float32)      # madmom stores complex64
    fb = log_filterbank64(frame_size, bpo, fmin, fmax).astype(np.float32)
    L = np.log10(mul * (mag.astype(np.float64) @ fb.astype(np.float64)) + add)
    if diff_ratio is None:
        return L
    win = np.hanning(frame_size)
    k = int(max(1, round((frame_size / 2 - np.argmax(win > diff_ratio * win.max())) / hop)))
    D = np.zeros_like(L)
    D[k:] = np.maximum(L[k:] - L[:-k], 0.0)
    return np.hstack([L, D])


# ---- tests ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("frame_size", [1024, 2048, 4096, 8192])
@pytest.mark.parametrize("hop", [441.0, 4410.0, 8820.0, SR / 7.0])
def test_frame_geometry(frame_size, hop):
    x = np.arange(1, 30001, dtype=np.float32)
    fs = ref.FramedSignal(ref.Signal(x, sample_rate=SR), frame_size=frame_size, hop_size=hop)
    mine = gather_frames(x, frame_size, hop)
    assert len(fs) == mine.shape[0] == int(np.ceil(30000 / hop))
    for t in range(len(fs)):
        np.testing.assert_array_equal(fs[t], mine[t])


@pytest.mark.parametrize("frame_size", [1024, 2048, 4096, 8192])
@pytest.mark.parametrize("dtype", ["f32", "i16"])
def test_stft_against_rfft_and_direct_dft(frame_size, dtype):
    x = synth_guitar(500 + frame_size, 0.75)
    if dtype == "i16":
        x = (x * 30000).astype(np.int16)
    hop = 441.0 if frame_size <= 4096 else 4410.0
    got = ref.ShortTimeFourierTransform(ref.FramedSignal(ref.Signal(x, sample_rate=SR), frame_size=frame_size,
                                                         hop_size=hop)).data
    want = stft64(x, frame_size, hop)
    assert got.dtype == np.complex64 and got.shape == want.shape
    peak = np.abs(want).max(axis=1, keepdims=True)
    # complex64 rounding of a float64 transform: half an ulp per component
    assert (np.abs(got - want) <= 1.5 * np.finfo(np.float32).eps * peak + 1e-30).all()
    # a direct O(N^2) DFT on three frames (first: left zero padding; middle; last: right zero padding)
    frames = gather_frames(x, frame_size, hop).astype(np.float64)
    win = np.hanning(frame_size) / (32767.0 if dtype == "i16" else 1.0)
    bins = np.unique(np.concatenate([np.arange(0, 8), np.linspace(8, frame_size // 2 - 1, 24).astype(int)]))
    for t in (0, len(frames) // 2, len(frames) - 1):
        direct = dft_direct(frames[t] * win, bins)
        assert np.abs(got[t, bins] - direct).max() <= 1.5 * np.finfo(np.float32).eps * max(peak[t, 0], 1e-30)


@pytest.mark.parametrize("frame_size,bpo,fmin,fmax,bands", [
    (1024, 3, 30, 17000, 21), (2048, 6, 30, 17000, 45), (4096, 12, 30, 17000, 91), (2048, 12, 30, 17000, 81),
    (8192, 24, 65, 2100, 105), (8192, 24, 60, 2600, 113), (4096, 24, 65, 2100, 87), (1024, 6, 30, 17000, 39),
    (4096, 6, 30, 17000, 49)])
def test_filterbank_against_closed_form(frame_size, bpo, fmin, fmax, bands):
    fb = ref.LogarithmicFilterbank(ref.fft_frequencies(frame_size // 2, SR), num_bands=bpo, fmin=fmin, fmax=fmax).data
    mine = log_filterbank64(frame_size, bpo, fmin, fmax)
    assert fb.shape == mine.shape == (frame_size // 2, bands)
    np.testing.assert_array_equal(fb != 0, mine != 0)                     # same support, bin for bin
    np.testing.assert_allclose(fb, mine, rtol=3e-7, atol=0)               # float32 triangle / float32 sum / divide
    fbu = ref.LogarithmicFilterbank(ref.fft_frequencies(1024, SR), num_bands=12, unique_filters=False).data
    assert fbu.shape == log_filterbank64(2048, 12, 30, 17000, unique=False).shape == (1024, 108)


def test_beat_front_end_end_to_end():
    x = synth_guitar(4242, 1.0)
    want = np.hstack([front_end64(x, f, 441.0, b, 30, 17000, diff_ratio=0.5) for f, b in ((1024, 3), (2048, 6), (4096, 12))])
    got = ref.rnn_beat_preprocessor()(x)
    assert got.shape == want.shape == (100, 314) and got.dtype == np.float32
    np.testing.assert_allclose(got, want, rtol=2e-6, atol=2e-6)          # float32 dot / log10 against float64


def test_onset_front_end_end_to_end():
    x = synth_guitar(4243, 1.0)
    want = np.hstack([front_end64(x, f, 441.0, 6, 30, 17000, mul=5.0, diff_ratio=0.25) for f in (1024, 2048, 4096)])
    got = ref.rnn_onset_preprocessor()(x)
    assert got.shape == want.shape == (100, 266)
    np.testing.assert_allclose(got, want, rtol=2e-6, atol=4e-6)


@pytest.mark.parametrize("fps,rows", [(10, 30), (5, 15)])
def test_chroma_key_front_end_int16(fps, rows):
    x = (synth_guitar(4244, 3.0) * 25000).astype(np.int16)
    want = front_end64(x, 8192, SR / fps, 24, 65, 2100)
    got = ref.log_filt_chain(8192, fps=fps)(x).data
    assert got.shape == want.shape == (rows, 105)
    np.testing.assert_allclose(got, want, rtol=2e-6, atol=2e-6)


def test_committed_goldens_against_the_second_statement():
    """the fixtures the GPU tests compare against agree with the independent float64 statement as well"""
    from pathlib import Path
    gold = Path(__file__).resolve().parent / "golden"
    x = np.load(gold / "guitar_2s_f32.npy")
    want = np.hstack([front_end64(x, f, 441.0, b, 30, 17000, diff_ratio=0.5) for f, b in ((1024, 3), (2048, 6), (4096, 12))])
    np.testing.assert_allclose(np.load(gold / "guitar_2s_beat314.npy"), want, rtol=2e-6, atol=2e-6)
    clip = np.load(gold / "refjob_3s_i16.npy")
    np.testing.assert_allclose(np.load(gold / "refjob_3s_deepchroma105.npy"), front_end64(clip, 8192, 4410.0, 24, 65, 2100),
                               rtol=2e-6, atol=2e-6)
    np.testing.assert_allclose(np.load(gold / "refjob_3s_key105.npy"), front_end64(clip, 8192, 8820.0, 24, 65, 2100),
                               rtol=2e-6, atol=2e-6)


# ---- the librosa restatement (N3 row) against a third-party port of the same functions -------------------------------
@pytest.mark.parametrize("frame_size,hop,dtype", [(1024, 441, "f32"), (2048, 441, "f32"), (4096, 441, "i16"), (8192, 4410, "f32"),
                                                  (2048, 512, "f32")])
def test_stft_against_scipy_signal_stft(frame_size, hop, dtype):
    """EXTERNAL second opinion on madmom's framing + window + transform: scipy.signal.stft with boundary='zeros'
    centres frame n on sample n * hop over zero padding, exactly FramedSignal's origin 0 -- a third-party
    implementation of the same definition, not written for this repo.  (scipy divides by the window sum, and an
    int16 signal enters madmom's transform through a window divided by 32767.)"""
    import scipy.signal as ss
    x = synth_guitar(1200 + frame_size, 1.37)
    if dtype == "i16":
        x = np.clip(np.round(x * 25000), -32768, 32767).astype(np.int16)
    st = ref.ShortTimeFourierTransform(ref.FramedSignal(ref.Signal(x, sample_rate=44100), frame_size=frame_size, hop_size=hop))
    w = np.hanning(frame_size)
    xs = x.astype(np.float64) / (32767.0 if dtype == "i16" else 1.0)
    _, _, Z = ss.stft(xs, fs=44100, window=w, nperseg=frame_size, noverlap=frame_size - hop, boundary="zeros", padded=True,
                      return_onesided=True, scaling="spectrum")
    Z = Z.T * w.sum()
    T = st.data.shape[0]
    assert T == int(np.ceil(len(x) / hop)) and Z.shape[0] >= T
    err = np.abs(st.data - Z[:T, :frame_size // 2]).max()
    assert err <= 2.5e-7 * np.abs(Z).max(), (err, np.abs(Z).max())       # complex64 storage of the oracle's float64 transform
    # with the Nyquist bin
    sn = ref.ShortTimeFourierTransform(ref.FramedSignal(ref.Signal(x, sample_rate=44100), frame_size=frame_size, hop_size=hop),
                                       include_nyquist=True)
    assert np.abs(sn.data - Z[:T]).max() <= 2.5e-7 * np.abs(Z).max()


def test_librosa_restatement_against_transformers_audio_utils():
    """oracle/librosa_ref.py (Slaney mel filterbank, centred zero-padded periodic-Hann STFT, power mel spectrogram,
    power_to_db) against `transformers.audio_utils` -- an independent numpy port of the same librosa functions that
    happens to be installed in this image (librosa itself is not).  Not our code, not the oracle's author: the
    closest thing to an external pin this row has."""
    au = pytest.importorskip("transformers.audio_utils")
    from oracle import librosa_ref as lr
    for sr, n_fft, n_mels in ((44100, 2048, 128), (22050, 2048, 128), (44100, 1024, 40)):
        theirs = au.mel_filter_bank(num_frequency_bins=1 + n_fft // 2, num_mel_filters=n_mels, min_frequency=0.0,
                                    max_frequency=sr / 2.0, sampling_rate=sr, norm="slaney", mel_scale="slaney")
        ours = lr.mel(sr=sr, n_fft=n_fft, n_mels=n_mels)
        assert theirs.shape == ours.T.shape
        np.testing.assert_array_equal(theirs != 0, ours.T != 0)                       # same support
        np.testing.assert_allclose(ours.T, theirs, rtol=2e-5, atol=1e-9)              # float32 table against float64
    y = synth_guitar(4100, 2.0)
    mel = au.mel_filter_bank(num_frequency_bins=1025, num_mel_filters=128, min_frequency=0.0, max_frequency=SR / 2.0,
                             sampling_rate=SR, norm="slaney", mel_scale="slaney")
    theirs_db = au.spectrogram(y.astype(np.float64), au.window_function(2048, "hann", periodic=True), frame_length=2048,
                               hop_length=512, power=2.0, center=True, pad_mode="constant", mel_filters=mel,
                               mel_floor=1e-10, log_mel="dB", reference=1.0, min_value=1e-10, db_range=None,
                               dtype=np.float64).T
    ours_db = lr.power_to_db(lr.melspectrogram(y, sr=SR), top_db=None)
    assert theirs_db.shape == ours_db.shape == (1 + len(y) // 512, 128)
    np.testing.assert_allclose(ours_db, theirs_db, rtol=0, atol=2e-3)                 # dB; float32 einsum in the restatement
    # power_to_db with a range clip == librosa's top_db
    np.testing.assert_allclose(lr.power_to_db(lr.melspectrogram(y, sr=SR), top_db=80.0),
                               au.power_to_db(lr.melspectrogram(y, sr=SR).astype(np.float64), reference=1.0, min_value=1e-10,
                                              db_range=80.0), rtol=0, atol=1e-4)
