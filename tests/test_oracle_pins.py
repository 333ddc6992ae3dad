"""Oracle pins: structural constants of madmom 0.16.1, analytic known-answer tests
and the committed golden fixtures.  The oracle is test infrastructure; "parity unpinned" by the
reference itself (it ships no vectors for this path), so these are what hold it in place."""
from pathlib import Path

import numpy as np
import pytest

from oracle import madmom_ref as ref

GOLD = Path(__file__).resolve().parent / "golden"
SR = 44100


def bins(frame_size):
    return ref.fft_frequencies(frame_size >> 1, SR)


@pytest.mark.parametrize("frame_size,bpo,fmin,fmax,unique,bands", [
    (2048, 12, 30, 17000, True, 81), (2048, 12, 30, 17000, False, 108),
    (8192, 24, 65, 2100, True, 105), (8192, 24, 60, 2600, True, 113),
    (1024, 3, 30, 17000, True, 21), (2048, 6, 30, 17000, True, 45), (4096, 12, 30, 17000, True, 91),
    (1024, 6, 30, 17000, True, 39), (4096, 6, 30, 17000, True, 49), (4096, 24, 65, 2100, True, 87),
])
def test_band_counts(frame_size, bpo, fmin, fmax, unique, bands):
    fb = ref.LogarithmicFilterbank(bins(frame_size), num_bands=bpo, fmin=fmin, fmax=fmax, unique_filters=unique)
    assert fb.num_bands == bands
    assert fb.data.dtype == np.float32


def test_feature_widths_314_266_1575():
    x = np.zeros(SR, np.float32)
    assert ref.rnn_beat_preprocessor()(x).shape == (100, 314)
    assert ref.rnn_onset_preprocessor()(x).shape == (100, 266)
    spec = ref.log_filt_chain(8192, fps=10)(x).data
    assert spec.shape == (10, 105)
    assert ref.dcp_context(spec).shape == (10, 1575)


def test_num_frames_sample_wav_281():
    assert ref.num_frames_for(123481, 441.0) == 281
    assert ref.num_frames_for(123481, 441.0, "extend") == 281
    assert ref.num_frames_for(441 * 5, 441.0) == 5 and ref.num_frames_for(441 * 5, 441.0, "extend") == 6
    with pytest.raises(ValueError):
        ref.num_frames_for(10, 441.0, "bogus")


def test_diff_frames():
    assert [ref.diff_frames_for(0.5, 441.0, f) for f in (1024, 2048, 4096)] == [1, 1, 2]
    assert [ref.diff_frames_for(0.25, 441.0, f) for f in (1024, 2048, 4096)] == [1, 2, 3]


def test_filterbank_structure():
    for frame_size, bpo in ((1024, 3), (2048, 6), (4096, 12), (2048, 12), (8192, 24)):
        fb = ref.LogarithmicFilterbank(bins(frame_size), num_bands=bpo).data
        assert ((fb != 0).sum(axis=1) <= 2).all()                 # every FFT bin feeds at most 2 bands
        np.testing.assert_allclose(fb.sum(axis=0), 1.0, atol=1e-6)  # area-normalised triangles
        for j in range(fb.shape[1]):                              # contiguous support
            nz = np.nonzero(fb[:, j])[0]
            assert nz[-1] - nz[0] + 1 == len(nz)


def test_stft_dtype_and_bins():
    x = np.random.default_rng(0).standard_normal(8000).astype(np.float32)
    s = ref.ShortTimeFourierTransform(ref.FramedSignal(ref.Signal(x, sample_rate=SR), frame_size=2048))
    assert s.data.dtype == np.complex64 and s.data.shape == (19, 1024)
    assert s.bin_frequencies[1] == pytest.approx(SR / 2048)


def test_kat_zero_input_is_exactly_zero():
    out = ref.rnn_beat_preprocessor()(np.zeros(SR // 2, np.float32))
    assert out.dtype == np.float32 and not out.any()


def test_kat_bin_centred_cosine():
    F, k0, A = 2048, 100, 0.5
    n = np.arange(SR)
    x = (A * np.cos(2 * np.pi * k0 * n / F)).astype(np.float32)
    s = ref.ShortTimeFourierTransform(ref.FramedSignal(ref.Signal(x, sample_rate=SR), frame_size=F))
    mag = np.abs(s.data[10])
    assert np.argmax(mag) == k0
    assert mag[k0] == pytest.approx(A * np.hanning(F).sum() / 2, rel=1e-4)


def test_kat_impulse_flat_spectrum():
    F = 1024
    x = np.zeros(SR, np.float32)
    x[5000] = 1.0
    fr = ref.FramedSignal(ref.Signal(x, sample_rate=SR), frame_size=F)
    n = 11                                    # frame 11 starts at 4851 - 512 = 4339 -> impulse at offset 661
    start = ref.frame_start(n, F, 441.0)
    s = ref.ShortTimeFourierTransform(fr)
    np.testing.assert_allclose(np.abs(s.data[n]), np.hanning(F)[5000 - start], rtol=1e-5)


def test_int16_window_scaling():
    rng = np.random.default_rng(3)
    xf = (rng.standard_normal(20000) * 0.2).clip(-1, 1)
    xi = np.round(xf * 32767).astype(np.int16)
    a = ref.log_filt_chain(8192, fps=5)(xi).data
    b = ref.log_filt_chain(8192, fps=5)((xi / 32767.0).astype(np.float32)).data
    np.testing.assert_allclose(a, b, rtol=1e-5, atol=1e-6)


def test_diff_first_rows_zero_and_positive():
    x = np.random.default_rng(1).standard_normal(SR).astype(np.float32)
    out = ref.rnn_beat_preprocessor()(x)
    for lo, b, k in ((0, 21, 1), (42, 45, 1), (132, 91, 2)):
        d = out[:, lo + b:lo + 2 * b]
        assert not d[:k].any() and (d >= 0).all()
        s = out[:, lo:lo + b]
        np.testing.assert_array_equal(d[k:], np.maximum(s[k:] - s[:-k], 0))


def test_fold_chroma_and_pcp():
    fb = ref.LogarithmicFilterbank(bins(4096), num_bands=24, fmin=65, fmax=2100)
    cls = ref.fold_classes(fb.center_frequencies)
    assert cls.min() == 0 and cls.max() == 11
    a440 = np.argmin(np.abs(fb.center_frequencies - 440.0))
    assert cls[a440] == 9                                       # A -> class 9 when C = 0
    pcp = ref.PitchClassProfileFilterbank(bins(4096))
    assert pcp.data.shape == (2048, 12) and set(np.unique(pcp.data)) == {0.0, 1.0}
    k440 = np.argmin(np.abs(bins(4096) - 440.0))
    assert pcp.data[k440, 0] == 1                               # class 0 = pitch class of fref


@pytest.mark.parametrize("name,fn", [
    ("guitar_2s_beat314", lambda x: ref.rnn_beat_preprocessor()(x)),
    ("guitar_2s_onset266", lambda x: ref.rnn_onset_preprocessor()(x)),
    ("guitar_2s_logfilt81", lambda x: ref.log_filtered_spectrogram(x)),
])
def test_oracle_reproduces_golden(name, fn):
    x = np.load(GOLD / "guitar_2s_f32.npy")
    np.testing.assert_array_equal(fn(x), np.load(GOLD / (name + ".npy")))


def test_oracle_reproduces_golden_real_audio():
    clip = np.load(GOLD / "refjob_3s_i16.npy")
    assert clip.dtype == np.int16
    np.testing.assert_array_equal(ref.log_filt_chain(8192, fps=10)(clip).data, np.load(GOLD / "refjob_3s_deepchroma105.npy"))
    np.testing.assert_array_equal(ref.log_filt_chain(8192, fps=5)(clip).data, np.load(GOLD / "refjob_3s_key105.npy"))
    n, t = np.load(GOLD / "refjob_total_frames.npy")
    assert (n, t) == (675192, 1532)                             # SURVEY.md §8c pin (6)
