"""librosa-style onset strength: oracle self-checks and the host-side pieces on the CPU,
GPU parity against oracle/librosa_ref.py (parity unpinned: librosa is not installable here).

Tolerances (stated here because they differ from the madmom rows): the dB mel spectrogram is compared
after the 80 dB top_db clip with atol 0.05 dB -- 80 dB below the clip's maximum is where the float32 FFT's
rounding noise (~1e-7 of the frame peak) reaches ~1e-3 of a band's amplitude; the envelope (a mean or
median of 128 positive differences) with rtol 1e-3, atol 5e-3 dB."""
import numpy as np
import pytest

from oracle import librosa_ref as lr

SR = 44100


def test_oracle_mel_structure():
    M = lr.mel(SR, 2048)
    assert M.shape == (128, 1025) and M.dtype == np.float32
    assert (M >= 0).all() and (M.sum(axis=1) > 0).all()
    assert ((M > 0).sum(axis=0) <= 2).all()                 # each bin feeds at most two neighbouring triangles
    assert M[:, -1].max() < 1e-12                           # Nyquist weight vanishes for fmax = sr / 2
    mel_f = lr.mel_frequencies(130, 0.0, SR / 2)
    assert mel_f[0] == 0.0 and abs(mel_f[-1] - SR / 2) < 1e-6
    assert np.allclose(np.diff(lr.hz_to_mel(mel_f)), np.diff(lr.hz_to_mel(mel_f))[0])   # uniform on the mel axis
    # Slaney normalisation: every triangle has unit area in Hz
    df = SR / 2048
    assert np.allclose(M.sum(axis=1) * df, 1.0, atol=0.15) and np.allclose(M[64:].sum(axis=1) * df, 1.0, atol=1e-2)


def test_oracle_onset_strength_properties():
    rng = np.random.default_rng(5)
    y = (rng.standard_normal(SR) * 0.05).astype(np.float32)
    y[SR // 2:SR // 2 + 2000] += np.sin(2 * np.pi * 880 * np.arange(2000) / SR).astype(np.float32)
    env = lr.onset_strength(y, SR)
    assert env.shape == (1 + len(y) // 512,) and env.dtype == np.float32
    assert (env[:3] == 0).all() and (env >= 0).all()        # lag + centre shift = 3 leading zeros
    peak = int(np.argmax(env))
    assert abs(peak - (SR // 2) // 512) <= 3                # the burst shows up at its frame (+ the centre shift of 2)
    assert lr.onset_strength(np.zeros(4096, np.float32), SR).max() == 0.0
    med = lr.onset_strength(y, SR, aggregate=np.median)
    assert med.shape == env.shape and med[peak] > 0


def test_product_mel_filterbank_matches_oracle():
    from audio_tabs_b200.filters import SlaneyMelFilterbank
    for sr, n_fft, n_mels in [(SR, 2048, 128), (22050, 2048, 128), (SR, 4096, 64)]:
        fb = SlaneyMelFilterbank(sr, n_fft, n_mels=n_mels)
        want = lr.mel(sr, n_fft, n_mels=n_mels)
        assert fb.shape == (n_fft // 2, n_mels)
        assert np.array_equal(np.asarray(fb), want[:, :-1].T)        # bit-identical on the bins the FFT produces
        start, length, woff, w = fb.banded()
        assert (length > 0).all() and len(w) == length.sum()


def test_onset_api_rejects_unsupported_options():
    from audio_tabs_b200 import onsets
    with pytest.raises(ValueError):
        onsets.onset_strength(y=np.zeros(10, np.float32), sr=SR, max_size=3)
    with pytest.raises(ValueError):
        onsets.onset_strength(S=np.zeros((128, 4), np.float32), sr=SR)
    with pytest.raises(ValueError):
        onsets._aggregate_code(np.max)


# ---- GPU parity --------------------------------------------------------------------------------------------
@pytest.mark.gpu
def test_mel_db_matches_oracle(cuda_device):
    from audio_tabs_b200.onsets import OnsetStrength
    from audio_tabs_b200.synth import synth_guitar
    y = synth_guitar(4100, 3.0)
    eng = OnsetStrength(sr=SR)
    packed = eng.fe.pack([y])
    got = eng.mel_db(packed).cpu().numpy()
    want = lr.power_to_db(lr.melspectrogram(y, sr=SR), top_db=None)
    assert got.shape == want.shape == (1 + len(y) // 512, 128)
    floor = want.max() - 80.0
    err = np.abs(np.maximum(got, floor) - np.maximum(want, floor))
    assert err.max() < 0.05, err.max()
    loud = want > want.max() - 40.0                              # well above the noise floor: float32-tight
    assert np.abs(got - want)[loud].max() < 1e-3


@pytest.mark.gpu
def test_mel_db_error_is_the_float32_fft_floor(cuda_device):
    """Why the dB tolerances of this file are looser than the 1e-4 of the madmom path: the
    oracle transforms in float64, and ANY float32 FFT leaves a noise floor of ~1e-7 of the frame's energy in
    every bin; in a mel band 60-80 dB below the clip's maximum that floor is a visible fraction of the band, and
    10 log10 turns it into 1e-2 dB.  Shown with a yardstick that shares no code with the kernel: the same
    spectrogram through a float32 cuFFT (torch.fft, test-only) misses the float64 oracle by as much as we do."""
    import torch
    from audio_tabs_b200.onsets import OnsetStrength, hann_periodic
    from audio_tabs_b200.synth import synth_guitar
    y = synth_guitar(4100, 3.0)
    eng = OnsetStrength(sr=SR)
    ours = eng.mel_db(eng.fe.pack([y])).cpu().numpy().astype(np.float64)
    want = lr.power_to_db(lr.melspectrogram(y, sr=SR), top_db=None).astype(np.float64)
    # yardstick: float32 frames x float32 periodic Hann -> cuFFT float32 -> |X|^2 -> the oracle's mel basis -> dB
    yp = torch.from_numpy(np.pad(y, 1024)).cuda()
    frames = yp.unfold(0, 2048, 512) * torch.from_numpy(hann_periodic(2048).astype(np.float32)).cuda()
    S32 = (torch.fft.rfft(frames, dim=1).abs() ** 2).cpu().numpy()
    basis = lr.mel(sr=SR, n_fft=2048)
    yard = lr.power_to_db(np.einsum("tf,mf->tm", S32, basis).astype(np.float32), top_db=None).astype(np.float64)
    assert yard.shape == want.shape == ours.shape
    top = want.max()
    for lo, hi in ((-40.0, 0.0), (-60.0, -40.0), (-80.0, -60.0)):                 # dB below the clip maximum
        sel = (want > top + lo) & (want <= top + hi)
        if sel.sum() < 50:
            continue
        e_ours, e_yard = np.abs(ours - want)[sel], np.abs(yard - want)[sel]
        # the same order of magnitude as another float32 FFT, at the maximum and at the 99th percentile
        assert e_ours.max() <= 4.0 * e_yard.max() + 1e-4, (lo, hi, e_ours.max(), e_yard.max())
        assert np.percentile(e_ours, 99) <= 4.0 * np.percentile(e_yard, 99) + 1e-4, (lo, hi)
    loud = want > top - 40.0
    assert np.abs(ours - want)[loud].max() < 1e-3                                  # float32-tight where it can be


@pytest.mark.gpu
@pytest.mark.parametrize("aggregate", [None, np.median])
def test_onset_strength_matches_oracle(cuda_device, aggregate):
    from audio_tabs_b200 import onsets
    from audio_tabs_b200.synth import synth_guitar
    for seed, seconds in [(4200, 4.0), (4201, 1.3)]:
        y = synth_guitar(seed, seconds)
        want = lr.onset_strength(y, SR, aggregate=aggregate)
        got = onsets.onset_strength(y=y, sr=SR, aggregate=aggregate)
        assert got.shape == want.shape and got.dtype == np.float32
        assert (got[:3] == 0).all()
        assert np.abs(got - want).max() <= 5e-3 + 1e-3 * np.abs(want).max(), np.abs(got - want).max()


@pytest.mark.gpu
def test_onset_strength_batch_ragged(cuda_device):
    """Clips of different length in one packed batch: no leakage of the lagged row or the top_db maximum
    across clip boundaries; a silent clip gives an all-zero envelope."""
    from audio_tabs_b200.onsets import OnsetStrength
    from audio_tabs_b200.synth import synth_guitar
    clips = [synth_guitar(4300, 2.0), np.zeros(30000, np.float32), synth_guitar(4301, 0.7) * 0.01,
             synth_guitar(4302, 0.02)]
    eng = OnsetStrength(sr=SR, aggregate=np.median)
    outs = eng.process_batch(clips)
    for c, got in zip(clips, outs):
        want = lr.onset_strength(c, SR, aggregate=np.median)
        assert got.shape == want.shape
        assert np.abs(got - want).max() <= 5e-3 + 1e-3 * np.abs(want).max()
    assert outs[1].max() == 0.0


@pytest.mark.gpu
@pytest.mark.parametrize("B", [1, 7, 32, 40, 64, 91, 128, 200, 256, 300])
@pytest.mark.parametrize("aggregate", [0, 1])
def test_onset_envelope_kernel_any_width(cuda_device, B, aggregate):
    """b200spec_onset_envelope on random rows: the register (bitonic sort, 1-8 values per lane) and the
    shared-memory (rank counting, B > 256) medians, means, top_db clip, lag and shift against numpy."""
    import torch
    from audio_tabs_b200 import _ffi
    rng = np.random.default_rng(100 + B)
    lens = [37, 0, 5, 64]
    lag, shift, top_db = 2, 3, 30.0
    L = (rng.standard_normal((sum(lens), B)) * 20).astype(np.float32)
    L[rng.random(L.shape) < 0.1] = 0.0                                  # ties
    off = np.concatenate(([0], np.cumsum(lens))).astype(np.int64)
    dL, doff = torch.from_numpy(L).cuda(), torch.from_numpy(off).cuda()
    env = torch.full((sum(lens),), -1.0, device="cuda")
    scratch = torch.empty(len(lens), device="cuda")
    _ffi.check(_ffi.lib().b200spec_onset_envelope(dL.data_ptr(), B, B, doff.data_ptr(), len(lens), sum(lens), lag, top_db,
                                                  aggregate, shift, scratch.data_ptr(), env.data_ptr(), None))
    got = env.cpu().numpy()
    agg = np.median if aggregate else np.mean
    for c, n in enumerate(lens):
        S = L[off[c]:off[c + 1]]
        want = np.zeros(n, np.float32)
        if n:
            S = np.maximum(S, S.max() - top_db)
            d = agg(np.maximum(0.0, S[lag:] - S[:-lag]), axis=1) if n > lag else np.zeros(0, np.float32)
            full = np.concatenate((np.zeros(lag + shift, np.float32), d))[:n]
            want[:len(full)] = full
        assert np.allclose(got[off[c]:off[c + 1]], want, rtol=1e-5, atol=1e-5), (B, aggregate, c)
