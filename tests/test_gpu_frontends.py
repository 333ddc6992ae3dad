"""GPU parity for every front end the reference reaches, the input formats, the batch engine and the
edge cases (ragged / empty / tiny clips), all through the C ABI, against the oracle and the
committed golden fixtures.  Tolerances: helpers.RTOL / ATOL (1e-4 / 1e-5)."""
from pathlib import Path

import numpy as np
import pytest

from helpers import assert_close, assert_stft_close, noise
from oracle import madmom_ref as ref

pytestmark = pytest.mark.gpu
GOLD = Path(__file__).resolve().parent / "golden"
SR = 44100


@pytest.fixture(scope="module")
def b2(cuda_device):
    import audio_tabs_b200
    return audio_tabs_b200


# ---- golden fixtures ---------------------------------------------------------------------------
def test_golden_beat_onset_logfilt(b2):
    from audio_tabs_b200.frontends import rnn_beat_frontend_fused, rnn_onset_frontend
    x = np.load(GOLD / "guitar_2s_f32.npy")
    assert_close(rnn_beat_frontend_fused()(x), np.load(GOLD / "guitar_2s_beat314.npy"), what="golden beat")
    assert_close(rnn_onset_frontend()(x), np.load(GOLD / "guitar_2s_onset266.npy"), what="golden onset")
    chain = b2.SequentialProcessor((b2.SignalProcessor(num_channels=1, sample_rate=SR), b2.FramedSignalProcessor(),
                                    b2.ShortTimeFourierTransformProcessor(),
                                    b2.LogarithmicFilteredSpectrogramProcessor(num_bands=12, fmin=30, fmax=17000)))
    assert_close(np.asarray(chain(x)), np.load(GOLD / "guitar_2s_logfilt81.npy"), what="golden logfilt")


def test_golden_stft(b2):
    x = np.load(GOLD / "guitar_2s_f32.npy")[:22050]
    got = np.asarray(b2.ShortTimeFourierTransformProcessor()(b2.FramedSignal(b2.Signal(x, sample_rate=SR))))
    assert_stft_close(got, np.load(GOLD / "guitar_0p5s_stft2048.npy"))


def test_golden_real_audio_int16(b2):
    """The reference's own recorded clip (int16): chroma/key front ends at 8192 and the beat front end."""
    from audio_tabs_b200.audio.chroma import cnn_key_frontend, deep_chroma_frontend
    from audio_tabs_b200.frontends import MultiResolutionFrontEnd, beat_specs
    clip = np.load(GOLD / "refjob_3s_i16.npy")
    assert_close(np.asarray(deep_chroma_frontend()(clip)), np.load(GOLD / "refjob_3s_deepchroma105.npy"), what="deepchroma")
    assert_close(np.asarray(cnn_key_frontend()(clip)), np.load(GOLD / "refjob_3s_key105.npy"), what="key")
    got = MultiResolutionFrontEnd(beat_specs(int16=True))(b2.Signal(clip, sample_rate=SR))
    assert_close(got, np.load(GOLD / "refjob_3s_beat314.npy"), what="beat int16")
    n, t = np.load(GOLD / "refjob_total_frames.npy")
    assert len(b2.FramedSignal(b2.Signal(np.zeros(n, np.int16), sample_rate=SR))) == t == 1532


# ---- chroma / key / chord front ends -------------------------------------------
@pytest.mark.parametrize("which,width", [("deep_chroma_frontend", 105), ("cnn_key_frontend", 105),
                                         ("cnn_chord_frontend", 113)])
def test_8192_frontends_float(b2, which, width):
    from audio_tabs_b200.audio import chroma
    from audio_tabs_b200.synth import synth_guitar
    x = synth_guitar(3000 + width, 8.0)
    fps = 5 if which == "cnn_key_frontend" else 10
    fmin, fmax = (60.0, 2600.0) if width == 113 else (65.0, 2100.0)
    want = ref.log_filt_chain(8192, fps=fps, fmin=fmin, fmax=fmax)(x).data
    got = np.asarray(getattr(chroma, which)()(x))
    assert got.shape == want.shape == (8 * fps, width)
    assert_close(got, want, what=which)


@pytest.mark.parametrize("hop,fps", [(441.0, None), (None, 10)])
def test_chord_chroma_config3(b2, hop, fps):
    """BASELINE config 3: frame 4096 -> log filterbank (24 bpo, 65-2100 Hz, 87 bands) -> 12-bin chroma."""
    from audio_tabs_b200.audio.chroma import chord_chroma_frontend
    from audio_tabs_b200.synth import synth_guitar
    x = synth_guitar(3003, 5.0)
    kw = dict(hop_size=hop) if hop else dict(fps=fps)
    spec = ref.log_filt_chain(4096, **({"hop_size": hop} if hop else {"fps": fps}))(x)
    assert spec.data.shape[1] == 87
    want = ref.fold_chroma(spec.data, spec.bin_frequencies)
    got = np.asarray(chord_chroma_frontend(4096, **kw)(x))
    assert got.shape == want.shape and got.shape[1] == 12
    assert_close(got, want, what="folded chroma")


def test_pitch_class_profile(b2):
    x = noise(5, SR * 2)
    stft = b2.ShortTimeFourierTransform(b2.FramedSignal(b2.Signal(x, sample_rate=SR), frame_size=4096))
    got = np.asarray(b2.PitchClassProfile(b2.Spectrogram(stft)))
    rs = ref.spectrogram(ref.ShortTimeFourierTransform(ref.FramedSignal(ref.Signal(x, sample_rate=SR), frame_size=4096)))
    want, _ = ref.pitch_class_profile(rs)
    assert_close(got, want.astype(np.float32), rtol=2e-4, atol=1e-4, what="pcp")


# ---- unfused stages and stand-alone kernels ----------------------------------------------------
def test_spectrogram_and_log_without_filterbank(b2):
    x = noise(6, 30000)
    fr = b2.FramedSignal(b2.Signal(x, sample_rate=SR), frame_size=1024)
    rs = ref.spectrogram(ref.ShortTimeFourierTransform(ref.FramedSignal(ref.Signal(x, sample_rate=SR), frame_size=1024)))
    got = np.asarray(b2.Spectrogram(b2.ShortTimeFourierTransform(fr)))
    assert_close(got, rs.data, rtol=1e-4, atol=1e-5, what="magnitude")
    got_log = np.asarray(b2.LogarithmicSpectrogram(b2.Spectrogram(b2.ShortTimeFourierTransform(fr)), mul=2, add=1))
    assert_close(got_log, ref.logarithmic_spectrogram(rs, mul=2, add=1).data, what="log spec")
    got_diff = np.asarray(b2.SpectrogramDifferenceProcessor(diff_frames=2, positive_diffs=True)(
        b2.LogarithmicSpectrogram(b2.Spectrogram(b2.ShortTimeFourierTransform(fr)), mul=2, add=1)))
    want_diff = ref.spectrogram_difference(ref.logarithmic_spectrogram(rs, mul=2, add=1).data, 2, positive_diffs=True)
    assert_close(got_diff, want_diff, what="diff on raw bins")


def test_standalone_stages_on_host_matrix(b2):
    """A chain rooted at a caller-supplied magnitude matrix runs K2/K3 stand-alone."""
    x = noise(8, 40000)
    rs = ref.spectrogram(ref.ShortTimeFourierTransform(ref.FramedSignal(ref.Signal(x, sample_rate=SR), frame_size=2048)))
    spec = b2.Spectrogram(rs.data, bin_frequencies=rs.bin_frequencies)
    filt = b2.FilteredSpectrogram(spec, num_bands=12, fmin=30, fmax=17000)
    rf = ref.filtered_spectrogram(rs, num_bands=12, fmin=30, fmax=17000)
    assert_close(np.asarray(filt), rf.data, what="stand-alone filter")
    log = b2.LogarithmicSpectrogram(b2.FilteredSpectrogram(spec, num_bands=12, fmin=30, fmax=17000), mul=1, add=1)
    rl = ref.logarithmic_spectrogram(rf)
    assert_close(np.asarray(log), rl.data, what="stand-alone filter+log")
    d = b2.SpectrogramDifference(b2.LogarithmicSpectrogram(b2.FilteredSpectrogram(spec)), diff_frames=1,
                                 positive_diffs=True)
    assert_close(np.asarray(d), ref.spectrogram_difference(rl.data, 1, positive_diffs=True), what="stand-alone diff")


def test_diff_only_and_unstacked(b2):
    x = noise(9, SR)
    mk = lambda stack: b2.SequentialProcessor((  # noqa: E731
        b2.FramedSignalProcessor(frame_size=2048), b2.ShortTimeFourierTransformProcessor(),
        b2.FilteredSpectrogramProcessor(num_bands=6), b2.LogarithmicSpectrogramProcessor(),
        b2.SpectrogramDifferenceProcessor(diff_ratio=0.5, positive_diffs=False, stack_diffs=stack)))
    sig = b2.Signal(x, sample_rate=SR)
    spec = ref.log_filtered_spectrogram(x, frame_size=2048, num_bands=6)
    want = ref.spectrogram_difference(spec, 1, positive_diffs=False)
    assert_close(np.asarray(mk(None)(sig)), want, what="signed diff")
    assert_close(np.asarray(mk(np.hstack)(sig)), np.hstack((spec, want)), what="hstack")
    assert_close(np.asarray(mk(np.vstack)(sig)), np.vstack((spec, want)), what="custom stack fn")


@pytest.mark.parametrize("max_bins,positive,stack", [(None, True, np.hstack), (3, True, None), (None, False, np.hstack),
                                                      (None, True, np.vstack)])
def test_online_difference_processor_continues_from_previous_rows(b2, max_bins, positive, stack):
    """SpectrogramDifferenceProcessor(..., reset=False): madmom's BufferProcessor semantics -- the buffer is as
    long as the first call's rows + diff_frames, later blocks are shifted in and differenced against the rows
    of the calls before; a block longer than the buffer does not fit (numpy's broadcast error)."""
    from audio_tabs_b200.synth import synth_guitar
    blocks = [synth_guitar(4100 + i, sec) for i, sec in enumerate((0.5, 0.2, 0.01, 0.5, 0.33))]

    def front(m):
        return m.SequentialProcessor((
            m.SignalProcessor(num_channels=1, sample_rate=SR), m.FramedSignalProcessor(frame_size=2048, fps=100),
            m.ShortTimeFourierTransformProcessor(), m.FilteredSpectrogramProcessor(num_bands=12, fmin=30, fmax=17000),
            m.LogarithmicSpectrogramProcessor(mul=1, add=1)))

    def values(y):
        return np.asarray(y.data if hasattr(y, "data") and not isinstance(y, np.ndarray) else y)

    kw = dict(diff_ratio=0.5, diff_max_bins=max_bins, positive_diffs=positive, stack_diffs=stack)
    ours, theirs = b2.SpectrogramDifferenceProcessor(**kw), ref.SpectrogramDifferenceProcessor(**kw)
    fo, fr = front(b2), front(ref)
    for i, x in enumerate(blocks):
        reset = i == 0
        got, want = values(ours(fo(x), reset=reset)), values(theirs(fr(x), reset=reset))
        assert got.shape == want.shape, (i, got.shape, want.shape)
        assert np.isfinite(got).all()
        assert_close(got, want, what="online block %d" % i)
    # a first call with reset=False starts like an offline call
    fresh = b2.SpectrogramDifferenceProcessor(**kw)
    assert_close(values(fresh(fo(blocks[1]), reset=False)),
                 values(ref.SpectrogramDifferenceProcessor(**kw)(fr(blocks[1]), reset=True)), what="first online call")
    # ... and its buffer is only as long as that first block (+ diff_frames): a longer block is refused
    with pytest.raises(ValueError, match="could not broadcast"):
        fresh(fo(blocks[0]), reset=False)
    with pytest.raises(ValueError, match="could not broadcast"):
        short = ref.SpectrogramDifferenceProcessor(**kw)
        short(fr(blocks[1]), reset=False)
        short(fr(blocks[0]), reset=False)


# ---- batch engine: formats, ragged batches, flux -----------------------------------------------
def _oracle_beat(x):
    return ref.rnn_beat_preprocessor()(x)


def test_ragged_batch_with_empty_and_tiny_clips(b2):
    from audio_tabs_b200.frontends import beat_specs
    from audio_tabs_b200.plan import FrontEnd
    from audio_tabs_b200.synth import synth_guitar
    clips = [synth_guitar(4000, 1.3), np.zeros(0, np.float32), noise(1, 100), synth_guitar(4001, 0.77),
             noise(2, 441), noise(3, 442), synth_guitar(4002, 2.05), noise(4, 1)]
    fe = FrontEnd(beat_specs(), device=0)
    outs = fe.process_batch(clips)
    assert [o.shape[0] for o in outs] == [ref.num_frames_for(len(c), 441.0) for c in clips]
    for c, o in zip(clips, outs):
        assert o.shape[1] == 314
        if len(c):
            assert_close(o, _oracle_beat(c), what="ragged clip of %d samples" % len(c))
    again = fe.process_batch(clips)
    for a, b in zip(outs, again):
        np.testing.assert_array_equal(a, b)                       # deterministic


def test_stereo_downmix_float_and_int16(b2):
    from audio_tabs_b200.frontends import beat_specs
    from audio_tabs_b200.plan import FrontEnd
    rng = np.random.default_rng(11)
    st = (rng.standard_normal((30000, 2)) * 0.2).astype(np.float32)
    out = FrontEnd(beat_specs(), device=0, channels=2).process_batch([st, st[:9000]])
    assert_close(out[0], _oracle_beat(ref.Signal(st, sample_rate=SR, num_channels=1).data), what="stereo f32")
    assert_close(out[1], _oracle_beat(ref.Signal(st[:9000], sample_rate=SR, num_channels=1).data), what="stereo f32 b")
    sti = rng.integers(-20000, 20000, size=(25000, 2)).astype(np.int16)
    outi = FrontEnd(beat_specs(int16=True), device=0, dtype="i16", channels=2).process_batch([sti])
    assert_close(outi[0], _oracle_beat(ref.Signal(sti, sample_rate=SR, num_channels=1).data), what="stereo i16")
    mono = FrontEnd(beat_specs(int16=True), device=0, dtype="i16").process_batch([sti[:, 0].copy()])
    assert_close(mono[0], _oracle_beat(sti[:, 0].copy()), what="mono i16")


def test_spectral_flux_and_projection_outputs(b2):
    import torch
    from audio_tabs_b200.frontends import log_filt_spec
    from audio_tabs_b200.plan import FrontEnd
    x = noise(12, SR * 2)
    spec = log_filt_spec(2048, 441.0, 12, diff_ratio=0.5, fold=True)
    fe = FrontEnd([spec], device=0)
    packed = fe.pack([x, x[:20000]])
    flux = torch.zeros(packed.total_frames, device="cuda")
    proj = torch.zeros((packed.total_frames, 12), device="cuda")
    out = fe.run_packed(packed, flux=[flux], proj=[proj]).cpu().numpy()
    L = ref.log_filtered_spectrogram(x)
    D = ref.spectrogram_difference(L, 1, positive_diffs=True)
    T = L.shape[0]
    assert_close(out[:T], np.hstack((L, D)), what="stacked")
    np.testing.assert_allclose(flux.cpu().numpy()[:T], ref.spectral_flux(D), rtol=1e-4, atol=1e-4)
    cf = ref.LogarithmicFilterbank(ref.fft_frequencies(1024, SR)).center_frequencies
    assert_close(proj.cpu().numpy()[:T], ref.fold_chroma(L, cf), rtol=1e-4, atol=1e-4, what="projection")
    L2 = ref.log_filtered_spectrogram(x[:20000])
    D2 = ref.spectrogram_difference(L2, 1, positive_diffs=True)
    assert_close(out[T:], np.hstack((L2, D2)), what="second clip restarts the diff")


def test_pinned_pipeline_equals_resident_path(b2):
    import torch
    from audio_tabs_b200.frontends import beat_specs
    from audio_tabs_b200.plan import FrontEnd
    fe = FrontEnd(beat_specs(), device=0)
    lens = [30000, 44100, 12345, 50000, 8000, 61000, 44100]
    xs = [noise(100 + i, n) for i, n in enumerate(lens)]
    packed = fe.pack(xs)
    want = fe.run_packed(packed).cpu()
    host_in = torch.from_numpy(np.concatenate(xs)).pin_memory()
    host_out = torch.empty(want.shape, dtype=torch.float32).pin_memory()
    # uniform groups, the default schedule (one clip, then pairs), an explicit schedule, three slots
    for kw in (dict(group_clips=2), dict(), dict(group_clips=[3, 1, 2]), dict(group_clips=[1, 2], n_slots=3),
               dict(group_clips=16)):
        host_out.zero_()
        fe.process_batch_pinned(host_in, lens, host_out, **kw)
        torch.cuda.synchronize()
        assert torch.equal(host_out, want), kw
    with pytest.raises(ValueError):
        fe.process_batch_pinned(host_in, lens, host_out, group_clips=[2, 0])


def test_device_resident_signal_and_tensor_output(b2):
    import torch
    x = noise(13, 50000)
    t = torch.from_numpy(x).cuda()
    chain = b2.SequentialProcessor((b2.SignalProcessor(num_channels=1, sample_rate=SR), b2.FramedSignalProcessor(),
                                    b2.ShortTimeFourierTransformProcessor(),
                                    b2.LogarithmicFilteredSpectrogramProcessor()))
    stage = chain(t)
    dev = stage.tensor()
    assert dev.is_cuda and dev.shape == (114, 81)
    assert_close(dev.cpu().numpy(), ref.log_filtered_spectrogram(x), what="device signal")


# ---- size-independent properties at BASELINE sizes ---------------------------------------------
def test_properties_at_full_clip_length(b2):
    """3-minute stems (BASELINE config 2 clip length): shapes, exact zeros, sign, determinism, and
    agreement of the fused 3-resolution buffer with single-resolution launches (a checksum of checksums)."""
    import torch
    from audio_tabs_b200.frontends import beat_specs
    from audio_tabs_b200.plan import FrontEnd, Packed
    from audio_tabs_b200.synth import synth_batch_device
    n = 180 * SR
    dev = torch.device("cuda", 0)
    sig = synth_batch_device(4, n, seed=5, device=dev)
    sig[2 * n:3 * n] = 0                                            # one silent stem
    specs = beat_specs()
    fe = FrontEnd(specs, device=0)
    packed = Packed(sig, [n] * 4, 441.0)
    assert packed.num_frames == [18000] * 4
    out = fe.run_packed(packed)
    assert out.shape == (72000, 314) and bool(torch.isfinite(out).all())
    assert not out[36000:54000].any()                              # log10(1 + 0) = 0, diff = 0
    col = 0
    for s in specs:
        b, k = s.num_bands, s.diff_frames
        d = out[:, col + b:col + 2 * b]
        assert bool((d >= 0).all()) and bool((out[:, col:col + b] >= 0).all())
        for c in range(4):
            assert not d[c * 18000:c * 18000 + k].any()            # first k rows of every clip
        single = FrontEnd([s], device=0).run_packed(packed)
        assert torch.equal(single, out[:, col:col + 2 * b])
        col += 2 * b
    assert torch.equal(out, fe.run_packed(packed))                 # idempotent / deterministic
    # linearity of the STFT under scaling by a power of two is exact
    st1 = FrontEnd([specs[1]], device=0).stft_packed(Packed(sig[:n], [n], 441.0))
    st2 = FrontEnd([specs[1]], device=0).stft_packed(Packed(sig[:n] * 0.5, [n], 441.0))
    assert torch.equal(st1 * 0.5, st2)
    # a spot check of the long clip against the oracle (first and last second)
    x = sig[:n].cpu().numpy()
    want_head = ref.rnn_beat_preprocessor()(x[:SR + 4096])[:90]
    assert_close(out[:90].cpu().numpy(), want_head, what="head of 3-min stem")


# ---- SuperFlux: SpectrogramDifference(diff_max_bins=3) -------------------------
@pytest.mark.parametrize("max_bins,stack", [(3, True), (2, False), (5, True)])
def test_superflux_difference(b2, max_bins, stack):
    """madmom's SuperFlux chain: frame 2048 @ 200 fps, 24 bands/oct, log10(1+x), lagged row widened by
    a maximum filter over `max_bins` bands (scipy.ndimage.maximum_filter, 'reflect') before subtraction."""
    from audio_tabs_b200.synth import synth_guitar
    x = synth_guitar(3100 + max_bins, 3.0)

    def chain(m, max_bins=max_bins, stack=stack):
        return m.SequentialProcessor((
            m.SignalProcessor(num_channels=1, sample_rate=SR), m.FramedSignalProcessor(frame_size=2048, fps=200),
            m.ShortTimeFourierTransformProcessor(),
            m.FilteredSpectrogramProcessor(num_bands=24, fmin=30, fmax=17000, norm_filters=False),
            m.LogarithmicSpectrogramProcessor(mul=1, add=1),
            m.SpectrogramDifferenceProcessor(diff_ratio=0.5, diff_max_bins=max_bins, positive_diffs=True,
                                             stack_diffs=np.hstack if stack else None)))
    want = chain(ref)(x)
    want = np.asarray(want.data if hasattr(want, "data") and not isinstance(want, np.ndarray) else want)
    got = np.asarray(chain(b2)(x))
    assert got.shape == want.shape and got.shape[0] == 600
    assert_close(got, want, what="superflux M=%d" % max_bins)
    # the maximum filter only ever lowers the positive difference
    plain = np.asarray(chain(b2, None, False)(x))
    sf = got[:, got.shape[1] // 2:] if stack else got
    assert (sf <= plain + 1e-6).all() and (sf < plain - 1e-4).any()


def test_superflux_flux_only_batch(b2):
    """FrontEnd.run_packed with diff_max_bins > 1 and out=False: flux rows equal the row sums of the difference."""
    import torch
    from audio_tabs_b200.frontends import log_filt_spec
    from audio_tabs_b200.plan import FrontEnd
    from audio_tabs_b200.synth import synth_guitar
    spec = log_filt_spec(2048, 441.0, 12, diff_ratio=0.5)
    spec_sf = log_filt_spec(2048, 441.0, 12, diff_ratio=0.5, diff_max_bins=3)
    clips = [synth_guitar(3200 + i, 1.0 + 0.37 * i) for i in range(3)]
    fe = FrontEnd([spec_sf], device=0)
    packed = fe.pack(clips)
    full = fe.run_packed(packed)
    flux = torch.empty(packed.total_frames, dtype=torch.float32, device="cuda")
    fe.run_packed(packed, out=False, flux=[flux])
    B = spec.num_bands
    assert_close(flux.cpu().numpy(), full[:, B:].sum(dim=1).cpu().numpy(), rtol=1e-5, atol=1e-5, what="superflux flux")
    o = 0
    for c in clips:      # per clip against the oracle, rows must not leak across clip boundaries
        t = ref.num_frames_for(len(c), 441.0)
        L = ref.log_filtered_spectrogram(c, frame_size=2048, num_bands=12)
        L = np.asarray(L.data if hasattr(L, "data") else L)
        D = ref.spectrogram_difference(L, spec.diff_frames, diff_max_bins=3, positive_diffs=True)
        assert_close(full[o:o + t, B:].cpu().numpy(), D.astype(np.float32), what="superflux batch diff")
        o += t


@pytest.mark.parametrize("frame_size", [1024, 2048, 4096, 8192])
@pytest.mark.parametrize("lag", [1, 3, 7, 16])
def test_difference_across_task_seams(b2, frame_size, lag):
    """The fused kernels cut a clip into tasks and run them without warm-up rows; the first `lag` rows of every task
    are rewritten by the seam kernel from the filtered rows.  Small batches get 2-frame tasks, so with lags up to
    B200SPEC_MAX_DIFF_FRAMES every row is a seam row here; the flux must agree whether it comes with the stacked
    matrix (seam path) or alone (warm-up rows: there is no matrix to read the lagged rows from)."""
    import dataclasses
    import torch
    from audio_tabs_b200.frontends import log_filt_spec
    from audio_tabs_b200.plan import FrontEnd
    from audio_tabs_b200.synth import synth_guitar
    spec = dataclasses.replace(log_filt_spec(frame_size, 441.0, 12, diff_ratio=0.5), diff_frames=lag, positive_diffs=True)
    clips = [synth_guitar(3300 + 7 * i + lag, sec) for i, sec in enumerate((0.9, 0.011, 0.35, 2.2))]
    fe = FrontEnd([spec], device=0)
    packed = fe.pack(clips)
    flux_with = torch.empty(packed.total_frames, dtype=torch.float32, device="cuda")
    full = fe.run_packed(packed, flux=[flux_with])
    flux_alone = torch.full((packed.total_frames,), -1.0, dtype=torch.float32, device="cuda")
    fe.run_packed(packed, out=False, flux=[flux_alone])
    B = spec.num_bands
    o = 0
    for c in clips:
        t = ref.num_frames_for(len(c), 441.0)
        L = ref.log_filtered_spectrogram(c, frame_size=frame_size, num_bands=12)
        L = np.asarray(L.data if hasattr(L, "data") else L)
        D = ref.spectrogram_difference(L, lag, positive_diffs=True) if t > lag else np.zeros_like(L)
        assert_close(full[o:o + t, :B].cpu().numpy(), L.astype(np.float32), what="spec, lag %d" % lag)
        assert_close(full[o:o + t, B:].cpu().numpy(), D.astype(np.float32), atol=2e-5, what="diff, lag %d" % lag)
        o += t
    want_flux = full[:, B:].sum(dim=1).cpu().numpy()
    assert_close(flux_with.cpu().numpy(), want_flux, rtol=1e-5, atol=1e-5, what="flux next to the matrix")
    assert_close(flux_alone.cpu().numpy(), want_flux, rtol=1e-4, atol=1e-4, what="flux alone")


def test_rows_do_not_depend_on_the_batch(b2):
    """Without warm-up rows every task starts on an even frame, so the two frames that share a complex FFT are
    always (2i, 2i + 1) and the lagged difference is formed from the very values that are stored: a clip's rows are
    bitwise the same whether it is processed alone or inside a large batch (different task sizes, different SMs)."""
    from audio_tabs_b200.frontends import beat_specs
    from audio_tabs_b200.plan import FrontEnd
    from audio_tabs_b200.synth import synth_guitar
    fe = FrontEnd(beat_specs(), device=0)
    clip = synth_guitar(3400, 4.0)
    alone = fe.process_batch([clip])[0]
    crowd = [synth_guitar(3401 + i, 30.0) for i in range(12)]
    batch = fe.process_batch(crowd[:5] + [clip] + crowd[5:])
    assert np.array_equal(batch[5], alone)


def test_large_batch_with_two_task_sizes_keeps_rows_bitwise(b2):
    """A batch large enough for the two-size task plan (long tasks, then short ones from the last clips that fill
    the last round): clips from the long-task region, from the short-task tail and the very last one must have
    exactly the rows they get when processed alone; nothing unwritten, nothing non-finite."""
    import torch
    from audio_tabs_b200.frontends import beat_specs
    from audio_tabs_b200.plan import FrontEnd, Packed
    from audio_tabs_b200.synth import synth_batch_device
    n_clips, n = 40, 120 * SR
    dev = torch.device("cuda", 0)
    sig = synth_batch_device(n_clips, n, seed=77, device=dev)
    fe = FrontEnd(beat_specs(), device=0)
    packed = Packed(sig, [n] * n_clips, fe.hop_size)
    out = torch.full((packed.total_frames, fe.width), float("nan"), dtype=torch.float32, device=dev)
    fe.run_packed(packed, out)
    torch.cuda.synchronize()
    assert bool(torch.isfinite(out).all())
    T = packed.total_frames // n_clips
    for c in (0, 17, 29, 33, n_clips - 1):
        one = Packed(sig[c * n:(c + 1) * n], [n], fe.hop_size)
        alone = fe.run_packed(one)
        torch.cuda.synchronize()
        assert torch.equal(out[c * T:(c + 1) * T], alone), "clip %d" % c


@pytest.mark.parametrize("bands,context", [(105, 15), (7, 3), (33, 4), (1, 5), (128, 1)])
def test_context_stack_flat_stream(b2, bands, context):
    """b200spec_context_stack writes the dense (rows, context*bands) matrix as one flat stream of aligned 128-bit
    stores: ragged clips (empty and one-row ones included), row widths that are no multiple of four floats, an
    output pointer that is not 16-byte aligned, a strided input, sentinels before and after -- bit-exact against
    the per-clip zero-padded stacking (madmom DeepChromaProcessor: FramedSignal(frame_size=15, hop_size=1) + _dcp_flatten)."""
    import ctypes as C
    import torch
    from audio_tabs_b200 import _ffi
    rng = np.random.default_rng(bands * 100 + context)
    lens = [37, 0, 1, 260, 5, 0, 1033, 2]
    T = sum(lens)
    ld = bands + 3
    x = rng.standard_normal((T, ld)).astype(np.float32)
    half, W = context // 2, context * bands
    want = np.zeros((T, W), np.float32)
    o = 0
    for n in lens:
        for t in range(n):
            for c in range(context):
                s_ = t + c - half
                if 0 <= s_ < n:
                    want[o + t, c * bands:(c + 1) * bands] = x[o + s_, :bands]
        o += n
    dev = torch.device("cuda", 0)
    xin = torch.from_numpy(x).to(dev)
    fo = torch.tensor(np.concatenate(([0], np.cumsum(lens))), dtype=torch.int64, device=dev)
    for shift in (0, 1, 3):                                   # float offset of the output inside its allocation
        buf = torch.full((T * W + 64,), -7.0, dtype=torch.float32, device=dev)
        out = buf[16 + shift:16 + shift + T * W]
        _ffi.check(_ffi.lib().b200spec_context_stack(C.c_void_p(xin.data_ptr()), ld, bands, C.c_void_p(fo.data_ptr()),
                                                     len(lens), T, context, C.c_void_p(out.data_ptr()),
                                                     C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)))
        torch.cuda.synchronize()
        got = out.cpu().numpy().reshape(T, W)
        assert np.array_equal(got, want), "shift %d" % shift
        assert bool((buf[:16 + shift] == -7.0).all()) and bool((buf[16 + shift + T * W:] == -7.0).all())


# ---- ingest: per-clip peak + fused peak normalisation ---------------------------
@pytest.mark.parametrize("dtype,channels", [("f32", 1), ("f32", 2), ("i16", 1), ("i16", 2)])
def test_clip_peak_and_fused_normalisation(b2, dtype, channels):
    """b200spec_clip_peak + d_clip_scale == the front end of the peak-normalised clip
    (services/audio.py:24-26 peak_normalize for float input; madmom Signal(norm=True) with eps = 0)."""
    import torch
    from audio_tabs_b200.frontends import beat_specs
    from audio_tabs_b200.plan import FrontEnd
    from audio_tabs_b200.synth import synth_guitar
    rng = np.random.default_rng(77)
    clips = []
    for i in range(3):
        y = synth_guitar(3300 + i, 0.8 + 0.5 * i) * (0.05 + 0.3 * i)      # different peaks per clip
        if channels == 2:
            y = np.stack([y, np.roll(y, 7) * 0.5 + rng.standard_normal(len(y)).astype(np.float32) * 0.01], axis=1)
        if dtype == "i16":
            y = np.clip(np.round(y * 20000), -32768, 32767).astype(np.int16)
        clips.append(np.ascontiguousarray(y))
    fe = FrontEnd(beat_specs(int16=(dtype == "i16")), device=0, dtype=dtype, channels=channels)
    packed = fe.pack(clips)
    eps = 1e-9 if dtype == "f32" else 0.0
    peaks = fe.peak_scales(packed, reciprocal=False).cpu().numpy()
    mono = [ref.remix(c, 1) if channels == 2 else c for c in clips]
    assert np.array_equal(peaks, np.array([np.abs(m.astype(np.float32)).max() for m in mono], np.float32))
    got = fe.run_packed(packed, clip_scale=fe.peak_scales(packed, eps=eps)).cpu().numpy()
    o = 0
    for m in mono:
        x = m.astype(np.float32)
        x = x / (np.abs(x).max() + np.float32(eps))
        want = ref.rnn_beat_preprocessor()(x)
        assert_close(got[o:o + len(want)], want, what="normalised %s/%d" % (dtype, channels))
        o += len(want)
    assert o == got.shape[0]


def test_device_signal_norm_through_processors(b2):
    """SignalProcessor(norm=True) on a CUDA tensor: same result as madmom's host-side normalisation."""
    import torch
    from audio_tabs_b200.synth import synth_guitar
    x = synth_guitar(3400, 2.0) * 0.23

    def chain(m):
        return m.SequentialProcessor((
            m.SignalProcessor(num_channels=1, sample_rate=SR, norm=True), m.FramedSignalProcessor(frame_size=2048, fps=100),
            m.ShortTimeFourierTransformProcessor(), m.LogarithmicFilteredSpectrogramProcessor(num_bands=12)))
    want = chain(ref)(x).data
    got_dev = np.asarray(chain(b2)(torch.from_numpy(x).cuda()))
    got_host = np.asarray(chain(b2)(x))
    assert_close(got_dev, want, what="norm on device tensor")
    assert_close(got_host, want, what="norm on host array")
    from audio_tabs_b200.frontends import log_filt_spec
    from audio_tabs_b200.plan import FrontEnd
    outs = FrontEnd([log_filt_spec(2048, 441.0, 12)], device=0).process_batch([x, x * 3.0], peak_normalize=True, eps=0.0)
    assert_close(outs[0], want, what="process_batch peak_normalize")
    assert_close(outs[1], want, what="process_batch peak_normalize (gain invariance)")


def test_context_stack_on_device(b2):
    """DeepChroma's +-7 frame context stacking (T, 105) -> (T, 1575) as a device kernel, ragged clips."""
    import torch
    from audio_tabs_b200.audio.chroma import context_stack, context_stack_device
    rng = np.random.default_rng(8)
    lens = [40, 3, 0, 17]
    x = rng.standard_normal((sum(lens), 105)).astype(np.float32)
    off = np.concatenate(([0], np.cumsum(lens))).astype(np.int64)
    got = context_stack_device(torch.from_numpy(x).cuda(), 15, torch.from_numpy(off).cuda()).cpu().numpy()
    assert got.shape == (sum(lens), 1575)
    for c, n in enumerate(lens):
        want = ref.dcp_context(x[off[c]:off[c + 1]], 15) if n else np.zeros((0, 1575), np.float32)
        assert np.array_equal(got[off[c]:off[c + 1]], want)
    assert np.array_equal(context_stack(torch.from_numpy(x[:40]).cuda(), 15).cpu().numpy(), context_stack(x[:40], 15))


def test_repeated_runs_are_bitwise_identical(b2):
    """No atomics or ordering-dependent sums on the data path: the same packed batch gives the same bits
    every time (a shared-memory race between the phases of a group would show up here first; the pool
    has compute-sanitizer closed).  64 clips of different length keep all 148 x 4 groups busy and exercise
    the dynamic task queue in different interleavings."""
    import torch
    from audio_tabs_b200.frontends import beat_specs
    from audio_tabs_b200.plan import FrontEnd
    from audio_tabs_b200.synth import synth_batch_device
    dev = torch.device("cuda", 0)
    n = 44100 * 20
    sig = synth_batch_device(64, n, seed=99, device=dev).reshape(-1)
    lens = [n - 1000 * (i % 7) for i in range(64)]
    clips = [sig[i * n:i * n + lens[i]] for i in range(64)]
    fe = FrontEnd(beat_specs(), device=0)
    packed = fe.pack(clips)
    first = fe.run_packed(packed).clone()
    assert torch.isfinite(first).all()
    for _ in range(4):
        assert torch.equal(fe.run_packed(packed), first)
    flux = [torch.empty(packed.total_frames, device=dev) for _ in fe.specs]
    again = fe.run_packed(packed, flux=flux)
    assert torch.equal(again, first)
    B = fe.specs[2].num_bands
    c = fe.col[2]
    assert torch.allclose(flux[2], first[:, c + B:c + 2 * B].sum(dim=1), rtol=1e-5, atol=1e-5)


def test_full_config2_batch_properties(b2):
    """BASELINE configs[1] at its FULL size (64 stems x 180 s -> (1 152 000, 314)): every clip's rows agree
    with running that clip alone (different task sizes / frame pairings: float32 rounding only), the tail
    of the last clip (right zero padding) and the head of the first match the oracle, nothing leaks
    across clip boundaries, all values finite and non-negative."""
    import torch
    from audio_tabs_b200.frontends import beat_specs
    from audio_tabs_b200.plan import FrontEnd, Packed
    from audio_tabs_b200.synth import synth_batch_device
    n, nc = 180 * SR, 64
    dev = torch.device("cuda", 0)
    sig = synth_batch_device(nc, n, seed=2000, device=dev)
    fe = FrontEnd(beat_specs(), device=0)
    packed = Packed(sig, [n] * nc, 441.0)
    out = fe.run_packed(packed)
    assert out.shape == (nc * 18000, 314)
    assert bool(torch.isfinite(out).all()) and bool((out >= 0).all())
    for c in (0, 31, 63):
        alone = fe.run_packed(Packed(sig[c * n:(c + 1) * n], [n], 441.0))
        rows = out[c * 18000:(c + 1) * 18000]
        assert float((rows - alone).abs().max()) <= 2e-6 * max(1.0, float(alone.abs().max()))
    # per-resolution checksums of the whole batch equal the sum of the per-clip checksums (double precision)
    total = out.double().sum(dim=0)
    per_clip = sum(out[c * 18000:(c + 1) * 18000].double().sum(dim=0) for c in range(nc))
    assert torch.allclose(total, per_clip, rtol=1e-12)
    x_last = sig[(nc - 1) * n:].cpu().numpy()
    start = 441 * (18000 - 130)                    # a multiple of the hop: frame j of the excerpt is frame 17870 + j
    want_tail = ref.rnn_beat_preprocessor()(x_last[start:])            # the last 130 frames run past the end of the stem
    assert want_tail.shape[0] == 130
    assert_close(out[-100:].cpu().numpy(), want_tail[-100:], what="tail of the last stem")
    x0 = sig[:n].cpu().numpy()
    assert_close(out[:90].cpu().numpy(), ref.rnn_beat_preprocessor()(x0[:SR + 4096])[:90], what="head of the first stem")


# ---- framing options: origin, end, fractional hop, hop > frame, task-boundary lengths ----------------
@pytest.mark.parametrize("kw", [
    dict(frame_size=2048, hop_size=441.0, origin="right"),          # 'future' frames: origin = -(frame/2)
    dict(frame_size=2048, hop_size=441.0, origin="left"),           # 'past' frames: origin = (frame-1)/2
    dict(frame_size=1024, hop_size=441.0, origin=100),
    dict(frame_size=4096, hop_size=441.0, end="extend"),
    dict(frame_size=2048, fps=123.4),                               # fractional hop 357.37...
    dict(frame_size=1024, hop_size=2500.0),                         # hop > frame: gaps between frames
    dict(frame_size=4096, fps=7),                                   # hop 6300: low-overlap regime
])
def test_framing_options(b2, kw):
    from audio_tabs_b200.synth import synth_guitar
    x = synth_guitar(3500 + len(str(kw)), 2.3)

    def chain(m):
        return m.SequentialProcessor((
            m.SignalProcessor(num_channels=1, sample_rate=SR), m.FramedSignalProcessor(**kw),
            m.ShortTimeFourierTransformProcessor(), m.FilteredSpectrogramProcessor(num_bands=12, fmin=30, fmax=17000),
            m.LogarithmicSpectrogramProcessor(mul=1, add=1),
            m.SpectrogramDifferenceProcessor(diff_ratio=0.5, positive_diffs=True, stack_diffs=np.hstack)))
    want = np.asarray(chain(ref)(x))
    got = np.asarray(chain(b2)(x))
    assert got.shape == want.shape
    assert_close(got, want, what=str(kw))


def test_clip_lengths_around_task_boundaries(b2):
    """Frame counts that straddle the 92-95-frame tasks, the tail batches of 2 / 4 frames and the frame pairs
    (odd counts, one frame, counts just above and below a task), all in one packed batch."""
    from audio_tabs_b200.frontends import beat_specs
    from audio_tabs_b200.plan import FrontEnd
    from audio_tabs_b200.synth import synth_guitar
    base = synth_guitar(3600, 3.0)
    counts = [1, 2, 3, 5, 91, 92, 93, 94, 95, 96, 97, 98, 187, 189, 191, 193]
    clips = [base[:441 * (c - 1) + 17 + i] for i, c in enumerate(counts)]
    fe = FrontEnd(beat_specs(), device=0)
    outs = fe.process_batch(clips)
    for c, clip, got in zip(counts, clips, outs):
        want = ref.rnn_beat_preprocessor()(clip)
        assert got.shape == want.shape == (c, 314)
        assert_close(got, want, what="%d frames" % c)


@pytest.mark.parametrize("window", [np.hamming, np.blackman, None], ids=["hamming", "blackman", "rectangular"])
def test_other_windows_use_the_table_path(b2, window):
    """np.hanning is evaluated in registers by the warp and pair kernels; any other window must come from the table."""
    x = noise(21, 50000) * 0.3
    for frame_size in (1024, 2048, 4096):
        rs = ref.spectrogram(ref.ShortTimeFourierTransform(
            ref.FramedSignal(ref.Signal(x, sample_rate=SR), frame_size=frame_size), window=window))
        want = ref.logarithmic_spectrogram(ref.filtered_spectrogram(rs, num_bands=12), mul=1, add=1).data
        want = np.hstack([want, ref.spectrogram_difference(want, 1, positive_diffs=True)])
        stft = b2.ShortTimeFourierTransform(b2.FramedSignal(b2.Signal(x, sample_rate=SR), frame_size=frame_size),
                                            window=window)
        log = b2.LogarithmicSpectrogram(b2.FilteredSpectrogram(b2.Spectrogram(stft), num_bands=12), mul=1, add=1)
        got = np.asarray(b2.SpectrogramDifferenceProcessor(diff_frames=1, positive_diffs=True, stack_diffs=np.hstack)(log))
        assert got.shape == want.shape
        assert_close(got, want, what="log spec + diff, frame %d" % frame_size)


# ---- round 2: per-clip status, gain on power spectrograms, non-current device ---------------------------
def test_clip_status_and_batch_isolation(b2):
    """A clip with NaN / Inf samples sets ITS status bit and leaves the other clips' rows bitwise unchanged
    ("a failed shard must not poison the batch")."""
    import torch
    from audio_tabs_b200 import _ffi
    from audio_tabs_b200.frontends import beat_specs
    from audio_tabs_b200.plan import FrontEnd
    from audio_tabs_b200.synth import synth_guitar
    clips = [synth_guitar(3500 + i, 1.0 + 0.3 * i) for i in range(4)]
    bad = [c.copy() for c in clips]
    bad[1][12345] = np.nan
    bad[3][777] = np.inf
    fe = FrontEnd(beat_specs(), device=0)
    clean, st0 = fe.process_batch(clips, return_status=True)
    got, st = fe.process_batch(bad, return_status=True)
    assert st0.dtype == np.int32 and st0.tolist() == [0, 0, 0, 0]
    assert st.tolist() == [0, _ffi.CLIP_NONFINITE, 0, _ffi.CLIP_NONFINITE]
    for i in (0, 2):
        assert np.array_equal(got[i], clean[i])
    assert not np.isfinite(got[1]).all() and not np.isfinite(got[3]).all()
    rows = np.nonzero(~np.isfinite(got[1]).all(axis=1))[0]            # only the frames that contain the sample
    assert rows.min() >= (12345 - 2048) // 441 - 2 and rows.max() <= (12345 + 2048) // 441 + 3
    # the single-frame kernel (frame 8192) reports it as well
    from audio_tabs_b200.frontends import log_filt_spec
    fe8 = FrontEnd([log_filt_spec(8192, 4410.0, 24, 65.0, 2100.0)], device=0)
    _, st8 = fe8.process_batch(bad, return_status=True)
    assert st8.tolist() == [0, 1, 0, 1]


def test_peak_normalisation_of_a_power_spectrogram(b2):
    """the per-clip gain on a power (|X|^2) filterbank is gain^2 -- dB mel spectrogram of the
    peak-normalised clip == fused gain on the raw clip."""
    import torch
    from audio_tabs_b200.onsets import mel_db_spec
    from audio_tabs_b200.plan import FrontEnd
    from audio_tabs_b200.synth import synth_guitar
    clips = [synth_guitar(3600 + i, 1.0) * s for i, s in enumerate((0.07, 0.9, 0.31))]
    fe = FrontEnd([mel_db_spec(44100)], device=0, end="extend")
    fused = fe.process_batch(clips, peak_normalize=True, eps=0.0)
    plain = fe.process_batch([c / np.abs(c).max() for c in clips])
    for a, b in zip(fused, plain):
        # 10 log10: a wrong (linear) gain would be off by 10 log10(g) = up to 11.5 dB here; what remains is the float32
        # noise floor of two differently scaled transforms in the quietest mel bands (2e-3 dB measured)
        assert np.abs(a - b).max() < 5e-3, np.abs(a - b).max()


def test_planless_calls_follow_the_pointer_device(b2):
    """context_stack / onset_envelope / magnitude launch on the device that owns their input,
    whatever the current device is (needs two GPUs)."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from audio_tabs_b200.audio.chroma import context_stack_device
    rng = np.random.default_rng(9)
    x = rng.standard_normal((50, 105)).astype(np.float32)
    with torch.cuda.device(0):
        t = torch.from_numpy(x).to("cuda:1")
        got = context_stack_device(t, 15)
        torch.cuda.synchronize(1)
    assert got.device.index == 1 and np.array_equal(got.cpu().numpy(), ref.dcp_context(x, 15))


# ---- round 2: all resolutions in one launch (b200spec_logfilt_multi) --------------------------------------
@pytest.mark.parametrize("dtype,channels", [("f32", 1), ("i16", 1), ("f32", 2)])
def test_one_launch_front_end_matches_oracle_and_per_resolution_launches(b2, dtype, channels):
    """k_front_multi: a group takes each (clip, chunk) task through 1024 / 2048 / 4096 in turn.  Same numbers as
    the oracle and (to a few ulp) as three b200spec_logfilt launches; ragged batch with an empty clip and clips
    shorter than a frame."""
    import torch
    from audio_tabs_b200.frontends import beat_specs
    from audio_tabs_b200.plan import FrontEnd
    from audio_tabs_b200.synth import synth_guitar
    lens = [44100, 0, 517, 30000, 1, 88200 + 13, 4096]
    clips = []
    for i, n in enumerate(lens):
        y = synth_guitar(3700 + i, max(n, 1) / 44100.0)[:n] if n else np.zeros(0, np.float32)
        if channels == 2:
            y = np.stack([y, np.roll(y, 5) * 0.25], axis=1) if n else np.zeros((0, 2), np.float32)
        if dtype == "i16":
            y = np.clip(np.round(y * 30000), -32768, 32767).astype(np.int16)
        clips.append(np.ascontiguousarray(y))
    specs = beat_specs(int16=(dtype == "i16"))
    fe1 = FrontEnd(specs, device=0, dtype=dtype, channels=channels, one_launch=True)
    fe3 = FrontEnd(specs, device=0, dtype=dtype, channels=channels, one_launch=False)
    assert fe1.one_launch and not fe3.one_launch
    packed = fe1.pack(clips)
    flux1 = [torch.empty(packed.total_frames, device="cuda") for _ in specs]
    flux3 = [torch.empty(packed.total_frames, device="cuda") for _ in specs]
    st = torch.zeros(len(clips), dtype=torch.int32, device="cuda")
    got1 = fe1.run_packed(packed, flux=flux1, clip_status=st).cpu().numpy()
    got3 = fe3.run_packed(packed, flux=flux3).cpu().numpy()
    o = 0
    for c in clips:
        mono = ref.remix(c, 1) if channels == 2 else c
        if len(mono) == 0:
            continue
        want = ref.rnn_beat_preprocessor()(mono)
        assert_close(got1[o:o + len(want)], want, what="one launch %s/%d" % (dtype, channels))
        o += len(want)
    assert o == got1.shape[0]
    # Not bitwise: the two paths cut the clips into different (clip, chunk) tasks, so a given frame travels through
    # the pair transform with a different partner frame, and the partner's rounding noise (1e-7 of the larger
    # frame) leaks into it.  Both are within the oracle tolerance; against each other they agree to a few ulp.
    np.testing.assert_allclose(got1, got3, rtol=2e-5, atol=4e-6)
    for a, b in zip(flux1, flux3):
        np.testing.assert_allclose(a.cpu().numpy(), b.cpu().numpy(), rtol=2e-5, atol=2e-5)
    assert st.cpu().tolist() == [0] * len(clips)


def test_one_launch_onset_front_end_and_large_batch(b2):
    """RNNOnsetProcessor's resolutions (6 / 6 / 6 bands per octave, log10(5x + 1), ratio 0.25 -> lags 1 / 2 / 3) in
    one launch, and a batch large enough for several tasks per group."""
    import torch
    from audio_tabs_b200.frontends import beat_specs, onset_specs
    from audio_tabs_b200.plan import FrontEnd
    from audio_tabs_b200.synth import synth_guitar
    x = synth_guitar(4242, 2.0)
    got = FrontEnd(onset_specs(), device=0, one_launch=True).process_batch([x])[0]
    assert_close(got, np.load(GOLD / "guitar_2s_onset266.npy"), what="one launch onset266")
    clips = [synth_guitar(3800 + i, 6.0 + 0.37 * i) for i in range(24)]
    a = FrontEnd(beat_specs(), device=0, one_launch=True).process_batch(clips)
    b = FrontEnd(beat_specs(), device=0, one_launch=False).process_batch(clips)
    for u, v in zip(a, b):
        np.testing.assert_allclose(u, v, rtol=2e-5, atol=4e-6)
    assert_close(a[5], ref.rnn_beat_preprocessor()(clips[5]), what="one launch, clip 5 of 24")


def test_one_launch_rejects_what_it_cannot_run(b2):
    from audio_tabs_b200.frontends import log_filt_spec
    from audio_tabs_b200.plan import FrontEnd
    with pytest.raises(ValueError):
        FrontEnd([log_filt_spec(4096, 441.0, 12), log_filt_spec(8192, 441.0, 24, 65.0, 2100.0)], device=0, one_launch=True)
    with pytest.raises(ValueError):
        FrontEnd([log_filt_spec(2048, 441.0, 12)], device=0, one_launch=True)       # a single resolution: nothing to fuse


# ---- round 2: remaining madmom options -------------------------------------------------------------------
@pytest.mark.parametrize("log", [np.log, np.log2, np.log1p])
def test_other_log_functions(b2, log):
    """LogarithmicSpectrogramProcessor(log=np.log / np.log2 / np.log1p): evaluated on the device as a scaled log10"""
    from audio_tabs_b200.synth import synth_guitar
    x = synth_guitar(3900, 1.0)

    def chain(m, lg):
        return m.SequentialProcessor((
            m.SignalProcessor(num_channels=1, sample_rate=SR), m.FramedSignalProcessor(frame_size=2048, fps=100),
            m.ShortTimeFourierTransformProcessor(), m.FilteredSpectrogramProcessor(num_bands=12),
            m.LogarithmicSpectrogramProcessor(log=lg, mul=2.0, add=1.0)))
    want = chain(ref, log)(x).data
    got = np.asarray(chain(b2, log)(x))
    assert_close(got, want, what="log=%s" % log.__name__)
    with pytest.raises(ValueError):
        chain(b2, np.sqrt)(x)
    # the stand-alone stage (host spectrogram in) takes the same functions
    mag = np.abs(ref.ShortTimeFourierTransform(ref.FramedSignal(ref.Signal(x, sample_rate=SR), frame_size=2048)).data)
    got2 = np.asarray(b2.LogarithmicSpectrogram(b2.Spectrogram(mag), log=log, mul=2.0, add=1.0))
    assert_close(got2, log(np.float32(2.0) * mag + np.float32(1.0)).astype(np.float32), what="stand-alone log=%s" % log.__name__)


def test_fused_multi_resolution_front_end_honours_norm(b2):
    """rnn_beat_frontend_fused() on a DeviceSignal with norm=True applies the per-clip gain, like the
    per-branch chain does"""
    import torch
    from audio_tabs_b200.audio.signal import SignalProcessor
    from audio_tabs_b200.frontends import MultiResolutionFrontEnd, beat_specs
    from audio_tabs_b200.synth import synth_guitar
    x = synth_guitar(3901, 1.5) * 0.17
    sig = SignalProcessor(num_channels=1, sample_rate=SR, norm=True)(torch.from_numpy(x).cuda())
    got = MultiResolutionFrontEnd(beat_specs())(sig)
    assert_close(got, ref.rnn_beat_preprocessor()(x / np.abs(x).max()), what="fused front end, norm=True")
    with pytest.raises(NotImplementedError):
        SignalProcessor(num_channels=1, sample_rate=SR, gain=6.0)(torch.from_numpy(x).cuda())


@pytest.mark.parametrize("frame_size,fft_size", [(2048, 4096), (3000, 4096), (2048, 1024), (1500, 2048), (6000, 8192)])
def test_fft_size_different_from_frame_size(b2, frame_size, fft_size):
    """madmom stft(fft_size=...): the windowed frame is zero-padded (or cut) at its end to fft_size points; any
    frame_size works as long as fft_size is one of the transform lengths"""
    from audio_tabs_b200.synth import synth_guitar
    x = synth_guitar(3950 + fft_size, 0.8)

    def stft_of(m):
        frames = m.FramedSignal(m.Signal(x, sample_rate=SR), frame_size=frame_size, hop_size=441.0)
        return m.ShortTimeFourierTransform(frames, fft_size=fft_size)
    want = stft_of(ref)
    got = stft_of(b2)
    assert got.shape == want.data.shape == (80, fft_size // 2)
    assert np.array_equal(got.bin_frequencies, want.bin_frequencies)
    assert_stft_close(np.asarray(got), want.data)

    def chain(m):
        return m.SequentialProcessor((
            m.SignalProcessor(num_channels=1, sample_rate=SR), m.FramedSignalProcessor(frame_size=frame_size, fps=100),
            m.ShortTimeFourierTransformProcessor(fft_size=fft_size), m.FilteredSpectrogramProcessor(num_bands=12),
            m.LogarithmicSpectrogramProcessor(mul=1, add=1),
            m.SpectrogramDifferenceProcessor(diff_ratio=0.5, positive_diffs=True, stack_diffs=np.hstack)))
    assert_close(np.asarray(chain(b2)(x)), chain(ref)(x), what="log-filtered chain, frame %d fft %d" % (frame_size, fft_size))
    with pytest.raises(ValueError):
        b2.ShortTimeFourierTransform(b2.FramedSignal(b2.Signal(x, sample_rate=SR), frame_size=frame_size), fft_size=3000)


@pytest.mark.parametrize("frame_size", [1024, 2048, 8192])
def test_circular_shift(b2, frame_size):
    """madmom stft(circular_shift=True): the halves of the windowed frame are swapped = bin k times (-1)^k"""
    from audio_tabs_b200.synth import synth_guitar
    x = synth_guitar(3960 + frame_size, 0.7)

    def stft_of(m, **kw):
        return m.ShortTimeFourierTransform(m.FramedSignal(m.Signal(x, sample_rate=SR), frame_size=frame_size), **kw)
    want = stft_of(ref, circular_shift=True).data
    got = np.asarray(stft_of(b2, circular_shift=True))
    assert_stft_close(got, want)
    plain = np.asarray(stft_of(b2))
    assert np.array_equal(got[:, 0::2], plain[:, 0::2]) and np.array_equal(got[:, 1::2], -plain[:, 1::2])
    # magnitudes, and everything downstream of them, do not change
    spec = np.asarray(b2.Spectrogram(stft_of(b2, circular_shift=True)))
    assert_close(spec, np.abs(want), rtol=1e-4, atol=1e-4 * np.abs(want).max(), what="spectrogram of the shifted STFT")
    with pytest.raises(ValueError):
        b2.ShortTimeFourierTransform(b2.FramedSignal(b2.Signal(x, sample_rate=SR), frame_size=1000), fft_size=1024,
                                     circular_shift=True)


@pytest.mark.parametrize("case", ["beat", "beat_one_launch", "beat_i16_stereo", "chroma_4096", "deep_8192", "mel_2048", "superflux"])
def test_guard_bands_around_inputs_and_outputs(b2, case):
    """compute-sanitizer is closed on this pool; this is the bounds check that stands in for it.  The packed samples
    sit between NaN guard bands at an ODD element offset (no alignment beyond the element size), the output rows
    between sentinel rows: a read outside the clips would put NaN into a result (NaN x 0 = NaN, so zero-weight taps
    and zero window ends count), a write outside the rows would destroy a sentinel.  Results must be finite,
    bitwise equal to the run on plain tensors, and the guards untouched."""
    import torch
    from audio_tabs_b200.frontends import beat_specs, log_filt_spec
    from audio_tabs_b200.onsets import mel_db_spec
    from audio_tabs_b200.plan import FrontEnd, Packed
    from audio_tabs_b200.synth import synth_guitar
    dev = torch.device("cuda", 0)
    dtype, channels, kw, proj = "f32", 1, {}, False
    if case == "beat":
        specs = beat_specs()
    elif case == "beat_one_launch":
        specs, kw = beat_specs(), dict(one_launch=True)
    elif case == "beat_i16_stereo":
        specs, dtype, channels = beat_specs(int16=True), "i16", 2
    elif case == "chroma_4096":
        specs, proj = [log_filt_spec(4096, 4410.0, 24, 65.0, 2100.0, fold=True)], True
    elif case == "deep_8192":
        specs = [log_filt_spec(8192, 4410.0, 24, 65.0, 2100.0)]
    elif case == "mel_2048":
        specs, kw = [mel_db_spec(44100)], dict(end="extend")
    else:
        specs = [log_filt_spec(2048, 441.0, 12, diff_ratio=0.5, diff_max_bins=3)]
    fe = FrontEnd(specs, device=0, dtype=dtype, channels=channels, **kw)
    lens = [44100, 1, 700, 3 * 8192 + 5, 0, 30011]
    clips = []
    for i, n in enumerate(lens):
        y = synth_guitar(9900 + i, max(n, 64) / SR)[:n]
        if channels == 2:
            y = np.stack([y, 0.5 * np.roll(y, 2)], axis=1)
        if dtype == "i16":
            y = np.clip(np.round(y * 25000), -32768, 32767).astype(np.int16)
        clips.append(y)
    flat = np.concatenate([c.reshape(-1) for c in clips])
    tdt = torch.int16 if dtype == "i16" else torch.float32
    G = 8191 * channels                                   # odd guard length (per channel): nothing may rely on alignment
    big = torch.empty(flat.size + 2 * G, dtype=tdt, device=dev)
    if dtype == "i16":
        big.fill_(32767)                                  # no NaN in int16: full-scale guards change any sum they enter
    else:
        big.fill_(float("nan"))
    big[G:G + flat.size] = torch.from_numpy(flat).to(dev)
    plain = torch.from_numpy(flat).to(dev)

    def run(sig, guarded):
        packed = Packed(sig, lens, fe.hop_size, fe.end)
        T = packed.total_frames
        width = specs[0].num_classes if proj else fe.width
        R = 37                                            # sentinel rows on both sides
        buf = torch.full((T + 2 * R, width), -12345.0, dtype=torch.float32, device=dev)
        out = buf[R:R + T]
        if proj:
            fe.run_packed(packed, out=False, proj=[out])
        else:
            fe.run_packed(packed, out)
        torch.cuda.synchronize()
        assert bool((buf[:R] == -12345.0).all()) and bool((buf[R + T:] == -12345.0).all()), "write outside the output rows"
        return out.cpu().numpy()

    want = run(plain, False)
    got = run(big[G:G + flat.size], True)
    assert np.isfinite(got).all(), "a guard value reached the results: read outside the packed clips"
    assert np.array_equal(got, want)
    assert not bool((want == -12345.0).any()), "an output element was never written"
    head, tail = big[:G], big[G + flat.size:]
    if dtype == "i16":
        assert bool((head == 32767).all()) and bool((tail == 32767).all())
    else:
        assert bool(torch.isnan(head).all()) and bool(torch.isnan(tail).all())


@pytest.mark.parametrize("frame_size,dtype", [(1024, "f32"), (2048, "i16"), (4096, "f32"), (8192, "f32")])
def test_include_nyquist(b2, frame_size, dtype):
    """madmom stft(include_nyquist=True): frame_size/2 + 1 bins, the last one the (real) Nyquist bin; every other
    bin is bitwise what the plain call gives.  Holds for the magnitude, log and difference stages on raw bins;
    a filterbank takes frame_size/2 bins and refuses."""
    from audio_tabs_b200.synth import synth_guitar
    x = synth_guitar(3990 + frame_size, 0.6)
    x = x + 0.2 * np.cos(np.pi * np.arange(len(x))).astype(np.float32)     # energy AT the Nyquist frequency
    if dtype == "i16":
        x = np.clip(np.round(x * 20000), -32768, 32767).astype(np.int16)

    def stft_of(m, **kw):
        return m.ShortTimeFourierTransform(m.FramedSignal(m.Signal(x, sample_rate=SR), frame_size=frame_size), **kw)
    want = stft_of(ref, include_nyquist=True)
    ours = stft_of(b2, include_nyquist=True)
    got = np.asarray(ours)
    N = frame_size // 2
    assert got.shape == want.data.shape == (len(ours), N + 1) and ours.num_bins == N + 1
    np.testing.assert_allclose(ours.bin_frequencies, want.bin_frequencies)
    assert_stft_close(got, want.data)
    assert np.abs(got[:, N]).max() > 10 * np.abs(got[:, N - 8:N - 2]).mean() and np.all(got[:, N].imag == 0)
    assert np.array_equal(got[:, :N], np.asarray(stft_of(b2)))
    circ = np.asarray(stft_of(b2, include_nyquist=True, circular_shift=True))
    assert_stft_close(circ, stft_of(ref, include_nyquist=True, circular_shift=True).data)
    spec = np.asarray(b2.Spectrogram(ours))
    assert spec.shape == (len(ours), N + 1)
    assert_close(spec, np.abs(want.data), rtol=1e-4, atol=1e-4 * np.abs(want.data).max(), what="magnitudes with Nyquist")
    log = np.asarray(b2.LogarithmicSpectrogram(b2.Spectrogram(ours), mul=1, add=1))
    # raw bins (no band sums): a bin far below the frame's peak carries the FFT's absolute rounding noise, 1e-4 max|X|
    # at most (the magnitude bound above); d/dm log10(1 + m) <= 0.4343
    tol = max(2e-5, 0.4343e-4 * float(np.abs(want.data).max()))
    assert_close(log, np.log10(np.abs(want.data).astype(np.float32) + 1), rtol=1e-4, atol=tol, what="log with Nyquist")
    d = np.asarray(b2.SpectrogramDifference(b2.LogarithmicSpectrogram(b2.Spectrogram(ours), mul=1, add=1), diff_frames=1,
                                            positive_diffs=True))
    assert d.shape == (len(ours), N + 1)
    assert_close(d, ref.spectrogram_difference(log, 1, positive_diffs=True), atol=2e-5, what="diff with Nyquist")   # of OUR log rows
    with pytest.raises(ValueError, match="include_nyquist"):
        np.asarray(b2.FilteredSpectrogram(b2.Spectrogram(ours), num_bands=12))


@pytest.mark.parametrize("seed", list(range(40)))
def test_randomized_configurations(b2, seed):
    """Seeded random configurations of the fused chain -- frame size, (fractional) hop, origin, bands per octave,
    frequency range, mul / add, difference lag, sample format, channel count, ragged clip lengths down to one
    sample -- against the oracle: the cases nobody thought of writing down."""
    from audio_tabs_b200.frontends import log_filt_spec
    from audio_tabs_b200.plan import FrontEnd
    from audio_tabs_b200.synth import synth_guitar
    rng = np.random.default_rng(9000 + seed)
    frame_size = int(rng.choice([1024, 2048, 4096, 8192]))
    hop = float(rng.choice([441.0, 512.0, 4410.0, 8820.0, SR / 7.0, float(rng.integers(64, 6000)) + float(rng.choice([0.0, 0.5, 0.25]))]))
    bpo = int(rng.choice([3, 6, 12, 24]))
    fmin = float(rng.choice([30.0, 65.0, 100.0, 27.5]))
    fmax = float(rng.choice([17000.0, 2100.0, 8000.0, 16000.0]))
    mul, add = float(rng.choice([1.0, 5.0, 0.5])), float(rng.choice([1.0, 0.5, 2.0]))
    ratio = rng.choice([None, 0.5, 0.25])
    dtype = str(rng.choice(["f32", "i16"]))
    channels = int(rng.choice([1, 1, 2]))
    origin = int(rng.choice([0, 0, 0, (frame_size - 1) // 2, -(frame_size // 2)]))      # 'center', 'online', 'stream'
    try:
        spec = log_filt_spec(frame_size, hop, bpo, fmin, fmax, mul=mul, add=add, diff_ratio=None if ratio is None else float(ratio),
                             int16=(dtype == "i16"), origin=origin)
    except ValueError:
        pytest.skip("filterbank not constructible for this draw")
    lens = [int(v) for v in rng.choice([1, 7, 300, frame_size - 1, frame_size, 3 * frame_size + 17, 40000, 66150], size=5)]
    clips = []
    for i, n in enumerate(lens):
        y = synth_guitar(9100 + 10 * seed + i, max(n, 64) / SR)[:n]
        if channels == 2:
            y = np.stack([y, np.roll(y, 3) * 0.5], axis=1)
        if dtype == "i16":
            y = np.clip(np.round(y * 28000), -32768, 32767).astype(np.int16)
        clips.append(np.ascontiguousarray(y))
    fe = FrontEnd([spec], device=0, dtype=dtype, channels=channels)
    got = fe.process_batch(clips)
    for c, g in zip(clips, got):
        mono = ref.remix(c, 1) if channels == 2 else c
        chain = [ref.SignalProcessor(num_channels=1, sample_rate=SR),
                 ref.FramedSignalProcessor(frame_size=frame_size, hop_size=hop, origin=origin),
                 ref.ShortTimeFourierTransformProcessor(),
                 ref.FilteredSpectrogramProcessor(num_bands=bpo, fmin=fmin, fmax=fmax),
                 ref.LogarithmicSpectrogramProcessor(mul=mul, add=add)]
        if ratio is not None:
            chain.append(ref.SpectrogramDifferenceProcessor(diff_ratio=float(ratio), positive_diffs=True, stack_diffs=np.hstack))
        want = ref.SequentialProcessor(chain)(mono)
        want = np.asarray(want.data if hasattr(want, "data") and not isinstance(want, np.ndarray) else want).astype(np.float32)
        what = "seed %d: frame %d hop %r bpo %d %s/%d origin %d diff %s" % (seed, frame_size, hop, bpo, dtype, channels, origin, ratio)
        if ratio is None:
            assert_close(g, want, what=what)
        else:
            # [spec | diff]: a difference of two log values carries the absolute error of both (near 0 the relative
            # term gives no room), so its absolute tolerance is twice that of a value
            B = want.shape[1] // 2
            assert_close(g[:, :B], want[:, :B], what=what + " (spec)")
            assert_close(g[:, B:], want[:, B:], atol=2e-5, what=what + " (diff)")
