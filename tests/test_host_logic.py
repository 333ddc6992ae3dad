"""Host-side logic that needs no GPU: filterbank construction (bit-identical to the oracle), madmom's
processor protocol, lazy chain bookkeeping, error types, sharding, synthetic inputs."""
import numpy as np
import pytest

import audio_tabs_b200 as b2
from audio_tabs_b200 import filters, sharding
from audio_tabs_b200.audio import spectrogram as sp
from audio_tabs_b200.audio.stft import fft_frequencies
from oracle import madmom_ref as ref

SR = 44100


@pytest.mark.parametrize("frame_size,bpo,fmin,fmax,unique,norm", [
    (1024, 3, 30, 17000, True, True), (2048, 6, 30, 17000, True, True), (4096, 12, 30, 17000, True, True),
    (2048, 12, 30, 17000, False, True), (8192, 24, 65, 2100, True, True), (8192, 24, 60, 2600, True, True),
    (4096, 24, 65, 2100, True, False), (2048, 12, 30, 17000, True, False),
])
def test_filterbank_bit_identical_to_oracle(frame_size, bpo, fmin, fmax, unique, norm):
    bf = fft_frequencies(frame_size >> 1, SR)
    ours = filters.LogarithmicFilterbank(bf, num_bands=bpo, fmin=fmin, fmax=fmax, unique_filters=unique,
                                         norm_filters=norm)
    want = ref.LogarithmicFilterbank(bf, num_bands=bpo, fmin=fmin, fmax=fmax, unique_filters=unique,
                                     norm_filters=norm)
    assert ours.dtype == np.float32
    np.testing.assert_array_equal(np.asarray(ours), want.data)
    np.testing.assert_array_equal(ours.center_frequencies, want.center_frequencies)
    start, length, woff, w = ours.banded()
    dense = np.zeros_like(np.asarray(ours))
    for j in range(len(start)):
        dense[start[j]:start[j] + length[j], j] = w[woff[j]:woff[j] + length[j]]
    np.testing.assert_array_equal(dense, np.asarray(ours))


def test_pcp_filterbank_matches_oracle():
    bf = fft_frequencies(2048, SR)
    np.testing.assert_array_equal(np.asarray(filters.PitchClassProfileFilterbank(bf)),
                                  ref.PitchClassProfileFilterbank(bf).data)
    fb = filters.LogarithmicFilterbank(bf, num_bands=24, fmin=65, fmax=2100)
    np.testing.assert_array_equal(filters.fold_classes(fb.center_frequencies), ref.fold_classes(fb.center_frequencies))


def test_filter_errors_like_madmom():
    with pytest.raises(ValueError):
        filters.TriangularFilter(5, 4, 9)
    with pytest.raises(TypeError):
        filters.Filterbank([1, 2, 3], [0.0])
    with pytest.raises(ValueError):
        filters.Filterbank(np.zeros((4, 2)), [0.0, 1.0])


def test_processor_protocol():
    seq = b2.SequentialProcessor([lambda x: x + 1, b2.SequentialProcessor([lambda x: x * 2])])
    assert len(seq) == 2 and seq(3) == 8
    par = b2.ParallelProcessor([lambda x: x + 1, lambda x: x - 1])
    assert par(1) == [2, 0]
    seq.append(lambda x: -x)
    assert seq(3) == -8

    class Kw(b2.Processor):
        def process(self, data, **kwargs):
            return kwargs.get("scale", 1) * data
    assert b2.SequentialProcessor([Kw(), np.negative])(2, scale=5) == -10
    with pytest.raises(NotImplementedError):
        b2.Processor()(1)


def test_framed_signal_geometry(lib_built):
    x = np.arange(123481, dtype=np.float32)
    fr = b2.FramedSignalProcessor(frame_size=2048, hop_size=441.0)(b2.Signal(x, sample_rate=SR))
    want = ref.FramedSignal(ref.Signal(x, sample_rate=SR), frame_size=2048)
    assert len(fr) == want.num_frames == 281 and fr.shape == (281, 2048) and fr.fps == 100.0
    for i in (0, 1, 2, 3, 140, 279, 280, -1):
        np.testing.assert_array_equal(fr[i], want[i])
    with pytest.raises(IndexError):
        fr[281]
    with pytest.raises(ValueError):
        b2.FramedSignal(b2.Signal(x, sample_rate=SR), end="bogus")
    fps = b2.FramedSignal(b2.Signal(x, sample_rate=SR), frame_size=8192, fps=30)
    assert fps.hop_size == SR / 30.0 and len(fps) == ref.num_frames_for(len(x), SR / 30.0)
    assert b2.FramedSignal(b2.Signal(x, sample_rate=SR), origin="online").origin == 1023
    assert b2.FramedSignal(b2.Signal(x, sample_rate=SR), origin="future").origin == -1024
    assert len(fr[10:20]) == 10


def test_signal_remix_matches_madmom(lib_built):
    rng = np.random.default_rng(0)
    st = rng.standard_normal((1000, 2)).astype(np.float32)
    np.testing.assert_array_equal(np.asarray(b2.Signal(st, sample_rate=SR, num_channels=1)),
                                  ref.Signal(st, sample_rate=SR, num_channels=1).data)
    sti = rng.integers(-32768, 32767, size=(1000, 2)).astype(np.int16)
    got = b2.Signal(sti, sample_rate=SR, num_channels=1)
    assert got.dtype == np.int16 and got.num_channels == 1
    np.testing.assert_array_equal(np.asarray(got), ref.Signal(sti, sample_rate=SR, num_channels=1).data)


def test_lazy_chain_bookkeeping(lib_built):
    from audio_tabs_b200.engine import _parse, _spec_from
    x = np.zeros(SR, np.float32)
    chain = b2.SequentialProcessor((
        b2.SignalProcessor(num_channels=1, sample_rate=SR), b2.FramedSignalProcessor(frame_size=4096, fps=100),
        b2.ShortTimeFourierTransformProcessor(), b2.FilteredSpectrogramProcessor(num_bands=12, fmin=30, fmax=17000),
        b2.LogarithmicSpectrogramProcessor(mul=1, add=1),
        b2.SpectrogramDifferenceProcessor(diff_ratio=0.5, positive_diffs=True, stack_diffs=np.hstack)))
    out = chain(x)                                   # nothing is computed yet
    assert out.shape == (100, 182) and out.dtype == np.float32
    rec = _parse(out)
    assert rec["stack"] and rec["diff"] == (2, True, 0) and rec["log"] == (1.0, 1.0)
    assert rec["filterbank"].shape == (2048, 91) and rec["magnitude"]
    spec = _spec_from(rec, stft=rec["stft"])
    assert (spec.frame_size, spec.hop_size, spec.num_bands, spec.diff_frames, spec.out_width) == (4096, 441.0, 91, 2, 182)
    # attributes the next madmom stage reads
    assert out.stft.frames.frame_size == 4096 and out.stft.window.shape == (4096,)
    assert out.diff.spectrogram.filterbank is rec["filterbank"]
    assert chain.processors[4].mul == 1 and chain.processors[5].diff_frames == 2    # cached like madmom


def test_error_types_like_madmom(lib_built):
    x = np.zeros((4000, 2), np.float32)
    frames = b2.FramedSignal(b2.Signal(x, sample_rate=SR), frame_size=1024)
    assert frames.shape == (10, 1024, 2)
    with pytest.raises(ValueError, match="frames must be a 2D array"):
        b2.ShortTimeFourierTransform(frames)
    mono = b2.ShortTimeFourierTransform(b2.FramedSignal(b2.Signal(x[:, 0], sample_rate=SR), frame_size=1024))
    with pytest.raises(TypeError):
        b2.FilteredSpectrogram(mono, filterbank=np.zeros((512, 3)))
    with pytest.raises(ValueError):
        b2.SpectrogramDifference(b2.Spectrogram(mono), diff_frames=0)
    assert mono.shape == (10, 512) and mono.dtype == np.complex64
    assert sp._diff_frames(0.25, 441.0, 4096) == 3


def test_int16_window_is_scaled_not_the_samples(lib_built):
    xi = np.zeros(5000, np.int16)
    s = b2.ShortTimeFourierTransformProcessor()(b2.FramedSignal(b2.Signal(xi, sample_rate=SR), frame_size=2048))
    np.testing.assert_array_equal(s.fft_window, np.hanning(2048) / 32767.0)
    np.testing.assert_array_equal(s.window, np.hanning(2048))


def test_sharding_partition():
    lens = [7, 3, 9, 1, 4, 4, 8, 2]
    for n in (1, 2, 3, 8):
        shards = sharding.partition(lens, n)
        assert sorted(i for s in shards for i in s) == list(range(len(lens)))
        loads = [sum(lens[i] for i in s) for s in shards]
        assert max(loads) - min(loads) <= max(lens)
    assert sharding.partition([5] * 8, 4) == [[0, 4], [1, 5], [2, 6], [3, 7]]     # equal lengths: round robin
    assert sharding.local_shard([5] * 8, 1, 4) == [1, 5]
    with pytest.raises(ValueError):
        sharding.partition(lens, 0)


def test_synth_guitar_is_deterministic_and_normalised():
    from audio_tabs_b200.synth import synth_guitar
    a, b = synth_guitar(7, 1.0), synth_guitar(7, 1.0)
    np.testing.assert_array_equal(a, b)
    assert a.dtype == np.float32 and a.shape == (SR,) and np.abs(a).max() == pytest.approx(1.0, abs=1e-6)
    assert not np.array_equal(a, synth_guitar(8, 1.0))


def test_context_stack_matches_oracle():
    from audio_tabs_b200.audio.chroma import context_stack
    spec = np.random.default_rng(0).standard_normal((20, 7)).astype(np.float32)
    np.testing.assert_array_equal(context_stack(spec, 15), ref.dcp_context(spec, 15))


def test_oracle_online_difference_equals_the_offline_one_row_by_row():
    """madmom's online mode (reset=False, BufferProcessor): feeding a spectrogram one row at a time gives, as the
    last row of every call, the row the offline difference has at that position."""
    from oracle import madmom_ref as ref
    rng = np.random.default_rng(5)
    L = rng.random((40, 9)).astype(np.float32)
    for k, max_bins, positive in [(1, None, True), (3, None, False), (2, 3, True)]:
        want = ref.spectrogram_difference(L, k, max_bins, positive)
        proc = ref.SpectrogramDifferenceProcessor(diff_frames=k, diff_max_bins=max_bins, positive_diffs=positive)
        first = proc(ref._Spec(L[:5]), reset=True)
        np.testing.assert_array_equal(first.data, want[:5])
        for t in range(5, 40):
            out = proc(ref._Spec(L[t:t + 1]), reset=False)
            assert out.data.shape == (5, 9)
            np.testing.assert_array_equal(out.data[-1], want[t])
        stacked = ref.SpectrogramDifferenceProcessor(diff_frames=k, positive_diffs=True, stack_diffs=np.hstack)
        a = stacked(ref._Spec(L[:6]), reset=True)
        b = stacked(ref._Spec(L[6:9]), reset=False)
        assert a.shape == b.shape == (6, 18)
        np.testing.assert_array_equal(b[:, :9], L[3:9])
        np.testing.assert_array_equal(b[-3:, 9:], ref.spectrogram_difference(L, k, None, True)[6:9])
