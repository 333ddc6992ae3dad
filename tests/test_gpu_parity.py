"""GPU parity: CUDA path (through the C ABI) vs the numpy oracle on the same seeded inputs.

Tolerances: rtol 1e-4 / atol 1e-5 on log-filtered, diff and chroma outputs; raw STFT
within 1e-4 of each frame's peak; frame counts bit-exact.
"""
import numpy as np
import pytest

from helpers import assert_close, assert_stft_close, noise
from oracle import madmom_ref as ref

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def b2(cuda_device):
    import audio_tabs_b200
    return audio_tabs_b200


@pytest.mark.parametrize("frame_size", [1024, 2048, 4096, 8192])
def test_stft_matches_oracle(b2, frame_size):
    x = noise(frame_size, 44100 * 2 + 123)
    frames = b2.FramedSignalProcessor(frame_size=frame_size, hop_size=441.0)(b2.Signal(x, sample_rate=44100))
    got = np.asarray(b2.ShortTimeFourierTransformProcessor()(frames))
    want = ref.ShortTimeFourierTransform(ref.FramedSignal(ref.Signal(x, sample_rate=44100), frame_size=frame_size)).data
    assert got.shape == want.shape
    assert_stft_close(got, want)


@pytest.mark.parametrize("frame_size,num_bands", [(1024, 3), (2048, 6), (4096, 12), (2048, 12), (8192, 24)])
def test_log_filtered_spectrogram(b2, frame_size, num_bands):
    x = noise(7 + frame_size, 44100 * 3 + 17)
    chain = b2.SequentialProcessor((
        b2.SignalProcessor(num_channels=1, sample_rate=44100),
        b2.FramedSignalProcessor(frame_size=frame_size, hop_size=441.0),
        b2.ShortTimeFourierTransformProcessor(),
        b2.FilteredSpectrogramProcessor(num_bands=num_bands, fmin=30, fmax=17000, norm_filters=True),
        b2.LogarithmicSpectrogramProcessor(mul=1, add=1),
    ))
    got = np.asarray(chain(x))
    want = ref.log_filtered_spectrogram(x, frame_size=frame_size, num_bands=num_bands)
    assert_close(got, want, what="logfilt %d" % frame_size)


def test_rnn_beat_frontend_314(b2):
    from audio_tabs_b200.frontends import rnn_beat_frontend, rnn_beat_frontend_fused
    from audio_tabs_b200.synth import synth_guitar
    x = synth_guitar(2001, 6.0)
    want = ref.rnn_beat_preprocessor()(x)
    got = rnn_beat_frontend()(x)
    assert got.shape == (600, 314)
    assert_close(got, want, what="beat front end (per-branch)")
    got2 = rnn_beat_frontend_fused()(x)
    assert_close(got2, want, what="beat front end (single buffer)")
    np.testing.assert_array_equal(got, got2)


def test_rnn_onset_frontend_266(b2):
    from audio_tabs_b200.frontends import rnn_onset_frontend
    x = noise(11, 44100 * 2)
    want = ref.rnn_onset_preprocessor()(x)
    got = rnn_onset_frontend()(x)
    assert got.shape == want.shape == (200, 266)
    assert_close(got, want, what="onset front end")
