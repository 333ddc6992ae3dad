"""Shared test helpers: seeded inputs and tolerance checks (SURVEY.md §8c tolerances)."""
import numpy as np

RTOL, ATOL = 1e-4, 1e-5          # fp32 log-filtered / diff / chroma outputs vs the oracle
STFT_REL_TO_PEAK = 1e-4          # raw STFT: |delta| <= 1e-4 * max_k |X[n, k]| per frame


def noise(seed, n, scale=0.1, dtype=np.float32):
    rng = np.random.default_rng(seed)
    x = rng.standard_normal(n) * scale
    if dtype == np.int16:
        return np.clip(x * 32767, -32768, 32767).astype(np.int16)
    return x.astype(dtype)


def assert_close(got, want, rtol=RTOL, atol=ATOL, what=""):
    got, want = np.asarray(got), np.asarray(want)
    assert got.shape == want.shape, (what, got.shape, want.shape)
    assert got.dtype == want.dtype, (what, got.dtype, want.dtype)
    err = np.abs(got.astype(np.float64) - want.astype(np.float64))
    bound = atol + rtol * np.abs(want.astype(np.float64))
    bad = err > bound
    assert not bad.any(), "%s: %d/%d outside rtol=%g atol=%g, max err %.3e at %s" % (
        what, bad.sum(), bad.size, rtol, atol, err.max(), np.unravel_index(np.argmax(err), err.shape))


def assert_stft_close(got, want, rel=STFT_REL_TO_PEAK):
    got, want = np.asarray(got), np.asarray(want)
    assert got.shape == want.shape and got.dtype == np.complex64
    peak = np.abs(want).max(axis=1, keepdims=True)
    err = np.abs(got.astype(np.complex128) - want.astype(np.complex128))
    bound = rel * np.maximum(peak, 1e-30) + 1e-12
    assert (err <= bound).all(), "stft max err/peak %.3e" % (err / np.maximum(peak, 1e-30)).max()
