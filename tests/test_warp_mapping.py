"""Design check of k_front_warp (audio_tabs_b200/csrc/frontend_warp_kernel.cuh) on the CPU: a lane-by-lane numpy
emulation of its data flow with the kernel's own index formulas -- 32 x 32 decomposition of the 1024-point FFT,
twiddle table layout, the transpose through the row-stride-33 tile, the mirror-bin exchange by shuffle (lane 0's
special case), the pair split (frame 1024) and the even/odd split with the per-lane twiddle base times W_64^d
(frame 2048, bin 512 apart) -- against numpy's FFT, plus the two properties the warp-private tile relies on:
between two __syncwarp() no lane touches an element another lane writes, and every 64-bit access pattern is free of
bank conflicts.  (compute-sanitizer's racecheck is closed on the GPU pool; the GPU tests check the kernel's results,
this checks that the mapping is hazard free by construction.)"""
import numpy as np
import pytest

N = 1024
STRIDE = 33            # kWarpTile = 32 * 33 float2


def _fft_by_lanes(z):
    """Z[c + 32 d] in lane c, element d -- the two DFT32 passes around the tile transpose."""
    tile = np.zeros(32 * STRIDE, complex)
    writes = []                                             # (lane, tile index) of phase 1
    for b in range(32):                                     # lane b holds z[32 a + b]
        x = z[32 * np.arange(32) + b]
        y = np.fft.fft(x)                                   # DFT32 over a -> index c
        for c in range(32):
            tile[c * STRIDE + b] = y[c] * np.exp(-2j * np.pi * b * c / N)      # s_tw[c * 32 + b] = W_1024^(b c)
            writes.append((b, c * STRIDE + b))
    reads, out = [], np.zeros((32, 32), complex)
    for c in range(32):                                     # lane c reads its row
        u = np.array([tile[c * STRIDE + b] for b in range(32)])
        reads += [(c, c * STRIDE + b) for b in range(32)]
        out[c] = np.fft.fft(u)                              # DFT32 over b -> index d
    return out, writes, reads


def _mirror(zl, c, d):
    """What lane c receives for its bin k = c + 32 d: Z[N - k], held by lane (32 - c) & 31 (the kernel's shuffle)."""
    partner = (32 - c) & 31
    give = zl[partner, (0 if d == 0 else 32 - d)] if partner == 0 else zl[partner, 31 - d]
    return give


def test_fft_decomposition_and_mirror_exchange():
    rng = np.random.default_rng(1)
    z = rng.standard_normal(N) + 1j * rng.standard_normal(N)
    zl, _, _ = _fft_by_lanes(z)
    Z = np.fft.fft(z)
    for c in range(32):
        np.testing.assert_allclose(zl[c], Z[c + 32 * np.arange(32)], rtol=0, atol=1e-9)
        for d in range(16):
            k = c + 32 * d
            np.testing.assert_allclose(_mirror(zl, c, d), Z[(N - k) % N], rtol=0, atol=1e-9)


def test_pair_split_frame_1024():
    """z = xA + i xB: |XA[k]| = |Z[k] + conj Z[N-k]| / 2, |XB[k]| = |Z[k] - conj Z[N-k]| / 2 for k < 512 (the kernel
    folds the 1/2 into the window)."""
    rng = np.random.default_rng(2)
    xa, xb = rng.standard_normal(N), rng.standard_normal(N)
    zl, _, _ = _fft_by_lanes(0.5 * (xa + 1j * xb))
    ma, mb = np.zeros(512), np.zeros(512)
    for c in range(32):
        for d in range(16):
            k, r = c + 32 * d, _mirror(zl, c, d)
            z = zl[c, d]
            ma[k], mb[k] = abs(z + np.conj(r)), abs(z - np.conj(r))
    np.testing.assert_allclose(ma, np.abs(np.fft.rfft(xa))[:512], atol=1e-9)
    np.testing.assert_allclose(mb, np.abs(np.fft.rfft(xb))[:512], atol=1e-9)


def test_even_odd_split_frame_2048():
    """z[m] = x[2m] + i x[2m+1]: X[k] = P + pt_k M, X[N-k] = conj(P - pt_k M), pt_k = -i W_2048^k formed as the lane's
    base (-sin, -cos)(pi lane / 1024) times W_64^d; bin 512 = 2 |Z[512]|."""
    rng = np.random.default_rng(3)
    x = rng.standard_normal(2048)
    zl, _, _ = _fft_by_lanes(0.5 * (x[0::2] + 1j * x[1::2]))
    mags = np.full(1024, np.nan)
    for c in range(32):
        th = np.pi * c / 1024.0
        ptc = complex(-np.sin(th), -np.cos(th))
        for d in range(16):
            k, r = c + 32 * d, _mirror(zl, c, d)
            z = zl[c, d]
            P, M = z + np.conj(r), z - np.conj(r)
            pt = ptc * complex(np.cos(np.pi * d / 32), -np.sin(np.pi * d / 32))        # W_64^d
            assert abs(pt - (-1j) * np.exp(-2j * np.pi * k / 2048)) < 1e-12
            mags[k] = abs(P + M * pt)
            if k > 0:
                mags[1024 - k] = abs(P - M * pt)
    mags[512] = 2 * abs(zl[0, 16])
    assert not np.isnan(mags).any()                          # every bin 0 .. 1023 is produced exactly by these writes
    np.testing.assert_allclose(mags, np.abs(np.fft.rfft(x))[:1024], atol=1e-9)


def _banks64(indices):
    """32-bit banks touched by one half-warp's 64-bit accesses to float2 elements `indices`."""
    banks = []
    for i in indices:
        banks += [(2 * i) % 32, (2 * i + 1) % 32]
    return banks


def test_tile_phases_are_hazard_free_and_conflict_free():
    _, writes, reads = _fft_by_lanes(np.zeros(N, complex))
    # phase 1 (stores) and phase 2 (loads) are separated by __syncwarp; inside a phase every element has ONE owner
    for acc in (writes, reads):
        owner = {}
        for lane, idx in acc:
            assert owner.setdefault(idx, lane) == lane
        assert max(i for _, i in acc) < 32 * STRIDE
    # 64-bit accesses go out per half-warp: 16 lanes x 2 words must hit 32 different banks
    for cc in range(32):                                     # store of element cc: lane b -> T[cc * 33 + b]
        for half in (range(16), range(16, 32)):
            b = _banks64([cc * STRIDE + lane for lane in half])
            assert len(set(b)) == 32
    for bcol in range(32):                                   # load of column b: lane c -> T[c * 33 + b]
        for half in (range(16), range(16, 32)):
            b = _banks64([lane * STRIDE + bcol for lane in half])
            assert len(set(b)) == 32


@pytest.mark.parametrize("nf,nbins", [(2, 512), (1, 1024)])
def test_magnitude_writes_have_one_owner_and_fit_the_tile(nf, nbins):
    """Phase 3 (between two __syncwarp): the magnitudes overwrite the tile; every float has one writing lane, the
    16 padding bins included, and everything stays inside the 32 x 33 float2 tile."""
    owner = {}

    def put(lane, idx):
        assert 0 <= idx < 2 * 32 * STRIDE
        assert owner.setdefault(idx, lane) == lane

    for lane in range(32):
        if lane < 16:
            for t in range(nf):
                put(lane, nf * (nbins + lane) + t)
        for d in range(16):
            k = lane + 32 * d
            if nf == 2:
                put(lane, 2 * k)
                put(lane, 2 * k + 1)
            else:
                put(lane, k)
                if k > 0:
                    put(lane, 1024 - k)
        if nf == 1 and lane == 0:
            put(lane, 512)
    assert sorted(owner) == list(range(nf * (nbins + 16)))   # bins 0 .. nbins + 15 of every frame, nothing else
