// fb_pack.h -- host-side packing of a banded filterbank into the slab form the fused kernel uses
// (fb_slabs in frontend_kernel.cuh).  Plain C++ (no CUDA) so tests/emu can check it on the CPU.
//
// Replaces the dense np.dot(spec, filterbank) of madmom.audio.spectrogram.FilteredSpectrogram
// (reached from /root/reference/backend/app/services/grid/beats.py:74): the used bins [kmin, kmax)
// are cut into NS * 128 slabs of L consecutive bins (L odd, so the 32 lanes of a warp never share a
// shared-memory bank).  A thread keeps four running sums per frame, indexed by (band & 3), so a slab
// may touch at most one band per residue.  madmom's triangular banks overlap only their neighbours,
// which makes that true wherever bands are wider than ~L/3 bins; the narrow low bands that break it
// are "direct": the band stage sums their few taps straight from the magnitudes.
#pragma once
#include <cstddef>
#include <vector>

namespace b2 {

struct FbBand {      // mirrors the int4 the kernel reads
  int x, y, z, w;    // slab band: {first partial index, count, 0, 0}; direct band: {dw offset, 0, first bin, taps}
};

struct FbPack {
  int L = 3, NS = 1, kmin = 0, kmax = 0, ndirect = 0;
  std::vector<float> w4;       // [NS][L][128][4]
  std::vector<FbBand> band;    // per band
  std::vector<float> dw;       // weights of the direct bands
};

// tbf = frames the kernel pushes through the stage at a time (FftCfg<F>::TBF); it only steers the
// choice of L: the slab pass costs ~NS*L*(1 + 5 tbf) instructions per thread, the band stage
// ~(widest band / L) * 2 tbf on the lane that owns the widest band.
inline FbPack fb_pack(int num_bins, int B, const int *band_start, const int *band_len, const int *band_woff,
                      const float *weights, int tbf = 2, int force_L = 0, int NT = 128) {
  FbPack P;
  int kmin = num_bins, kmax = 0;
  for (int j = 0; j < B; ++j) {
    if (band_len[j] <= 0) continue;
    if (band_start[j] < kmin) kmin = band_start[j];
    if (band_start[j] + band_len[j] > kmax) kmax = band_start[j] + band_len[j];
  }
  if (kmin > kmax) kmin = kmax;
  const int range = kmax - kmin;
  int wmax = 1;
  for (int j = 0; j < B; ++j)
    if (band_len[j] > wmax) wmax = band_len[j];
  int L = 3, NS = 1;
  long best = -1;
  for (int l = (force_L ? force_L : 3); l <= (force_L ? force_L : 15); l += 2) {
    int ns = (range + NT * l - 1) / (NT * l);
    if (ns < 1) ns = 1;
    const long cost = (long)ns * l * (1 + 5 * tbf) + (long)((wmax + l - 1) / l) * 2 * tbf + 4L * ns * tbf;
    if (best < 0 || cost < best) best = cost, L = l, NS = ns;
  }
  const int nslab = NS * NT;
  std::vector<char> direct((size_t)(B > 0 ? B : 1), 0);
  for (int g = 0; g < nslab; ++g) {
    const int lo = kmin + g * L, hi = lo + L;
    std::vector<int> touch;
    for (int j = 0; j < B; ++j)
      if (band_len[j] > 0 && !direct[j] && band_start[j] < hi && band_start[j] + band_len[j] > lo) touch.push_back(j);
    for (;;) {
      int cnt[4] = {0, 0, 0, 0};
      bool clash = false;
      for (int j : touch) clash |= (++cnt[j & 3] > 1);
      if (!clash) break;
      size_t narrow = 0;                               // drop the narrowest band of the slab
      for (size_t i = 1; i < touch.size(); ++i)
        if (band_len[touch[i]] < band_len[touch[narrow]]) narrow = i;
      direct[touch[narrow]] = 1;
      touch.erase(touch.begin() + (long)narrow);
    }
  }
  P.L = L;
  P.NS = NS;
  P.kmin = kmin;
  P.kmax = kmax;
  P.w4.assign((size_t)NS * L * NT * 4, 0.f);
  P.band.assign((size_t)(B > 0 ? B : 1), FbBand{0, 0, 0, 0});
  for (int j = 0; j < B; ++j) {
    const int len = band_len[j], st = band_start[j];
    if (len <= 0) continue;
    const float *w = weights + band_woff[j];
    if (direct[j]) {
      P.band[j] = FbBand{(int)P.dw.size(), 0, st, len};
      P.dw.insert(P.dw.end(), w, w + len);
      ++P.ndirect;
    } else {
      for (int i = 0; i < len; ++i) {
        const int g = (st + i - kmin) / L, ii = (st + i - kmin) % L;
        P.w4[(((size_t)(g / NT) * L + ii) * NT + (size_t)(g % NT)) * 4 + (size_t)(j & 3)] = w[i];
      }
      const int g0 = (st - kmin) / L, g1 = (st + len - 1 - kmin) / L;
      P.band[j] = FbBand{g0 * 4 + (j & 3), g1 - g0 + 1, 0, 0};
    }
  }
  return P;
}

}  // namespace b2
