// CPU emulation of the fused filterbank stage: fb_pack.h (host packing) + fb_core.cuh (the thread
// mapping k_front runs), checked against the dense banded product in float64.  Test infrastructure
// only.  usage: emu_fb <file>  with lines "F B" / start[B] / len[B] / weights[sum len]
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../audio_tabs_b200/csrc/fb_core.cuh"
#include "../../audio_tabs_b200/csrc/fb_pack.h"

template <int F>
int run(int B, const std::vector<int> &st, const std::vector<int> &ln, const std::vector<int> &wo,
        const std::vector<float> &w) {
  using C = b2::FftCfg<F>;
  constexpr int TBF = C::TBF, MS = C::MS, N = C::N;
  const b2::FbPack P = b2::fb_pack(N, B, st.data(), ln.data(), wo.data(), w.data(), TBF);
  // no slab may hold two bands with the same index modulo 4
  for (int g = 0; g < P.NS * 128; ++g)
    for (int r = 0; r < 4; ++r) {
      int owner = -1;
      for (int j = 0; j < B; ++j) {
        if (ln[j] <= 0 || P.band[j].w > 0 || (j & 3) != r) continue;
        const int lo = P.kmin + g * P.L, hi = lo + P.L;
        if (st[j] < hi && st[j] + ln[j] > lo) {
          if (owner >= 0) { printf("slab %d residue %d shared by bands %d and %d\n", g, r, owner, j); return 1; }
          owner = j;
        }
      }
    }
  std::vector<float> mags((size_t)TBF * MS, 0.f);
  srand(1234 + F + B);
  for (int t = 0; t < TBF; ++t)
    for (int k = 0; k < N; ++k) mags[(size_t)t * MS + k] = (float)rand() / RAND_MAX;
  const int pstride = P.NS * 128 * 4;
  std::vector<float> part((size_t)TBF * pstride, -1.f);
  for (int tid = 0; tid < 128; ++tid)
    b2::fb_slabs_dispatch<TBF, b2::MagLinear<MS>>(P.L, reinterpret_cast<const float4 *>(P.w4.data()), mags.data(), part.data(), P.NS,
                                   P.kmin, pstride, tid);
  double worst = 0;
  for (int j = 0; j < B; ++j) {
    float y[TBF];
    int4 bd;
    bd.x = P.band[j].x; bd.y = P.band[j].y; bd.z = P.band[j].z; bd.w = P.band[j].w;
    b2::fb_band_sum<TBF, b2::MagLinear<MS>>(bd, part.data(), pstride, mags.data(), P.dw.data(), y);
    for (int t = 0; t < TBF; ++t) {
      double ref = 0;
      for (int i = 0; i < ln[j]; ++i) ref += (double)w[wo[j] + i] * mags[(size_t)t * MS + st[j] + i];
      const double err = fabs(ref - y[t]) / (fabs(ref) + 1e-6);
      if (err > worst) worst = err;
    }
  }
  printf("F=%d B=%d: L=%d NS=%d kmin=%d kmax=%d direct=%d (%zu taps) max rel err %.2e\n", F, B, P.L, P.NS, P.kmin,
         P.kmax, P.ndirect, P.dw.size(), worst);
  return worst < 2e-6 ? 0 : 1;
}

int main(int argc, char **argv) {
  if (argc < 2) return 2;
  FILE *fh = fopen(argv[1], "r");
  if (!fh) return 2;
  int F, B, bad = 0, cases = 0;
  while (fscanf(fh, "%d %d", &F, &B) == 2) {
    std::vector<int> st(B), ln(B), wo(B);
    int nnz = 0;
    for (int j = 0; j < B; ++j) if (fscanf(fh, "%d", &st[j]) != 1) return 2;
    for (int j = 0; j < B; ++j) {
      if (fscanf(fh, "%d", &ln[j]) != 1) return 2;
      wo[j] = nnz;
      nnz += ln[j];
    }
    std::vector<float> w(nnz);
    for (int i = 0; i < nnz; ++i) if (fscanf(fh, "%f", &w[i]) != 1) return 2;
    switch (F) {
      case 1024: bad += run<1024>(B, st, ln, wo, w); break;
      case 2048: bad += run<2048>(B, st, ln, wo, w); break;
      case 4096: bad += run<4096>(B, st, ln, wo, w); break;
      case 8192: bad += run<8192>(B, st, ln, wo, w); break;
      default: return 2;
    }
    ++cases;
  }
  fclose(fh);
  if (bad || !cases) { printf("FAILED (%d of %d)\n", bad, cases); return 1; }
  printf("OK\n");
  return 0;
}
