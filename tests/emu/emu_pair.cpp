// CPU emulation of the pair transform (two real frames as one complex FFT of F points) in
// audio_tabs_b200/csrc/fft_core.cuh: every "thread" of passes 1-3 with the index mapping k_front_pair
// uses, compared with float64 DFTs of the two windowed frames.  Test infrastructure only.
#include <cmath>
#include <complex>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../audio_tabs_b200/csrc/fft_core.cuh"

template <int F>
double run_one(unsigned seed) {
  constexpr int F2 = 2 * F;
  using C = b2::FftCfg<F2>;            // N = F complex points
  static_assert(C::N == F, "pair geometry");
  const double PI = 3.14159265358979323846;
  std::vector<float> xa(F), xb(F), w(F);
  srand(seed);
  for (int i = 0; i < F; ++i) {
    xa[i] = (float)rand() / RAND_MAX * 2.f - 1.f;
    xb[i] = (float)rand() / RAND_MAX * 2.f - 1.f;
    w[i] = (float)(0.5 * (0.5 - 0.5 * cos(2 * PI * i / (F - 1))));   // hanning * 1/2
  }
  std::vector<float2> tw2(256), tw3(C::TW3C), wr(C::WR);   // tw3: compact form, row r = W_N^(2^r q)
  for (int k1 = 0; k1 < 16; ++k1)
    for (int n2 = 0; n2 < 16; ++n2) {
      double a = -2 * PI * (n2 * k1) / 256.0;
      tw2[k1 * 16 + n2] = make_float2((float)cos(a), (float)sin(a));
    }
  for (int q = 0; q <= 128; ++q)
    for (int r = 0; r < C::LOG2R3; ++r) {
      double a = -2 * PI * ((double)(1 << r) * q) / C::N;
      tw3[r * 129 + q] = make_float2((float)cos(a), (float)sin(a));
    }
  for (int e = 0; e < C::WR; ++e) {
    double a = -2 * PI * e / C::WR;
    wr[e] = make_float2((float)cos(a), (float)sin(a));
  }
  std::vector<float2> buf(C::BUF);
  std::vector<float> MA(F / 2 + 16, -1.f), MB(F / 2 + 16, -1.f);
  for (int b = 0; b < C::BPF; ++b) {
    auto load = [&](int n1) {
      int n = n1 * C::BPF + b;
      return make_float2(w[n] * xa[n], w[n] * xb[n]);
    };
    b2::fft_pass1<F2>(load, buf.data() + b);
  }
  for (int t2 = 0; t2 < C::BPF; ++t2) {
    float2 t[16];
    for (int n2 = 0; n2 < 16; ++n2) t[n2] = tw2[(t2 & 15) * 16 + n2];
    b2::fft_pass2<F2>(t, buf.data() + (t2 & 15) * C::S1 + (t2 >> 4));
  }
  auto mag = [](float2 v) { return sqrtf(v.x * v.x + v.y * v.y); };
  for (int u = 0; u < 128; ++u)
    b2::fft_pair_pass3_unit<F2>(u, buf.data() + b2::fft_col_offset<F2>(u),
                                buf.data() + b2::fft_col_offset<F2>((256 - u) & 255), tw3.data() + u,
                                [&](int bin, float2 p, float2 m) { MA[bin] = mag(p); MB[bin] = mag(m); });
  std::vector<float2> Z(C::R3);
  for (int k3 = 0; k3 < C::R3; ++k3) Z[k3] = b2::fft_pair_col128<F2>(k3, buf.data(), wr.data());
  for (int j = 0; j < C::R3 / 2; ++j) {
    float2 a = Z[j], b = Z[C::R3 - 1 - j];
    MA[128 + 256 * j] = mag(make_float2(a.x + b.x, a.y - b.y));
    MB[128 + 256 * j] = mag(make_float2(a.x - b.x, a.y + b.y));
  }
  double maxerr = 0, peak = 0;
  for (int k = 0; k < F / 2; ++k) {
    std::complex<double> sa = 0, sb = 0;
    for (int n = 0; n < F; ++n) {
      double ang = -2 * PI * ((double)((long long)n * k % F)) / F;
      std::complex<double> e(cos(ang), sin(ang));
      sa += 2.0 * (double)w[n] * (double)xa[n] * e;
      sb += 2.0 * (double)w[n] * (double)xb[n] * e;
    }
    if (MA[k] < 0 || MB[k] < 0) { printf("F=%d bin %d not emitted\n", F, k); return 1e9; }
    maxerr = fmax(maxerr, fmax(fabs(std::abs(sa) - MA[k]), fabs(std::abs(sb) - MB[k])));
    peak = fmax(peak, fmax(std::abs(sa), std::abs(sb)));
  }
  printf("pair F=%d max abs err %.3e, peak %.3e, rel-to-peak %.3e\n", F, maxerr, peak, maxerr / peak);
  return maxerr / peak;
}

int main() {
  double e = 0;
  for (unsigned seed = 1; seed <= 2; ++seed) {
    e = fmax(e, run_one<1024>(seed));
    e = fmax(e, run_one<2048>(seed));
    e = fmax(e, run_one<4096>(seed));
  }
  if (e > 2e-6) { printf("FAIL\n"); return 1; }
  printf("OK\n");
  return 0;
}
