// CPU emulation of the thread-mapped FFT passes in audio_tabs_b200/csrc/fft_core.cuh.
// Test infrastructure only: runs every "thread" of every pass sequentially (exactly the index
// mapping k_front uses, including the in-place pass 2) and compares the resulting half spectrum
// with a naive float64 DFT of the windowed frame.
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <complex>
#include "../../audio_tabs_b200/csrc/fft_core.cuh"

template <int F>
double run_one(unsigned seed) {
  using C = b2::FftCfg<F>;
  const double PI = 3.14159265358979323846;
  std::vector<float> x(F), w(F);
  srand(seed);
  for (int i = 0; i < F; ++i) {
    x[i] = (float)rand() / RAND_MAX * 2.f - 1.f;
    w[i] = (float)(0.5 * (0.5 - 0.5 * cos(2 * PI * i / (F - 1))));  // hanning * 1/2
  }
  std::vector<float2> tw2(16 * 16), tw3(C::TW3), pt(C::PT);
  for (int k1 = 0; k1 < 16; ++k1)
    for (int n2 = 0; n2 < 16; ++n2) {
      double a = -2 * PI * (n2 * k1) / 256.0;
      tw2[k1 * 16 + n2] = make_float2((float)cos(a), (float)sin(a));
    }
  for (int q = 0; q <= 128; ++q)
    for (int n3 = 0; n3 < C::R3; ++n3) {
      double a = -2 * PI * ((double)n3 * q) / C::N;
      tw3[n3 * 129 + q] = make_float2((float)cos(a), (float)sin(a));
    }
  for (int k3 = 0; k3 < C::R3; ++k3)
    for (int q = 0; q <= 128; ++q) {
      double a = -2 * PI * (q + 256.0 * k3) / F;
      pt[k3 * 129 + q] = make_float2((float)sin(a), (float)-cos(a));  // -i * exp(i a)
    }
  std::vector<float2> buf(C::BUF), X(C::N);
  std::vector<int> hits(C::N, 0);
  for (int b = 0; b < C::BPF; ++b) {
    auto load = [&](int n1) {
      int m = n1 * C::BPF + b;
      return make_float2(w[2 * m] * x[2 * m], w[2 * m + 1] * x[2 * m + 1]);
    };
    b2::fft_pass1<F>(load, buf.data() + b);
  }
  for (int t2 = 0; t2 < C::BPF; ++t2) {
    float2 t[16];
    for (int n2 = 0; n2 < 16; ++n2) t[n2] = tw2[(t2 & 15) * 16 + n2];
    b2::fft_pass2<F>(t, buf.data() + (t2 & 15) * C::S1 + (t2 >> 4));
  }
  auto emit = [&](int k, float2 v) { X[k] = v; hits[k]++; };
  std::vector<float2> wr(C::WR);
  for (int e = 0; e < C::WR; ++e) {
    double a = -2 * PI * e / C::WR;
    wr[e] = make_float2((float)cos(a), (float)sin(a));
  }
  if (seed & 1) {
    b2::fft_pass3_special<F>(buf.data(), tw3.data(), pt.data(), emit);     // single-thread form
  } else {
    for (int lane = 0; lane < 2 * C::R3; ++lane) {                        // lane-parallel form
      int bin;
      float2 v = b2::fft_pass3_selfpaired<F>(lane, buf.data(), wr.data(), pt.data(), bin);
      emit(bin, v);
    }
  }
  for (int u = 1; u < 128; ++u)
    b2::fft_pass3_unit<F>(u, buf.data() + b2::fft_col_offset<F>(u), buf.data() + b2::fft_col_offset<F>(256 - u),
                          tw3.data() + u, pt.data() + u, emit);
  double maxerr = 0, peak = 0;
  for (int k = 0; k < C::N; ++k) {
    if (hits[k] != 1) { printf("F=%d bin %d emitted %d times\n", F, k, hits[k]); return 1e9; }
    std::complex<double> s = 0;
    for (int n = 0; n < F; ++n) {
      double a = -2 * PI * ((double)((long long)n * k % F)) / F;
      s += 2.0 * (double)w[n] * (double)x[n] * std::complex<double>(cos(a), sin(a));
    }
    double e = std::abs(s - std::complex<double>(X[k].x, X[k].y));
    if (e > maxerr) maxerr = e;
    if (std::abs(s) > peak) peak = std::abs(s);
  }
  printf("F=%d max abs err %.3e, peak %.3e, rel-to-peak %.3e\n", F, maxerr, peak, maxerr / peak);
  return maxerr / peak;
}

int main() {
  double e = 0;
  for (unsigned seed = 1; seed <= 2; ++seed) {   // odd seeds: single-thread special, even: lane-parallel
    e = fmax(e, run_one<1024>(seed));
    e = fmax(e, run_one<2048>(seed));
    e = fmax(e, run_one<4096>(seed));
    e = fmax(e, run_one<8192>(seed));
  }
  if (e > 2e-6) { printf("FAIL\n"); return 1; }
  printf("OK\n");
  return 0;
}
