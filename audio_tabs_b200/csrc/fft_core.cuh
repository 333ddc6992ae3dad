// fft_core.cuh -- register-level butterflies and the three thread-mapped passes of the
// frame FFT used by every front-end kernel.
//
// Replaces the inner loop of madmom.audio.stft.stft() (per-frame scipy.fftpack.fft, reached from
// /root/reference/backend/app/services/grid/beats.py:74 via RNNBeatProcessor).
//
// A real frame of F samples is transformed as ONE complex FFT of N = F/2 points on
// z[m] = x[2m] + i x[2m+1] (window folded into the load, pre-scaled by 1/2), followed by the
// even/odd split  X[k] = E[k] + W_F^k O[k].  N = 16 * 16 * R3 with R3 = F/512 in {2,4,8,16}:
//
//   pass 1  thread b in [0,16*R3):  DFT16 over n1 of z[n1*16*R3 + b]            -> buf[k1][b]
//   pass 2  thread (k1,n3):         twiddle W_256^(n2*k1), DFT16 over n2, IN PLACE (k2 replaces n2)
//   pass 3  unit u in [0,128):      butterflies q=u and q=256-u of radix R3 are combined BEFORE the
//           DFT (P = a + conj(b), M = a - conj(b)), so E and O come out of two DFT_R3 directly and
//           both bins k and N-k of the real spectrum are produced in registers.
//
// Everything here is __host__ __device__ so tests/emu can run the exact thread mapping on the CPU
// (test infrastructure only; the product path is the CUDA build).
#pragma once

#if defined(__CUDACC__)
#define B2_HD __host__ __device__ __forceinline__
#else
#include <cmath>
#define B2_HD inline
struct float2 {
  float x, y;
};
static inline float2 make_float2(float x, float y) {
  float2 r;
  r.x = x;
  r.y = y;
  return r;
}
#endif

namespace b2 {

constexpr int kGroupThreads = 128;  // threads cooperating on one frame (or FPG small frames)

template <int F>
struct FftCfg {
  static_assert(F == 1024 || F == 2048 || F == 4096 || F == 8192, "unsupported frame size");
  static constexpr int N = F / 2;          // complex FFT length
  static constexpr int R3 = N / 256;       // radix of the last pass
  static constexpr int BPF = 16 * R3;      // radix-16 butterflies per frame in pass 1 and pass 2
  static constexpr int FPG = (BPF >= kGroupThreads) ? 1 : kGroupThreads / BPF;  // frames per group step
  static constexpr int IT12 = (BPF >= kGroupThreads) ? BPF / kGroupThreads : 1; // pass-1/2 butterflies per thread
  static constexpr int S1 = BPF + 1;       // buf1 row stride (float2), +1 keeps pass-2 reads conflict free
  static constexpr int BUF = 16 * S1;      // float2 elements of the single in-place buffer of a frame
  static constexpr int TW3 = 129 * R3;     // tw3[n3*129 + q] = W_N^(n3*q), q in [0,128]
  static constexpr int LOG2R3 = (R3 == 2) ? 1 : (R3 == 4) ? 2 : (R3 == 8) ? 3 : 4;
  static constexpr int TW3C = 129 * LOG2R3; // compact form (pair kernels): row r holds n3 = 2^r -- the only rows tw3_rows reads
  static constexpr int PT = 129 * R3;      // pt[k3*129 + q] = -i * W_F^(q + 256*k3)
  static constexpr int WR = 2 * R3;        // wr[e] = W_(2 R3)^e, used by the self-paired columns
  // frames per tail batch (filterbank/log/diff stage): 4, 4, 2, 1 -- measured best on B200 (profiles/README.md)
#if defined(B2_TB_1024) && defined(B2_TB_2048) && defined(B2_TB_4096)   // tuning override
  static constexpr int TB = (F == 1024) ? B2_TB_1024 : (F == 2048) ? B2_TB_2048 : (F == 4096) ? B2_TB_4096 : 1;
#elif defined(B2_TAIL_STEPS)
  static constexpr int TB = (F == 8192) ? 1 : B2_TAIL_STEPS * FPG;
#else
  static constexpr int TB = (F == 1024) ? 4 : (F == 2048) ? 4 : (F == 4096) ? 2 : 1;
#endif
  static constexpr int TBF = TB < 4 ? TB : 4;                        // frames per filterbank / band-stage call
  static constexpr int MS = N + 16;        // floats per frame in the magnitude buffer (+ room for the last slab)
};

// ---- complex helpers ---------------------------------------------------------------------------
// On sm_100 every complex add / subtract / rotate-by-i is ONE packed-FP32 instruction on a 64-bit
// register pair and a complex multiply is TWO (FMUL2 + FFMA2): the packed instructions take a
// half-swap (LO_HI), a per-half negation (NP / PN) and a scalar broadcast (.F32) as operand
// modifiers, and ptxas folds the component shuffles written below into them (checked with
// cuobjdump: no MOV / PRMT is left).  The host build (tests/emu) uses the scalar form.
#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ >= 1000) && !defined(B2_NO_PACKED_F32)
#define B2_PACKED_F32 1
B2_HD float2 cadd(float2 a, float2 b) { return __fadd2_rn(a, b); }
B2_HD float2 csub(float2 a, float2 b) { return __fadd2_rn(a, make_float2(-b.x, -b.y)); }
B2_HD float2 cadd_conj(float2 a, float2 b) { return __fadd2_rn(a, make_float2(b.x, -b.y)); }   // a + conj(b)
B2_HD float2 csub_conj(float2 a, float2 b) { return __fadd2_rn(a, make_float2(-b.x, b.y)); }   // a - conj(b)
B2_HD float2 cadd_negi(float2 a, float2 b) { return __fadd2_rn(a, make_float2(b.y, -b.x)); }   // a - i b
B2_HD float2 cadd_posi(float2 a, float2 b) { return __fadd2_rn(a, make_float2(-b.y, b.x)); }   // a + i b
B2_HD float2 emul(float2 a, float2 b) { return __fmul2_rn(a, b); }                               // element-wise
B2_HD float2 crscale(float2 a, float s) { return __fmul2_rn(a, make_float2(s, s)); }
B2_HD float2 cmul(float2 a, float2 b) {
  return __ffma2_rn(make_float2(a.y, a.x), make_float2(-b.y, b.y), __fmul2_rn(a, make_float2(b.x, b.x)));
}
#else
B2_HD float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
B2_HD float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
B2_HD float2 cadd_conj(float2 a, float2 b) { return make_float2(a.x + b.x, a.y - b.y); }
B2_HD float2 csub_conj(float2 a, float2 b) { return make_float2(a.x - b.x, a.y + b.y); }
B2_HD float2 cadd_negi(float2 a, float2 b) { return make_float2(a.x + b.y, a.y - b.x); }
B2_HD float2 cadd_posi(float2 a, float2 b) { return make_float2(a.x - b.y, a.y + b.x); }
B2_HD float2 emul(float2 a, float2 b) { return make_float2(a.x * b.x, a.y * b.y); }
B2_HD float2 crscale(float2 a, float s) { return make_float2(a.x * s, a.y * s); }
B2_HD float2 cmul(float2 a, float2 b) {
  return make_float2(fmaf(-a.y, b.y, a.x * b.x), fmaf(a.x, b.y, a.y * b.x));
}
#endif
B2_HD float2 cconj(float2 a) { return make_float2(a.x, -a.y); }
// Store of a complex value into SHARED memory.  Written as a plain float2 assignment ptxas copies
// every result of a packed instruction into a staging pair first (MOV, MOV, STS.64 -- checked with
// cuobjdump on CUDA 12.9); the explicit st.shared.v2.f32 takes the pair the arithmetic left it in.
B2_HD void cstore(float2 *p, float2 v) {
#if defined(B2_PACKED_F32)
  asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"((unsigned)__cvta_generic_to_shared(p)), "f"(v.x), "f"(v.y)
               : "memory");
#else
  *p = v;
#endif
}
B2_HD float2 mul_neg_i(float2 a) { return make_float2(a.y, -a.x); }  // a * (-i)

constexpr float kH = 0.70710678118654752440f;   // sqrt(1/2)
constexpr float kC1 = 0.92387953251128675613f;  // cos(pi/8)
constexpr float kS1 = 0.38268343236508977173f;  // sin(pi/8)

B2_HD float2 mul_w8_1(float2 a) { return crscale(cadd_negi(a, a), kH); }    // * (h,-h):  h (a - i a)
B2_HD float2 mul_w8_3(float2 a) { return crscale(cadd_posi(a, a), -kH); }   // * (-h,-h): -h (a + i a)

// ---- in-register DFTs (forward, W = exp(-2 pi i / R)) ------------------------------------------
B2_HD void dft2(float2 &a, float2 &b) {
  float2 t = csub(a, b);
  a = cadd(a, b);
  b = t;
}

// natural order in, natural order out
B2_HD void dft4(float2 &x0, float2 &x1, float2 &x2, float2 &x3) {
  float2 t0 = cadd(x0, x2), t1 = csub(x0, x2);
  float2 t2 = cadd(x1, x3), d = csub(x1, x3);   // t3 = -i d, folded into the two adds below
  x0 = cadd(t0, t2);
  x2 = csub(t0, t2);
  x1 = cadd_negi(t1, d);
  x3 = cadd_posi(t1, d);
}

// in place; X[k] ends up at v[pos8(k)]
B2_HD void dft8(float2 (&v)[8]) {
  dft4(v[0], v[2], v[4], v[6]);  // n2 = 0: y[0][k1] at v[2*k1]
  dft4(v[1], v[3], v[5], v[7]);  // n2 = 1: y[1][k1] at v[2*k1+1]
  v[3] = mul_w8_1(v[3]);
  v[5] = mul_neg_i(v[5]);
  v[7] = mul_w8_3(v[7]);
  dft2(v[0], v[1]);
  dft2(v[2], v[3]);
  dft2(v[4], v[5]);
  dft2(v[6], v[7]);
}

// in place; X[k] ends up at v[pos16(k)]
B2_HD void dft16(float2 (&v)[16]) {
#pragma unroll
  for (int n2 = 0; n2 < 4; ++n2) dft4(v[n2], v[4 + n2], v[8 + n2], v[12 + n2]);  // y[n2][k1] at v[4*k1+n2]
  // twiddle y[n2][k1] *= W16^(n2*k1)
  v[5] = cmul(v[5], make_float2(kC1, -kS1));    // (1,1) W^1
  v[9] = mul_w8_1(v[9]);                        // (n2=1,k1=2) W^2
  v[13] = cmul(v[13], make_float2(kS1, -kC1));  // (1,3) W^3
  v[6] = mul_w8_1(v[6]);                        // (2,1) W^2
  v[10] = mul_neg_i(v[10]);                     // (2,2) W^4
  v[14] = mul_w8_3(v[14]);                      // (2,3) W^6
  v[7] = cmul(v[7], make_float2(kS1, -kC1));    // (3,1) W^3
  v[11] = mul_w8_3(v[11]);                      // (3,2) W^6
  v[15] = cmul(v[15], make_float2(-kC1, kS1));  // (3,3) W^9
#pragma unroll
  for (int k1 = 0; k1 < 4; ++k1) dft4(v[4 * k1], v[4 * k1 + 1], v[4 * k1 + 2], v[4 * k1 + 3]);
}

// index in v[] that holds output bin k after the in-place DFT of radix R
template <int R>
B2_HD constexpr int dft_pos(int k) {
  return R == 16 ? (((k & 3) << 2) | (k >> 2)) : R == 8 ? (2 * (k & 3) + (k >> 2)) : k;
}

template <int R>
struct Dft;
template <>
struct Dft<2> {
  static B2_HD void run(float2 (&v)[2]) { dft2(v[0], v[1]); }
};
template <>
struct Dft<4> {
  static B2_HD void run(float2 (&v)[4]) { dft4(v[0], v[1], v[2], v[3]); }
};
template <>
struct Dft<8> {
  static B2_HD void run(float2 (&v)[8]) { dft8(v); }
};
template <>
struct Dft<16> {
  static B2_HD void run(float2 (&v)[16]) { dft16(v); }
};

// ---- pass 1: windowed load + DFT16 over n1 -----------------------------------------------------
// One in-place buffer per frame: buf[k1 * S1 + j], j in [0, BPF).  `load(n1)` returns the windowed
// complex sample z[n1 * BPF + b]; `p` = buf + b.
template <int F, class Load>
B2_HD void fft_pass1(Load load, float2 *p) {
  using C = FftCfg<F>;
  float2 v[16];
#pragma unroll
  for (int n1 = 0; n1 < 16; ++n1) v[n1] = load(n1);
  dft16(v);
#pragma unroll
  for (int k1 = 0; k1 < 16; ++k1) cstore(p + k1 * C::S1, v[dft_pos<16>(k1)]);
}

// two pass-1 butterflies of one thread with all 32 (pairs of) loads issued before the arithmetic
template <int F, class LoadA, class LoadB>
B2_HD void fft_pass1_x2(LoadA load_a, LoadB load_b, float2 *p, int d) {
  using C = FftCfg<F>;
  float2 v[16], w[16];
#pragma unroll
  for (int n1 = 0; n1 < 16; ++n1) v[n1] = load_a(n1);
#pragma unroll
  for (int n1 = 0; n1 < 16; ++n1) w[n1] = load_b(n1);
  dft16(v);
  dft16(w);
#pragma unroll
  for (int k1 = 0; k1 < 16; ++k1) cstore(p + k1 * C::S1, v[dft_pos<16>(k1)]);
#pragma unroll
  for (int k1 = 0; k1 < 16; ++k1) cstore(p + d + k1 * C::S1, w[dft_pos<16>(k1)]);
}

// ---- pass 2: twiddle W_256^(n2*k1), DFT16 over n2, in place ------------------------------------
// thread (k1, n3): p = buf + k1 * S1 + n3; element n2 sits at p[n2 * R3]; output k2 replaces it.
template <int F>
B2_HD void fft_pass2(const float2 (&tw2)[16], float2 *p) {
  using C = FftCfg<F>;
  float2 v[16];
#pragma unroll
  for (int n2 = 0; n2 < 16; ++n2) v[n2] = p[n2 * C::R3];
#pragma unroll
  for (int n2 = 1; n2 < 16; ++n2) v[n2] = cmul(v[n2], tw2[n2]);
  dft16(v);
#pragma unroll
  for (int k2 = 0; k2 < 16; ++k2) cstore(p + k2 * C::R3, v[dft_pos<16>(k2)]);
}

// two independent pass-2 butterflies of one thread (p and p + d) with their loads issued together: twice the
// instruction-level parallelism where a thread has two to do (frame 4096 of the pair kernel)
template <int F>
B2_HD void fft_pass2_x2(const float2 (&tw2)[16], float2 *p, int d) {
  using C = FftCfg<F>;
  float2 v[16], w[16];
#pragma unroll
  for (int n2 = 0; n2 < 16; ++n2) v[n2] = p[n2 * C::R3];
#pragma unroll
  for (int n2 = 0; n2 < 16; ++n2) w[n2] = p[d + n2 * C::R3];
#pragma unroll
  for (int n2 = 1; n2 < 16; ++n2) v[n2] = cmul(v[n2], tw2[n2]);
#pragma unroll
  for (int n2 = 1; n2 < 16; ++n2) w[n2] = cmul(w[n2], tw2[n2]);
  dft16(v);
  dft16(w);
#pragma unroll
  for (int k2 = 0; k2 < 16; ++k2) cstore(p + k2 * C::R3, v[dft_pos<16>(k2)]);
#pragma unroll
  for (int k2 = 0; k2 < 16; ++k2) cstore(p + d + k2 * C::R3, w[dft_pos<16>(k2)]);
}

// column q = k1 + 16*k2 of the pass-2 output starts at this offset (elements n3 = 0..R3-1 follow)
template <int F>
B2_HD constexpr int fft_col_offset(int q) {
  return (q & 15) * FftCfg<F>::S1 + (q >> 4) * FftCfg<F>::R3;
}

// Twiddles of the last pass: W^(n3 q), n3 = 1..R3-1, from the table rows n3 = 1, 2, 4, 8 and R3-5 complex
// products -- shared-memory wavefronts, not FP32 issue slots, are the scarcer resource in these kernels
// (profiles/README.md); the extra rounding (<= 3 products deep) is ~2e-7 relative.  -DB2_TW3_TABLE reads
// all R3-1 rows instead.
// COMPACT: the table holds only the rows n3 = 1, 2, 4, 8 (row r = log2 n3), which is all this routine reads --
// the pair kernels keep it that way in shared memory (FftCfg::TW3C).
template <int R3, bool COMPACT = false>
B2_HD void tw3_rows(const float2 *tw3q, float2 (&w)[R3]) {
#if defined(B2_TW3_TABLE)
  if (!COMPACT) {
#pragma unroll
    for (int n3 = 1; n3 < R3; ++n3) w[n3] = tw3q[n3 * 129];
    return;
  }
#endif
  {
    int r = 0;
#pragma unroll
    for (int n3 = 1; n3 < R3; n3 *= 2, ++r) w[n3] = tw3q[(COMPACT ? r : n3) * 129];
  }
#pragma unroll
  for (int n3 = 3; n3 < R3; ++n3)
    if (n3 & (n3 - 1)) {
      const int hi = n3 >= 8 ? 8 : n3 >= 4 ? 4 : 2;
      w[n3] = cmul(w[hi], w[n3 - hi]);
    }
}

// cos / sin of j pi / 16, j in [0,16): W_(2 R3)^k3 = (cos_pi16(j), -sin_pi16(j)) with j = k3 * 16 / R3
// (functions, not namespace-scope tables: the index is a constant only after unrolling)
B2_HD constexpr float cos_pi16(int j) {
  constexpr float t[16] = {1.f, 0.98078528040323044913f, 0.92387953251128675613f, 0.83146961230254523708f,
                           0.70710678118654752440f, 0.55557023301960222474f, 0.38268343236508977173f,
                           0.19509032201612826785f, 0.f, -0.19509032201612826785f, -0.38268343236508977173f,
                           -0.55557023301960222474f, -0.70710678118654752440f, -0.83146961230254523708f,
                           -0.92387953251128675613f, -0.98078528040323044913f};
  return t[j];
}
B2_HD constexpr float sin_pi16(int j) { return j < 8 ? cos_pi16(8 - j) : cos_pi16(j - 8); }

// ---- pass 3: last radix-R3 pass fused with the real-spectrum split -----------------------------
// unit u in [1,127]: pa / pb = columns u and 256-u; tw3u = tw3 + u (tw3[n3*129 + q] = W_N^(n3 q));
// ptu = pt + u (pt[k3*129 + q] = -i W_F^(q + 256 k3)).  Emits bins k = u + 256*k3 and N - k.
template <int F, class Emit>
B2_HD void fft_pass3_unit(int u, const float2 *pa, const float2 *pb, const float2 *tw3u, const float2 *ptu,
                          Emit emit) {
  using C = FftCfg<F>;
  constexpr int R3 = C::R3;
  float2 P[R3], M[R3];
#pragma unroll
  for (int n3 = 0; n3 < R3; ++n3) {
    float2 a = pa[n3];
    float2 b = pb[n3];
    P[n3] = cadd_conj(a, b);  // a + conj(b)
    M[n3] = csub_conj(a, b);  // a - conj(b)
  }
  {
    float2 w[R3];
    tw3_rows<R3>(tw3u, w);
#pragma unroll
    for (int n3 = 1; n3 < R3; ++n3) {
      P[n3] = cmul(P[n3], w[n3]);
      M[n3] = cmul(M[n3], w[n3]);
    }
  }
  Dft<R3>::run(P);
  Dft<R3>::run(M);
  const float2 pt0 = ptu[0];   // pt[k3*129 + u] = pt[u] * W_(2 R3)^k3: one load, the rest are constants
#pragma unroll
  for (int k3 = 0; k3 < R3; ++k3) {
    float2 E = P[dft_pos<R3>(k3)];
#if defined(B2_TW3_TABLE)
    float2 T = cmul(M[dft_pos<R3>(k3)], ptu[k3 * 129]);
#else
    float2 T = M[dft_pos<R3>(k3)];
    if (k3 > 0) T = cmul(T, make_float2(cos_pi16(k3 * (16 / R3)), -sin_pi16(k3 * (16 / R3))));
    T = cmul(T, pt0);
#endif
    const int k = u + 256 * k3;
    emit(k, cadd(E, T));
    emit(C::N - k, cconj(csub(E, T)));
  }
}

// the two self-paired butterflies q = 0 (bins 256*k3) and q = 128 (bins 128 + 256*k3)
template <int F, class Emit>
B2_HD void fft_pass3_special(const float2 *buf, const float2 *tw3, const float2 *pt, Emit emit) {
  using C = FftCfg<F>;
  constexpr int R3 = C::R3;
  float2 Z[R3];
  const float2 *p0 = buf + fft_col_offset<F>(0);
#pragma unroll
  for (int n3 = 0; n3 < R3; ++n3) Z[n3] = p0[n3];
  Dft<R3>::run(Z);
#pragma unroll
  for (int k3 = 0; k3 < R3; ++k3) {
    float2 a = Z[dft_pos<R3>(k3)];
    float2 b = cconj(Z[dft_pos<R3>((R3 - k3) % R3)]);
    emit(256 * k3, cadd(cadd(a, b), cmul(csub(a, b), pt[k3 * 129])));
  }
  const float2 *p1 = buf + fft_col_offset<F>(128);
#pragma unroll
  for (int n3 = 0; n3 < R3; ++n3) {
    float2 a = p1[n3];
    Z[n3] = n3 == 0 ? a : cmul(a, tw3[n3 * 129 + 128]);
  }
  Dft<R3>::run(Z);
#pragma unroll
  for (int k3 = 0; k3 < R3; ++k3) {
    float2 a = Z[dft_pos<R3>(k3)];
    float2 b = cconj(Z[dft_pos<R3>(R3 - 1 - k3)]);
    emit(128 + 256 * k3, cadd(cadd(a, b), cmul(csub(a, b), pt[k3 * 129 + 128])));
  }
}

// Lane-parallel form of the two self-paired columns (what k_front runs: 2*R3 lanes of one warp,
// one output bin per lane, no second code path for a whole thread):
//   h = lane / R3 (0: q = 0, 1: q = 128), k3 = lane % R3, bin k = 128 h + 256 k3
//   X[k] = E + pt[k] D,  E = 2 sum_n w_n Re(c_n),  D = 2 i sum_n w_n Im(c_n),
//   w_n = W_(2 R3)^(n (2 k3 + h)) = wr[(n (2 k3 + h)) mod 2 R3],  c_n = element n of column q.
// mirror = X[N - k] = conj(E - pt[k] D): for lane 0 (k = 0) the Nyquist bin X[N].
template <int F>
B2_HD float2 fft_pass3_selfpaired(int lane, const float2 *buf, const float2 *wr, const float2 *pt, int &bin, float2 &mirror) {
  using C = FftCfg<F>;
  constexpr int R3 = C::R3;
  const int h = lane / R3, k3 = lane % R3;
  const float2 *c = buf + (h ? fft_col_offset<F>(128) : fft_col_offset<F>(0));
  const int step = 2 * k3 + h;
  int e = 0;
  float2 E = make_float2(0.f, 0.f), D = make_float2(0.f, 0.f);
#pragma unroll
  for (int n = 0; n < R3; ++n) {
    const float2 w = wr[e];
    const float2 cn = c[n];
    E.x = fmaf(w.x, cn.x, E.x);
    E.y = fmaf(w.y, cn.x, E.y);
    D.x = fmaf(w.x, cn.y, D.x);
    D.y = fmaf(w.y, cn.y, D.y);
    e = (e + step) & (2 * R3 - 1);
  }
  const float2 T = cmul(make_float2(-2.f * D.y, 2.f * D.x), pt[k3 * 129 + 128 * h]);
  bin = 128 * h + 256 * k3;
  mirror = make_float2(fmaf(2.f, E.x, -T.x), fmaf(-2.f, E.y, T.y));
  return make_float2(fmaf(2.f, E.x, T.x), fmaf(2.f, E.y, T.y));
}

template <int F>
B2_HD float2 fft_pass3_selfpaired(int lane, const float2 *buf, const float2 *wr, const float2 *pt, int &bin) {
  float2 mirror;
  return fft_pass3_selfpaired<F>(lane, buf, wr, pt, bin, mirror);
}

// =================================================================================================
// Pair transform: TWO real frames of F samples as ONE complex FFT of F points, z[n] = xA[n] + i xB[n]
// (window folded into the load, pre-scaled by 1/2).  Geometry = FftCfg<2F> (N = F = 16*16*R3).
//   XA[k] = (Z[k] + conj Z[F-k]) / 2,   XB[k] = -i (Z[k] - conj Z[F-k]) / 2,   k in [0, F/2)
// With the same column pairing as above (q and 256-q combined BEFORE the last radix: P = a + conj b,
// M = a - conj b) the two DFT_R3 deliver XA and i*XB directly: no split twiddles, and every one of the
// 2*R3 outputs is a needed bin -- k3 < R3/2 gives bin q + 256 k3, k3 >= R3/2 gives the conjugate of bin
// (256-q) + 256 (R3-1-k3).  Only magnitudes are emitted (the fused log-filtered path).
// -------------------------------------------------------------------------------------------------
// unit u in [0,127]: pa / pb = columns u and (256-u)&255 (u = 0 pairs column 0 with itself and simply
// emits bins 256 j twice, plus bin F/2 which lands in the padding of the magnitude buffer).
// tw3u = compact twiddle table + u: tw3c[r*129 + q] = W_N^(2^r q), r < log2 R3.
// emit(bin, |XA[bin]|, |XB[bin]|) -- the caller takes the magnitudes.
template <int F2, class Emit>
B2_HD void fft_pair_pass3_unit(int u, const float2 *pa, const float2 *pb, const float2 *tw3u, Emit emit) {
  using C = FftCfg<F2>;
  constexpr int R3 = C::R3;
  float2 P[R3], M[R3];
#pragma unroll
  for (int n3 = 0; n3 < R3; ++n3) {
    float2 a = pa[n3];
    float2 b = pb[n3];
    P[n3] = cadd_conj(a, b);
    M[n3] = csub_conj(a, b);
  }
  {
    float2 w[R3];
    tw3_rows<R3, true>(tw3u, w);     // tw3u: compact table (rows n3 = 1, 2, 4, 8)
#pragma unroll
    for (int n3 = 1; n3 < R3; ++n3) {
      P[n3] = cmul(P[n3], w[n3]);
      M[n3] = cmul(M[n3], w[n3]);
    }
  }
  Dft<R3>::run(P);
  Dft<R3>::run(M);
#pragma unroll
  for (int k3 = 0; k3 < R3; ++k3) {
    const int bin = k3 < R3 / 2 ? u + 256 * k3 : ((256 - u) + 256 * (R3 - 1 - k3));
    emit(bin, P[dft_pos<R3>(k3)], M[dft_pos<R3>(k3)]);
  }
}

// The self-paired column q = 128 (bins 128 + 256 j), lane-parallel on R3 lanes of one warp:
// lane k3 forms Z[128 + 256 k3] = sum_n c[n] W_(2 R3)^(n (2 k3 + 1)); the caller exchanges Z with lane
// R3-1-k3 (the mirror bin) and emits |Z + conj Z'| (frame A) and |Z - conj Z'| (frame B) for k3 < R3/2.
// The sum is cut into kCol128Parts<F2> parts of consecutive terms so that the kernel can spread it over
// R3 * parts lanes (part = lane / R3) and add the parts with xor shuffles: the warp that owns column 128 is
// the last one to reach the barrier after pass 3, and this is the extra work it does.  Measured on B200: two
// parts at R3 = 16 (frame 4096 of the pair kernel) 6.76 -> 6.71 ms; four parts at R3 = 8 / 4 LOSE 2.4 % / 1.2 %
// (the sums are short there and the shuffles cost more than they save), so those stay on one lane per bin.
template <int F2>
constexpr int kCol128Parts = FftCfg<F2>::R3 >= 16 ? 2 : 1;

template <int F2>
B2_HD float2 fft_pair_col128_part(int k3, int part, const float2 *buf, const float2 *wr) {
  using C = FftCfg<F2>;
  constexpr int R3 = C::R3, TPP = R3 / kCol128Parts<F2>;   // terms per part
  static_assert(TPP >= 1, "more parts than terms");
  const float2 *c = buf + fft_col_offset<F2>(128) + part * TPP;
  const int step = 2 * k3 + 1;
  int e = (part * TPP * step) & (2 * R3 - 1);
  float2 Z = make_float2(0.f, 0.f);
#pragma unroll
  for (int n = 0; n < TPP; ++n) {
    const float2 w = wr[e];
    const float2 cn = c[n];
    Z.x = fmaf(w.x, cn.x, fmaf(-w.y, cn.y, Z.x));
    Z.y = fmaf(w.x, cn.y, fmaf(w.y, cn.x, Z.y));
    e = (e + step) & (2 * R3 - 1);
  }
  return Z;
}

template <int F2>
B2_HD float2 fft_pair_col128(int k3, const float2 *buf, const float2 *wr) {   // all parts on one lane (tests/emu)
  float2 Z = make_float2(0.f, 0.f);
  for (int part = 0; part < kCol128Parts<F2>; ++part) {
    const float2 z = fft_pair_col128_part<F2>(k3, part, buf, wr);
    Z.x += z.x;
    Z.y += z.y;
  }
  return Z;
}

}  // namespace b2
