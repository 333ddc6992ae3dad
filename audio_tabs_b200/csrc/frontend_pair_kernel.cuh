// frontend_pair_kernel.cuh -- the fused log-filtered front end with the PAIR transform: two consecutive
// real frames of F samples are carried through ONE complex FFT of F points (z = xA + i xB, see the pair
// section of fft_core.cuh), which removes the even/odd split twiddles, halves the window loads and halves
// the group barriers per frame compared with k_front (frontend_kernel.cuh).  Same execution model, same
// tables (plus the F-point twiddles), same tail (front_tail).  MODE_LOGFILT only: the raw STFT keeps k_front.
//
// Replaces, for one resolution, the madmom 0.16.1 chain reached from
// /root/reference/backend/app/services/grid/beats.py:74 (RNNBeatProcessor):
//   signal_frame -> frame*fft_window -> fftpack.fft[:F/2] -> np.abs -> np.dot(., filterbank)
//   -> np.log10(mul*y+add) -> SpectrogramDifference(positive) -> np.hstack
#pragma once
#include <type_traits>

#include "frontend_kernel.cuh"

namespace b2 {

template <int F>
struct PairCfg {
  static_assert(F == 1024 || F == 2048 || F == 4096, "pair transform: frame sizes 1024, 2048, 4096");
  using C2 = FftCfg<2 * F>;                       // geometry of the F-point complex FFT
  static constexpr int PPS = C2::FPG;             // pairs per group step
  static constexpr int FPS = 2 * PPS;             // frames per group step
  static constexpr int TB = FftCfg<F>::TB >= FPS ? FftCfg<F>::TB : FPS;   // frames per tail batch (multiple of FPS)
#ifdef B2_PAIR_TBF_2048   // tuning override: frames per filterbank / band-stage call at frame 2048
  static constexpr int TBF = (F == 2048) ? B2_PAIR_TBF_2048 : (TB < 4 ? TB : 4);
#else
  static constexpr int TBF = TB < 4 ? TB : 4;
#endif
  static constexpr int MS = FftCfg<F>::MS;        // floats per frame in the magnitude buffer
  // Frame 4096: a 33 KB FFT buffer per pair plus 16 KB of magnitudes allow three groups per SM only.
  // -DB2_PAIR_INPLACE makes pass 3 write the magnitudes IN PLACE of the columns it has consumed
  // (MagInPlace) so that four groups fit; measured on B200 it loses (7.67 ms against 7.54 ms for three
  // groups with 163 registers: the scattered addressing and the extra barrier cost more than the fourth
  // group brings), so it is off by default.
#ifdef B2_PAIR_INPLACE
  static constexpr bool INPLACE = (F == 4096);
#else
  static constexpr bool INPLACE = false;
#endif
#ifdef B2_PAIR_MAG_LINEAR   // tuning: one row of MS floats per frame instead of frames interleaved per bin
  using MA = typename std::conditional<INPLACE, MagInPlace<2 * F, MS>, MagLinear<MS>>::type;
  static constexpr bool INTERLEAVED = false;
  static constexpr bool PLANES = false;
#else
  static constexpr bool INTERLEAVED = !INPLACE && (TB == TBF) && (TB == 2 || TB == 4);
#ifdef B2_PAIR_MAG_PLANES   // four-frame batches: one plane per pair of frames (conflict-free pass-3 stores;
                            // measured slower on B200: 2.14 / 3.56 against 2.09 / 3.55 ms at frames 1024 / 2048)
  static constexpr bool PLANES = INTERLEAVED && TB == 4;
#else
  static constexpr bool PLANES = false;
#endif
  using MA = typename std::conditional<INPLACE, MagInPlace<2 * F, MS>,
             typename std::conditional<PLANES, MagPairPlanes<MS>,
             typename std::conditional<INTERLEAVED, MagInterleaved<TB, MS>, MagLinear<MS>>::type>::type>::type;
#endif
  // TMA staging: once pass 3 has consumed the FFT buffer, one thread starts a bulk copy (cp.async.bulk,
  // completion on an mbarrier) of the raw samples the FIRST step of the next tail batch needs into that
  // buffer; it lands while the filterbank / band stage runs, and pass 1 then reads shared memory instead
  // of waiting for L2.  float32 mono only (the copy is byte-wise); the span of all FPS frames of a step
  // must fit the PPS * BUF complex slots of the buffer.  Measured on B200 (config 2): frame 4096 unchanged
  // (7.14 ms with and without: the other groups already cover the L2 latency), frames 1024 / 2048 5 % / 4 %
  // slower (the extra barrier between reading the raw samples and overwriting them), so it is compiled in
  // only with -DB2_PAIR_TMA.
#ifdef B2_PAIR_TMA
  static constexpr bool TMA = true;
#else
  static constexpr bool TMA = false;
#endif
  static_assert(TB % FPS == 0, "tail batch must hold whole FFT steps");
  static_assert(!INPLACE || (TB == FPS && PPS == 1), "in-place magnitudes: one pair per tail batch");
};

template <int F>
inline size_t pair_smem_layout(FrontParams &p, int G) {
  using P = PairCfg<F>;
  using C2 = typename P::C2;
  auto al = [](size_t v) { return (v + 15) & ~size_t(15); };
  size_t o = 0;
  p.o_win = (int)o; o = al(o + sizeof(float) * F);
  p.o_tw3 = (int)o; o = al(o + sizeof(float2) * C2::TW3);
  p.o_pt = (int)o;                                 // no split twiddles
  p.o_wr = (int)o;  o = al(o + sizeof(float2) * C2::WR);
  p.part_stride = p.fb_ns * kGroupThreads * 4;
  p.o_w4 = (int)o;   o = al(o + (p.fb_w4_global ? 16 : sizeof(float4) * p.fb_ns * p.fb_L * kGroupThreads));
  p.o_band = (int)o; o = al(o + sizeof(int4) * (p.num_bands > 0 ? p.num_bands : 1));
  p.o_dw = (int)o;   o = al(o + sizeof(float) * (p.fb_ndw > 0 ? p.fb_ndw : 1));
  p.o_proj = (int)o; o = al(o + proj_table_bytes(p));
  p.o_groups = (int)o;
  size_t g = 0;
  g = al(g + sizeof(float2) * P::PPS * C2::BUF);   // in-place FFT buffers (one per pair)
  p.mag_stride = P::MS;
  p.g_mags = 0;         // in place: the magnitudes live in the FFT buffer
  if (!P::INPLACE) { p.g_mags = (int)g; g = al(g + sizeof(float) * P::TB * P::MS); }
  p.g_partial = (int)g; g = al(g + sizeof(float) * P::TBF * p.part_stride);
  p.g_hist = (int)g;    g = al(g + sizeof(float) * (p.diff_frames > 0 ? p.diff_frames : 1) * p.num_bands);
  p.g_lrow = (int)g;    g = al(g + sizeof(float) * P::TBF * p.num_bands);
  p.g_red = (int)g;     g = al(g + sizeof(float) * 4 * P::TBF);
  p.g_task = (int)g;    g = al(g + 16);
  p.group_bytes = (int)g;
  return o + g * G;
}

#if defined(__CUDACC__)

template <int F, int IN, int G>
__global__ void __launch_bounds__(kGroupThreads *G, 1) k_front_pair(const FrontParams p) {
  using P = PairCfg<F>;
  using C2 = typename P::C2;
  constexpr int F2 = 2 * F, PPS = P::PPS, FPS = P::FPS, R3 = C2::R3, S1 = C2::S1, TB = P::TB, TBF = P::TBF, MS = P::MS;
  extern __shared__ __align__(16) unsigned char smem[];
  float *s_win = reinterpret_cast<float *>(smem + p.o_win);
  float2 *s_tw3 = reinterpret_cast<float2 *>(smem + p.o_tw3);
  float2 *s_wr = reinterpret_cast<float2 *>(smem + p.o_wr);
  float4 *s_w4 = reinterpret_cast<float4 *>(smem + p.o_w4);
  int4 *s_band = reinterpret_cast<int4 *>(smem + p.o_band);
  float *s_dw = reinterpret_cast<float *>(smem + p.o_dw);

  for (int i = threadIdx.x; i < F; i += blockDim.x) s_win[i] = p.window[i];
  for (int i = threadIdx.x; i < C2::TW3; i += blockDim.x) s_tw3[i] = p.tw3[i];   // F-point tables (pair_tw3 / pair_wr)
  for (int i = threadIdx.x; i < C2::WR; i += blockDim.x) s_wr[i] = p.wr[i];
  if (!p.fb_w4_global)
    for (int i = threadIdx.x; i < p.fb_ns * p.fb_L * kGroupThreads; i += blockDim.x) s_w4[i] = p.fb_w4[i];
  for (int i = threadIdx.x; i < p.num_bands; i += blockDim.x) s_band[i] = p.fb_band[i];
  for (int i = threadIdx.x; i < p.fb_ndw; i += blockDim.x) s_dw[i] = p.fb_dw[i];
  TailCtx::stage_proj(p, smem);
  for (int gi = 0; gi < G; ++gi)      // magnitudes (and their padding, which zero-weight taps may read) start out finite
    for (int i = threadIdx.x; i < (P::INPLACE ? 2 * C2::BUF : TB * MS); i += blockDim.x)
      reinterpret_cast<float *>(smem + p.o_groups + (size_t)gi * p.group_bytes + p.g_mags)[i] = 0.f;
  __syncthreads();

  // virtual warp roles rotated by the group number (see k_front)
  const int g = threadIdx.x / kGroupThreads;
  const int tid = ((((threadIdx.x >> 5) + g) & 3) << 5) | (threadIdx.x & 31);
  unsigned char *gmem = smem + p.o_groups + (size_t)g * p.group_bytes;
  float2 *buf = reinterpret_cast<float2 *>(gmem);
  float *s_mags = reinterpret_cast<float *>(gmem + p.g_mags);
  volatile int *s_task = reinterpret_cast<volatile int *>(gmem + p.g_task);
  TailCtx tctx{s_w4, s_band, s_dw, s_mags, reinterpret_cast<float *>(gmem + p.g_partial),
               reinterpret_cast<float *>(gmem + p.g_hist), reinterpret_cast<float *>(gmem + p.g_lrow),
               reinterpret_cast<float *>(gmem + p.g_red), g, tid};
  tctx.resolve(p);

  // per-thread constants ---------------------------------------------------------------------
  float2 tw2r[16];                       // pass-2 twiddles of this thread's k1
#pragma unroll
  for (int n2 = 0; n2 < 16; ++n2) tw2r[n2] = __ldg(&p.tw2[(tid & 15) * 16 + n2]);
  const int sl12 = (PPS > 1) ? tid / C2::BPF : 0;           // pair slot this thread serves in pass 1/2
  const int b12 = (PPS > 1) ? tid % C2::BPF : tid;          // butterfly index in pass 1/2 (first iteration)
  float2 *p1 = buf + sl12 * C2::BUF + b12;                              // pass-1 store base
  const float *w1 = s_win + b12;                                        // window value of this butterfly's n1 = 0
  float2 wab0 = make_float2(0.f, 0.f), wab1 = wab0;                     // Hann window in registers (FrontParams::win_fly)
  if (p.win_fly) {
    wab0 = __ldg(p.win_ab + b12);
    if (C2::IT12 > 1) wab1 = __ldg(p.win_ab + b12 + kGroupThreads);
  }
  float2 *p2 = buf + sl12 * C2::BUF + (b12 & 15) * S1 + (b12 >> 4);     // pass-2 in-place base (k1, n3)
  const int u = tid;                                                    // pass-3 unit (0..127, all active)
  const int pa_off = fft_col_offset<F2>(u), pb_off = fft_col_offset<F2>((256 - u) & 255);

  const int total_tasks = p.task_off[p.n_clips];
  const int kd = p.diff_frames;
  constexpr bool kTma = P::TMA && (IN == IN_F32_MONO);
  const long long sig_bytes = kTma ? p.clip_off[p.n_clips] * 4 : 0;      // the packed buffer holds at least all clips
  uint64_t *s_bar = reinterpret_cast<uint64_t *>(gmem + p.g_task + 8);   // one mbarrier per group
  uint32_t bar_parity = 0;
  bool staged = false;                   // the samples of the next step's frames are on their way into buf
  int st_delta = 0;                      // offset (floats) of the first frame's first sample inside buf
  long long st_s0 = 0;                   // sample index (in the clip) of that first sample
  if (kTma) {
    if (tid == 0) mbar_init(s_bar, 1);
    group_bar(g);
  }

  for (;;) {
    if (tid == 0) *s_task = atomicAdd(p.task_counter, 1);
    group_bar(g);
    const int task = *s_task;
    group_bar(g);
    if (task >= total_tasks) break;

    int lo = 0, hi = p.n_clips;          // clip that owns this task: task_off[c] <= task < task_off[c+1]
    while (hi - lo > 1) {
      int mid = (lo + hi) >> 1;
      if (p.task_off[mid] <= task) lo = mid; else hi = mid;
    }
    const int c = lo;
    const long long samp0 = p.clip_off[c];
    const long long nsamp = p.clip_off[c + 1] - samp0;
    const long long row0 = p.frame_off[c];
    const int T = (int)(p.frame_off[c + 1] - row0);
    const int f0 = (task - p.task_off[c]) * p.chunk;
    const int f1 = min(T, f0 + p.chunk);
    const int fs = kd > 0 ? max(0, f0 - kd) : f0;            // warm-up rows for the difference
    Samples<IN> S{clip_base<IN>(p.sig, samp0)};
    float cscale = p.clip_scale != nullptr ? __ldg(p.clip_scale + c) : 1.f;
    if (p.power) cscale *= cscale;       // a power spectrogram scales with the gain squared
    int nonfinite = 0;

    int hslot = kd > 0 ? fs % kd : 0;                        // difference ring slot of frame fb
    staged = false;                                          // (a task never ends with a copy in flight: fn + FPS <= f1)
    for (int fb = fs; fb < f1; fb += TB) {
      // =============== FFT of the TB frames of this tail batch, FPS frames (PPS pairs) per step ===============
#pragma unroll 1
      for (int sub = 0; sub < TB; sub += FPS) {
        const int f = fb + sub;
        if (f >= f1) break;
        // ---------------- pass 1: z[n] = w[n] (xA[n] + i xB[n]), DFT16 ----------------
        const int fA = f + 2 * sl12;                         // frames of this thread's pair (fB = fA + 1)
        if (fA < f1) {
          const long long sA = (long long)((double)fA * p.hop) - (F / 2) - p.origin;
          const long long sB = (long long)((double)(fA + 1) * p.hop) - (F / 2) - p.origin;
          const bool hasB = fA + 1 < f1;
          const bool interior = (sA >= 0) && (sB + F <= nsamp) && hasB;     // sB >= sA
          if (kTma && staged && sub == 0) {
            // raw samples are in buf (TMA): take this thread's elements into registers, let the whole group
            // finish reading, then transform and overwrite the buffer
            mbar_wait(s_bar, bar_parity);
            const float *raw = reinterpret_cast<const float *>(buf) + st_delta;
            const float *ra = raw + (int)(sA - st_s0) + b12, *rb = raw + (int)(sB - st_s0) + b12;
            float xa[16 * C2::IT12], xb[16 * C2::IT12];
#pragma unroll
            for (int it = 0; it < C2::IT12; ++it)
#pragma unroll
              for (int n1 = 0; n1 < 16; ++n1) {
                xa[it * 16 + n1] = ra[it * kGroupThreads + n1 * C2::BPF];
                xb[it * 16 + n1] = rb[it * kGroupThreads + n1 * C2::BPF];
              }
            group_bar(g);
#pragma unroll
            for (int it = 0; it < C2::IT12; ++it) {
              const float *wp = w1 + it * kGroupThreads;
              fft_pass1<F2>([&](int n1) {
                const float w = wp[n1 * C2::BPF];
                return crscale(make_float2(xa[it * 16 + n1], xb[it * 16 + n1]), w);
              }, p1 + it * kGroupThreads);
            }
          } else
#if defined(B2_P1_X2)
          if (C2::IT12 == 2 && interior && p.win_fly) {     // both butterflies of this thread, loads issued together
            const void *qa = S.ptr(sA + b12), *qb = S.ptr(sB + b12);
            fft_pass1_x2<F2>(
                [&](int n1) {
                  const float w = fmaf(wab0.x, p.win_cs[n1].x, fmaf(wab0.y, p.win_cs[n1].y, p.win_h));
                  return crscale(make_float2(Samples<IN>::at_ptr(qa, n1 * C2::BPF), Samples<IN>::at_ptr(qb, n1 * C2::BPF)), w);
                },
                [&](int n1) {
                  const float w = fmaf(wab1.x, p.win_cs[n1].x, fmaf(wab1.y, p.win_cs[n1].y, p.win_h));
                  return crscale(make_float2(Samples<IN>::at_ptr(qa, n1 * C2::BPF + kGroupThreads),
                                             Samples<IN>::at_ptr(qb, n1 * C2::BPF + kGroupThreads)), w);
                },
                p1, kGroupThreads);
          } else
#endif
#pragma unroll 1
          for (int it = 0; it < C2::IT12; ++it) {
            const int b = b12 + it * kGroupThreads;
            const float *wp = w1 + it * kGroupThreads;
            if (interior && p.win_fly) {
              const void *qa = S.ptr(sA + b), *qb = S.ptr(sB + b);
              const float2 ab = it ? wab1 : wab0;
              fft_pass1<F2>([&](int n1) {
                const float w = fmaf(ab.x, p.win_cs[n1].x, fmaf(ab.y, p.win_cs[n1].y, p.win_h));
                return crscale(make_float2(Samples<IN>::at_ptr(qa, n1 * C2::BPF), Samples<IN>::at_ptr(qb, n1 * C2::BPF)), w);
              }, p1 + it * kGroupThreads);
            } else if (interior) {
              const void *qa = S.ptr(sA + b), *qb = S.ptr(sB + b);
              fft_pass1<F2>([&](int n1) {
                const float w = wp[n1 * C2::BPF];
                return crscale(make_float2(Samples<IN>::at_ptr(qa, n1 * C2::BPF), Samples<IN>::at_ptr(qb, n1 * C2::BPF)), w);
              }, p1 + it * kGroupThreads);
            } else {
              fft_pass1<F2>([&](int n1) {
                const float w = wp[n1 * C2::BPF];
                const long long a = sA + n1 * C2::BPF + b, bb = sB + n1 * C2::BPF + b;
                const float xa = (a >= 0 && a < nsamp) ? S.at(a) : 0.f;
                const float xb = (hasB && bb >= 0 && bb < nsamp) ? S.at(bb) : 0.f;
                return crscale(make_float2(xa, xb), w);
              }, p1 + it * kGroupThreads);
            }
          }
        }
        if (tid < 32) {
          // one warp pulls the samples that only the NEXT step's frames touch into L1
          constexpr int ESZ = (IN == IN_F32_MONO) ? 4 : (IN == IN_F32_STEREO) ? 8 : (IN == IN_I16_MONO) ? 2 : 4;
          const long long e0 = ((long long)((double)(f + FPS - 1) * p.hop) + (F / 2) - p.origin) * ESZ;
          long long e1 = ((long long)((double)(f + 2 * FPS - 1) * p.hop) + (F / 2) - p.origin) * ESZ;
          if (e1 > nsamp * ESZ) e1 = nsamp * ESZ;
          const char *bytes = reinterpret_cast<const char *>(S.base);
          for (long long a = (e0 & ~127LL) + tid * 128; a < e1; a += 32 * 128)
            if (a >= 0) asm volatile("prefetch.global.L1 [%0];" ::"l"(bytes + a));
        }
        group_bar(g);
        // ---------------- pass 2: twiddle, DFT16, in place ----------------
        if (fA < f1) {
          // frame 4096: both butterflies of the thread with their loads issued together (the kernels are latency
          // bound, not throughput bound: -1.9 % on B200; -DB2_NO_P2_X2 restores the loop)
#if !defined(B2_NO_P2_X2)
          if (C2::IT12 == 2) fft_pass2_x2<F2>(tw2r, p2, kGroupThreads >> 4);
          else
#endif
#pragma unroll 1
          for (int it = 0; it < C2::IT12; ++it) fft_pass2<F2>(tw2r, p2 + it * (kGroupThreads >> 4));
        }
        group_bar(g);
        // ---------------- pass 3: last radix on (column u + conj column 256-u): both frames' magnitudes ----------------
#pragma unroll 1
        for (int sl = 0; sl < PPS; ++sl) {
          if (f + 2 * sl >= f1) break;
          const float2 *fbuf = buf + sl * C2::BUF;
          float *magsA = s_mags + (sub + 2 * sl) * MS, *magsB = magsA + MS;
          auto put = [&](int bin, float ma, float mb) {
            if (P::INPLACE) {
              *reinterpret_cast<float2 *>(s_mags + MagInPlace<F2, MS>::at(bin)) = make_float2(ma, mb);
            } else if (P::PLANES) {        // plane of this pair, frames A and B side by side
              *reinterpret_cast<float2 *>(s_mags + (sub / 2 + sl) * 2 * MS + 2 * bin) = make_float2(ma, mb);
            } else if (P::INTERLEAVED) {   // frames sub + 2 sl and the next one sit side by side
              *reinterpret_cast<float2 *>(s_mags + bin * TB + sub + 2 * sl) = make_float2(ma, mb);
            } else {
              magsA[bin] = ma;
              magsB[bin] = mb;
            }
          };
          fft_pair_pass3_unit<F2>(u, fbuf + pa_off, fbuf + pb_off, s_tw3 + u,
                                  [&](int bin, float2 xa, float2 xb) { put(bin, cabs_fast(xa), cabs_fast(xb)); });
          if (tid < 32) {          // the self-paired column 128: one bin per lane, mirror bin by shuffle
            const int k3 = tid & (R3 - 1);
            constexpr int NPART = kCol128Parts<F2>;                     // R3 * NPART <= 32 lanes share the sum
            float2 Z = make_float2(0.f, 0.f);
            if (tid < R3 * NPART) Z = fft_pair_col128_part<F2>(k3, tid / R3, fbuf, s_wr);
#pragma unroll
            for (int d = R3; d < R3 * NPART; d <<= 1) {
              Z.x += __shfl_xor_sync(0xffffffffu, Z.x, d);
              Z.y += __shfl_xor_sync(0xffffffffu, Z.y, d);
            }
            const float zx = __shfl_sync(0xffffffffu, Z.x, R3 - 1 - k3), zy = __shfl_sync(0xffffffffu, Z.y, R3 - 1 - k3);
            if (tid < R3 / 2)       // |Z + conj Z'| (frame A), |Z - conj Z'| (frame B)
              put(128 + 256 * tid, cabs_fast(make_float2(Z.x + zx, Z.y - zy)), cabs_fast(make_float2(Z.x - zx, Z.y + zy)));
          }
        }
        group_bar(g);   // pass-3 reads done before the next pass 1 overwrites buf; magnitudes visible
      }
      // =============== stage the next batch's first step (TMA), then the tail of this batch ===============
      if (kTma) {
        if (staged) bar_parity ^= 1;                         // the wait of this batch's first step is done
        staged = false;
        const int fn = fb + TB;                              // first frame of the next batch
        if (fn + FPS <= f1) {
          const long long s_first = (long long)((double)fn * p.hop) - (F / 2) - p.origin;
          const long long s_last = (long long)((double)(fn + FPS - 1) * p.hop) - (F / 2) - p.origin;
          const long long g0 = samp0 + s_first;              // index in the packed buffer
          const int delta = (int)(g0 & 3);                   // the bulk copy starts on a 16-byte boundary
          const long long bytes = ((s_last - s_first + F + delta + 3) & ~3LL) * 4;
          if (s_first >= 0 && s_last + F <= nsamp && bytes <= (long long)sizeof(float2) * PPS * C2::BUF &&
              (g0 - delta) * 4 + bytes <= sig_bytes && ((reinterpret_cast<uintptr_t>(p.sig) & 15) == 0)) {
            staged = true;
            st_delta = delta;
            st_s0 = s_first;
            if (tid == 0) {
              asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic accesses to buf are done (barrier)
              mbar_expect_tx(s_bar, (uint32_t)bytes);
              tma_load_1d(buf, reinterpret_cast<const float *>(p.sig) + (g0 - delta), (uint32_t)bytes, s_bar);
            }
          }
        }
      }
      hslot = front_tail<TB, TBF, typename P::MA>(p, tctx, fb, f0, f1, row0, hslot, cscale, nonfinite);
      if (P::INPLACE) group_bar(g);   // the band stage has read its magnitudes before pass 1 overwrites the buffer
    }
    if (p.clip_status != nullptr && nonfinite) atomicOr(p.clip_status + c, 1);
  }
}

#endif  // __CUDACC__

}  // namespace b2
