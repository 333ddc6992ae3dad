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
//
// Two kernels are built from the same per-task routine (pair_run_task):
//   k_front_pair<F, IN, G>   one resolution per launch (b200spec_logfilt)
//   k_front_multi<IN, G>     up to three resolutions of one hop in ONE launch (b200spec_logfilt_multi): a group
//                            takes a (clip, chunk of frames) task through every resolution in turn, so the
//                            chunk's samples leave HBM once (the later resolutions find them in L2 / L1) and
//                            all columns of its rows are written before the group moves on.
#pragma once
#include <type_traits>

#include "frontend_kernel.cuh"

namespace b2 {

// pairs of frames a group carries through one FFT step.  The default fills one step with 128 butterflies per
// pass (one per thread) at frames 1024 / 2048 and 256 (two per thread, interleaved) at frame 4096.
template <int F>
struct PairPps {
#ifndef B2_PAIR_PPS_1024
#define B2_PAIR_PPS_1024 2
#endif
#ifndef B2_PAIR_PPS_2048
#define B2_PAIR_PPS_2048 1
#endif
  static constexpr int value = (F == 1024) ? B2_PAIR_PPS_1024 : (F == 2048) ? B2_PAIR_PPS_2048 : 1;
};

template <int F, int PPS_ = PairPps<F>::value, int TBF_MAX = 4>
struct PairCfg {
  static_assert(F == 1024 || F == 2048 || F == 4096, "pair transform: frame sizes 1024, 2048, 4096");
  using C2 = FftCfg<2 * F>;                       // geometry of the F-point complex FFT
  static constexpr int PPS = PPS_;                // pairs per group step
  static constexpr int FPS = 2 * PPS;             // frames per group step
  static constexpr int NB = PPS * C2::BPF;        // radix-16 butterflies per pass and step
  static_assert(NB % kGroupThreads == 0, "a step must give every thread the same number of butterflies");
  static constexpr int IT = NB / kGroupThreads;   // butterflies per thread and pass (1 or 2)
  static_assert(IT == 1 || IT == 2, "one or two butterflies per thread");
  // thread tid runs butterflies j = tid + it * 128 of the step: pair slot j / BPF, butterfly j % BPF
  static constexpr int TB = FftCfg<F>::TB >= FPS ? FftCfg<F>::TB : FPS;   // frames per tail batch (multiple of FPS)
  static constexpr int TBF = TB < TBF_MAX ? TB : TBF_MAX;                 // frames per filterbank / band-stage call
  static constexpr int MS = FftCfg<F>::MS;        // floats per frame in the magnitude buffer
  // frames of a filterbank call interleaved per bin (one 64-bit store per pair in pass 3, one vector load per
  // bin in the filterbank); sub-batch h of a longer tail batch starts h * MS floats in
  static constexpr bool INTERLEAVED = (TBF == 2 || TBF == 4);
  using MA = typename std::conditional<INTERLEAVED, MagInterleaved<TBF, MS>, MagLinear<MS>>::type;
  // offset between the two pass-2 butterflies of a thread (IT == 2): the next 128 butterflies are either in the
  // same pair (BPF = 256: columns n3 + 8) or in a later pair slot
  static constexpr int D2 = (C2::BPF > kGroupThreads) ? (kGroupThreads >> 4) : (kGroupThreads / C2::BPF) * C2::BUF;
  static_assert(TB % FPS == 0 && TB % TBF == 0, "tail batch must hold whole FFT steps and whole filterbank calls");
};

// byte offsets of one resolution's tables, starting at `o`; returns the end.  with_window = false leaves the
// window in global memory (o_win = -1): the pair kernel only reads the table at clip edges and for windows
// that are not np.hanning (FrontParams::win_fly).
template <int F>
inline size_t pair_table_layout(FrontParams &p, size_t o, bool with_window) {
  using C2 = FftCfg<2 * F>;
  auto al = [](size_t v) { return (v + 15) & ~size_t(15); };
  p.o_win = -1;
  if (with_window) { p.o_win = (int)o; o = al(o + sizeof(float) * F); }
  p.o_tw3 = (int)o; o = al(o + sizeof(float2) * C2::TW3C);
  p.o_pt = (int)o;                                 // no split twiddles
  p.o_wr = (int)o;  o = al(o + sizeof(float2) * C2::WR);
  p.part_stride = p.fb_ns * kGroupThreads * 4;
  p.o_w4 = (int)o;   o = al(o + (p.fb_w4_global ? 16 : sizeof(float4) * p.fb_ns * p.fb_L * kGroupThreads));
  p.o_band = (int)o; o = al(o + sizeof(int4) * (p.num_bands > 0 ? p.num_bands : 1));
  p.o_dw = (int)o;   o = al(o + sizeof(float) * (p.fb_ndw > 0 ? p.fb_ndw : 1));
  p.o_proj = (int)o; o = al(o + proj_table_bytes(p));
  return o;
}

// byte offsets inside a group's block; returns its size
template <class P>
inline size_t pair_group_layout(FrontParams &p) {
  using C2 = typename P::C2;
  auto al = [](size_t v) { return (v + 15) & ~size_t(15); };
  size_t g = 0;
  g = al(g + sizeof(float2) * P::PPS * C2::BUF);   // in-place FFT buffers (one per pair)
  p.mag_stride = P::MS;
  p.g_mags = (int)g;    g = al(g + sizeof(float) * P::TB * P::MS);
  p.g_partial = (int)g; g = al(g + sizeof(float) * P::TBF * p.part_stride);
  p.g_hist = (int)g;    g = al(g + sizeof(float) * (p.diff_frames > 0 ? p.diff_frames : 1) * p.num_bands);
  p.g_lrow = (int)g;    g = al(g + (p.num_classes > 0 ? sizeof(float) * P::TBF * p.num_bands : 0));   // rows for the projection only
  p.g_red = (int)g;     g = al(g + sizeof(float) * 4 * P::TBF);
  p.g_task = (int)g;    g = al(g + 16);
  return g;
}

template <int F>
inline size_t pair_smem_layout(FrontParams &p, int G) {
  // with the Hann window formed in registers (win_fly) the table is read at clip edges only: it stays in global
  // memory and L1 gets its 4-16 KB (frame 4096: 6.67 -> 6.59 ms on B200)
  const size_t o = pair_table_layout<F>(p, 0, !p.win_fly);
  p.o_groups = (int)o;
  p.group_bytes = (int)pair_group_layout<PairCfg<F>>(p);
  return o + (size_t)p.group_bytes * G;
}

// ---- all resolutions of one hop in one launch: parameters and shared-memory plan ---------------------------------------
constexpr int kMultiMaxRes = 3;
struct MultiParams {
  int n_res;
  FrontParams r[kMultiMaxRes];   // sig / clip_off / frame_off / n_clips / task_off / task_counter / chunk are the same in all
};

// geometry of a resolution inside k_front_multi: every step fills the group's 33 KB FFT buffer at frames 4096
// (one pair) and 2048 (two pairs) with two butterflies per thread and pass; frame 1024 keeps two pairs per step
// (four would need an eight-frame magnitude buffer).  Frames 2048 / 4096 run two frames per filterbank call,
// which keeps the partial sums at 4 KB.
template <int F>
struct MultiCfg {
  using type = PairCfg<F, (F == 4096) ? 1 : 2, (F == 1024) ? 4 : 2>;
};

// tables of every resolution side by side (windows stay in global memory), then G group blocks sized for the
// largest resolution
inline size_t multi_smem_layout(MultiParams &m, int G) {
  size_t o = 0, gb = 0;
  for (int i = 0; i < m.n_res; ++i) {
    FrontParams &p = m.r[i];
    size_t g = 0;
    switch (p.frame_size) {
      case 1024: o = pair_table_layout<1024>(p, o, false); g = pair_group_layout<MultiCfg<1024>::type>(p); break;
      case 2048: o = pair_table_layout<2048>(p, o, false); g = pair_group_layout<MultiCfg<2048>::type>(p); break;
      default: o = pair_table_layout<4096>(p, o, false); g = pair_group_layout<MultiCfg<4096>::type>(p); break;
    }
    if (g > gb) gb = g;
  }
  for (int i = 0; i < m.n_res; ++i) {
    m.r[i].o_groups = (int)o;
    m.r[i].group_bytes = (int)gb;
    m.r[i].g_task = (int)gb - 16;      // the task slot is shared by all resolutions: the last 16 bytes of the block
  }
  return o + gb * G;
}

#if defined(__CUDACC__)

// one resolution's tables for bulk_stage_begin: table(dst in shared memory, src in global memory, bytes)
template <int F, class Table>
__device__ __forceinline__ void pair_tables(const FrontParams &p, unsigned char *smem, Table &&table) {
  using C2 = FftCfg<2 * F>;
  if (p.o_win >= 0) table(smem + p.o_win, p.window, (uint32_t)sizeof(float) * F);
  table(smem + p.o_tw3, p.tw3, (uint32_t)sizeof(float2) * C2::TW3C);   // F-point tables (pair_tw3 / pair_wr)
  table(smem + p.o_wr, p.wr, (uint32_t)sizeof(float2) * C2::WR);
  if (!p.fb_w4_global) table(smem + p.o_w4, p.fb_w4, (uint32_t)sizeof(float4) * p.fb_ns * p.fb_L * kGroupThreads);
  table(smem + p.o_band, p.fb_band, (uint32_t)sizeof(int4) * p.num_bands);
  table(smem + p.o_dw, p.fb_dw, (uint32_t)sizeof(float) * p.fb_ndw);
}

// One task -- frames [f0, f1) of clip c -- of one resolution, run by the 128 threads of group g.
// ZERO_PAD: the magnitude buffer may hold another resolution's data (k_front_multi): its padding, which
// zero-weight filterbank taps read, is cleared first.
template <class P, int F, int IN, bool ZERO_PAD>
__device__ __forceinline__ void pair_run_task(const FrontParams &p, unsigned char *smem, unsigned char *gmem, int g, int tid,
                                              const float2 (&tw2r)[16], int c, int f0, int f1) {
  using C2 = typename P::C2;
  constexpr int F2 = 2 * F, PPS = P::PPS, FPS = P::FPS, R3 = C2::R3, S1 = C2::S1, TB = P::TB, TBF = P::TBF, MS = P::MS;
  constexpr int IT = P::IT, BPF = C2::BPF;
  const float *win = p.o_win >= 0 ? reinterpret_cast<const float *>(smem + p.o_win) : p.window;
  const float2 *s_tw3 = reinterpret_cast<const float2 *>(smem + p.o_tw3);
  const float2 *s_wr = reinterpret_cast<const float2 *>(smem + p.o_wr);
  float2 *buf = reinterpret_cast<float2 *>(gmem);
  float *s_mags = reinterpret_cast<float *>(gmem + p.g_mags);
  TailCtx tctx{reinterpret_cast<const float4 *>(smem + p.o_w4), reinterpret_cast<const int4 *>(smem + p.o_band),
               reinterpret_cast<const float *>(smem + p.o_dw), s_mags, reinterpret_cast<float *>(gmem + p.g_partial),
               reinterpret_cast<float *>(gmem + p.g_hist), reinterpret_cast<float *>(gmem + p.g_lrow),
               reinterpret_cast<float *>(gmem + p.g_red), g, tid};
  tctx.resolve(p);

  // butterflies of this thread in passes 1 and 2: j = tid + it * 128 -> (pair slot, butterfly)
  const int sl0 = tid / BPF, b0 = tid % BPF;                            // it = 0
  constexpr int SL_STEP = (BPF > kGroupThreads) ? 0 : kGroupThreads / BPF;     // slot / butterfly advance of it = 1
  constexpr int B_STEP = (BPF > kGroupThreads) ? kGroupThreads : 0;
  float2 *p2 = buf + sl0 * C2::BUF + (b0 & 15) * S1 + (b0 >> 4);        // pass-2 in-place base (k1, n3) of it = 0
  float2 wab0 = make_float2(0.f, 0.f), wab1 = wab0;                     // Hann window in registers (FrontParams::win_fly)
  if (p.win_fly) {
    wab0 = __ldg(p.win_ab + b0);
    wab1 = (B_STEP > 0) ? __ldg(p.win_ab + b0 + B_STEP) : wab0;
  }
  const int u = tid;                                                    // pass-3 unit (0..127, all active)
  const int pa_off = fft_col_offset<F2>(u), pb_off = fft_col_offset<F2>((256 - u) & 255);

  const int kd = p.diff_frames;
  const long long samp0 = p.clip_off[c];
  const long long nsamp = p.clip_off[c + 1] - samp0;
  const long long row0 = p.frame_off[c];
  const int fs = (kd > 0 && !p.seam_fix) ? max(0, f0 - kd) : f0;     // warm-up rows for the difference
  Samples<IN> S{clip_base<IN>(p.sig, samp0)};
  float cscale = p.clip_scale != nullptr ? __ldg(p.clip_scale + c) : 1.f;
  if (p.power) cscale *= cscale;         // a power spectrogram scales with the gain squared
  int nonfinite = 0;

  if (ZERO_PAD) {                        // bins N + 1 .. N + 15 of every frame of the batch (bin N is written by pass 3)
    constexpr int N = F / 2, NPAD = MS - N - 1;
    for (int i = tid; i < NPAD * TB; i += kGroupThreads) {
      const int fi = i / NPAD, k = N + 1 + i % NPAD;
      if (P::INTERLEAVED) s_mags[(fi / TBF) * TBF * MS + k * TBF + fi % TBF] = 0.f;
      else s_mags[fi * MS + k] = 0.f;
    }
  }

  int hslot = kd > 0 ? fs % kd : 0;                          // difference ring slot of frame fb
  for (int fb = fs; fb < f1; fb += TB) {
    // =============== FFT of the TB frames of this tail batch, FPS frames (PPS pairs) per step ===============
#pragma unroll 1
    for (int sub = 0; sub < TB; sub += FPS) {
      const int f = fb + sub;
      if (f >= f1) break;
      // ---------------- pass 1: z[n] = w[n] (xA[n] + i xB[n]), DFT16 ----------------
#pragma unroll 1
      for (int it = 0; it < IT; ++it) {
        const int sl = sl0 + it * SL_STEP, b = b0 + it * B_STEP;
        const int fA = f + 2 * sl;                           // frames of this butterfly's pair (fB = fA + 1)
        if (fA >= f1) break;
        const long long sA = (long long)((double)fA * p.hop) - (F / 2) - p.origin;
        const long long sB = (long long)((double)(fA + 1) * p.hop) - (F / 2) - p.origin;
        const bool hasB = fA + 1 < f1;
        const bool interior = (sA >= 0) && (sB + F <= nsamp) && hasB;     // sB >= sA
        float2 *p1 = buf + sl * C2::BUF + b;                 // pass-1 store base
        const float *wp = win + b;                           // window value of this butterfly's n1 = 0
        if (interior && p.win_fly) {
          const void *qa = S.ptr(sA + b), *qb = S.ptr(sB + b);
          const float2 ab = it ? wab1 : wab0;
          fft_pass1<F2>([&](int n1) {
            const float w = fmaf(ab.x, p.win_cs[n1].x, fmaf(ab.y, p.win_cs[n1].y, p.win_h));
            return crscale(make_float2(Samples<IN>::at_ptr(qa, n1 * BPF), Samples<IN>::at_ptr(qb, n1 * BPF)), w);
          }, p1);
        } else if (interior) {
          const void *qa = S.ptr(sA + b), *qb = S.ptr(sB + b);
          fft_pass1<F2>([&](int n1) {
            const float w = wp[n1 * BPF];
            return crscale(make_float2(Samples<IN>::at_ptr(qa, n1 * BPF), Samples<IN>::at_ptr(qb, n1 * BPF)), w);
          }, p1);
        } else {
          fft_pass1<F2>([&](int n1) {
            const float w = wp[n1 * BPF];
            const long long a = sA + n1 * BPF + b, bb = sB + n1 * BPF + b;
            const float xa = (a >= 0 && a < nsamp) ? S.at(a) : 0.f;
            const float xb = (hasB && bb >= 0 && bb < nsamp) ? S.at(bb) : 0.f;
            return crscale(make_float2(xa, xb), w);
          }, p1);
        }
      }
      if (tid < 32) {
        // one warp pulls the samples that only the NEXT step's frames touch into L1, so pass 1 does not wait on
        // L2 / HBM.  (Requesting the step after that into L2 as well, for the chord / key rates whose frames barely
        // overlap, was measured on B200 and LOSES: config 3b 3.36 ms against 3.17 ms.)
        constexpr int ESZ = (IN == IN_F32_MONO) ? 4 : (IN == IN_F32_STEREO) ? 8 : (IN == IN_I16_MONO) ? 2 : 4;
        const long long e0 = ((long long)((double)(f + FPS - 1) * p.hop) + (F / 2) - p.origin) * ESZ;
        long long e1 = ((long long)((double)(f + 2 * FPS - 1) * p.hop) + (F / 2) - p.origin) * ESZ;
        if (e1 > nsamp * ESZ) e1 = nsamp * ESZ;
        const char *bytes = reinterpret_cast<const char *>(S.base);
        for (long long a = (e0 & ~127LL) + tid * 128; a < e1; a += 32 * 128)
          if (a >= 0) B2_PREFETCH(bytes + a);
      }
      group_bar(g);
      // ---------------- pass 2: twiddle, DFT16, in place ----------------
      // two butterflies per thread: both with their loads issued together (the kernels are latency bound, not
      // throughput bound: -1.9 % at frame 4096 on B200)
      if (f + 2 * sl0 < f1) {
        if (IT == 2 && f + 2 * (sl0 + SL_STEP) < f1) fft_pass2_x2<F2>(tw2r, p2, P::D2);
        else fft_pass2<F2>(tw2r, p2);
      }
      group_bar(g);
      // ---------------- pass 3: last radix on (column u + conj column 256-u): both frames' magnitudes ----------------
#pragma unroll 1
      for (int sl = 0; sl < PPS; ++sl) {
        if (f + 2 * sl >= f1) break;
        const float2 *fbuf = buf + sl * C2::BUF;
        const int fi = sub + 2 * sl;                         // index of frame A inside the tail batch
        float *magsA = P::INTERLEAVED ? s_mags + (fi / TBF) * TBF * MS + fi % TBF : s_mags + fi * MS;
        auto put = [&](int bin, float ma, float mb) {
          if (P::INTERLEAVED) {            // frames fi and fi + 1 sit side by side
            *reinterpret_cast<float2 *>(magsA + bin * TBF) = make_float2(ma, mb);
          } else {
            magsA[bin] = ma;
            magsA[MS + bin] = mb;
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
    // =============== the tail of this batch ===============
    hslot = front_tail<TB, TBF, typename P::MA>(p, tctx, fb, f0, f1, row0, hslot, cscale, nonfinite);
  }
  if (p.clip_status != nullptr && nonfinite) atomicOr(p.clip_status + c, 1);
}

// the clip that owns task `task`: task_off[c] <= task < task_off[c+1]
__device__ __forceinline__ int task_clip(const int *task_off, int n_clips, int task) {
  int lo = 0, hi = n_clips;
  while (hi - lo > 1) {
    int mid = (lo + hi) >> 1;
    if (task_off[mid] <= task) lo = mid; else hi = mid;
  }
  return lo;
}

template <int F, int IN, int G>
__global__ void __launch_bounds__(kGroupThreads *G, 1) k_front_pair(const FrontParams p) {
  using P = PairCfg<F>;
  extern __shared__ __align__(16) unsigned char smem[];
  __shared__ __align__(8) uint64_t s_stage_bar;     // plan tables arrive by TMA bulk copies (bulk_stage.cuh)
  bulk_stage_begin(&s_stage_bar, [&](auto &&table) { pair_tables<F>(p, smem, table); });
  TailCtx::stage_proj(p, smem);
  for (int gi = 0; gi < G; ++gi)      // magnitudes (and their padding, which zero-weight taps may read) start out finite
    for (int i = threadIdx.x; i < P::TB * P::MS; i += blockDim.x)
      reinterpret_cast<float *>(smem + p.o_groups + (size_t)gi * p.group_bytes + p.g_mags)[i] = 0.f;
  bulk_stage_wait(&s_stage_bar);

  // virtual warp roles rotated by the group number (see k_front)
  const int g = threadIdx.x / kGroupThreads;
  const int tid = ((((threadIdx.x >> 5) + g) & 3) << 5) | (threadIdx.x & 31);
  unsigned char *gmem = smem + p.o_groups + (size_t)g * p.group_bytes;
  volatile int *s_task = reinterpret_cast<volatile int *>(gmem + p.g_task);
  float2 tw2r[16];                       // pass-2 twiddles of this thread's k1
#pragma unroll
  for (int n2 = 0; n2 < 16; ++n2) tw2r[n2] = __ldg(&p.tw2[(tid & 15) * 16 + n2]);
  const int total_tasks = p.task_off[p.n_clips];

  for (;;) {
    if (tid == 0) *s_task = atomicAdd(p.task_counter, 1);
    group_bar(g);
    const int task = *s_task;
    group_bar(g);
    if (task >= total_tasks) break;
    const int c = task_clip(p.task_off, p.n_clips, task);
    const int T = (int)(p.frame_off[c + 1] - p.frame_off[c]);
    const int ch = task_chunk(p, c);
    const int f0 = (task - p.task_off[c]) * ch;
    pair_run_task<P, F, IN, false>(p, smem, gmem, g, tid, tw2r, c, f0, min(T, f0 + ch));
  }
}

template <int IN, int G>
__global__ void __launch_bounds__(kGroupThreads *G, 1) k_front_multi(const MultiParams m) {
  extern __shared__ __align__(16) unsigned char smem[];
  __shared__ __align__(8) uint64_t s_stage_bar;     // every resolution's tables arrive by TMA bulk copies, one barrier
  bulk_stage_begin(&s_stage_bar, [&](auto &&table) {
    for (int i = 0; i < m.n_res; ++i) {
      const FrontParams &p = m.r[i];
      if (p.frame_size == 1024) pair_tables<1024>(p, smem, table);
      else if (p.frame_size == 2048) pair_tables<2048>(p, smem, table);
      else pair_tables<4096>(p, smem, table);
    }
  });
  for (int i = 0; i < m.n_res; ++i) TailCtx::stage_proj(m.r[i], smem);
  const FrontParams &p0 = m.r[0];
  for (int i = threadIdx.x; i < (p0.group_bytes * G) / 4; i += blockDim.x)     // every group block starts out zero (finite)
    reinterpret_cast<float *>(smem + p0.o_groups)[i] = 0.f;
  bulk_stage_wait(&s_stage_bar);

  const int g = threadIdx.x / kGroupThreads;
  const int tid = ((((threadIdx.x >> 5) + g) & 3) << 5) | (threadIdx.x & 31);
  unsigned char *gmem = smem + p0.o_groups + (size_t)g * p0.group_bytes;
  volatile int *s_task = reinterpret_cast<volatile int *>(gmem + p0.g_task);
  float2 tw2r[16];                       // pass-2 twiddles W_256^(n2 k1): the same for every frame size
#pragma unroll
  for (int n2 = 0; n2 < 16; ++n2) tw2r[n2] = __ldg(&p0.tw2[(tid & 15) * 16 + n2]);
  const int total_tasks = p0.task_off[p0.n_clips];
  (void)s_task;

  // The G groups of a CTA take G consecutive tasks and walk through the resolutions IN STEP (a CTA barrier per
  // resolution): all warps of the SM then run the same code at the same time -- the three resolutions' unrolled
  // FFT bodies are 370 KB of SASS, and groups in different resolutions evict each other's instructions -- and
  // neighbouring chunks of one clip share their boundary samples in L1.  The groups do the same amount of work
  // per resolution, so the barrier costs little.
  __shared__ int s_base;
  for (;;) {
    if (threadIdx.x == 0) s_base = atomicAdd(p0.task_counter, G);
    __syncthreads();
    const int base = s_base;
    if (base >= total_tasks) break;
    const int task = base + g;
    const bool valid = task < total_tasks;
    int c = 0, f0 = 0, f1 = 0;
    if (valid) {
      c = task_clip(p0.task_off, p0.n_clips, task);
      const int T = (int)(p0.frame_off[c + 1] - p0.frame_off[c]);
      f0 = (task - p0.task_off[c]) * p0.chunk;
      f1 = min(T, f0 + p0.chunk);
    }
#pragma unroll 1
    for (int i = 0; i < m.n_res; ++i) {
      const FrontParams &p = m.r[i];
      if (valid) {
        if (p.frame_size == 1024) pair_run_task<typename MultiCfg<1024>::type, 1024, IN, true>(p, smem, gmem, g, tid, tw2r, c, f0, f1);
        else if (p.frame_size == 2048) pair_run_task<typename MultiCfg<2048>::type, 2048, IN, true>(p, smem, gmem, g, tid, tw2r, c, f0, f1);
        else pair_run_task<typename MultiCfg<4096>::type, 4096, IN, true>(p, smem, gmem, g, tid, tw2r, c, f0, f1);
      }
      __syncthreads();   // every group is done with this resolution (and its buffers) before any starts the next
    }
  }
}

#endif  // __CUDACC__

}  // namespace b2
