// frontend_kernel.cuh -- the fused front-end kernel (K1 framing+window+FFT, K2 magnitude+filterbank
// +log, K3 difference/flux/projection) as one persistent sm_100a kernel per resolution.
//
// Replaces, for one resolution, the madmom 0.16.1 chain that
// /root/reference/backend/app/services/grid/beats.py:74 (RNNBeatProcessor),
// chords/extract.py:54 (DeepChromaProcessor) and theory/key.py:101 (CNNKeyRecognitionProcessor) run:
//   signal_frame -> frame*fft_window -> fftpack.fft[:F/2] -> np.abs -> np.dot(., filterbank)
//   -> np.log10(mul*y+add) -> SpectrogramDifference(positive) -> np.hstack
//
// Execution model
//   grid   = one CTA per SM (persistent), G groups of 128 threads per CTA
//   group  = pulls tasks (clip, chunk of frames) from a global counter; processes FPG frames per
//            step entirely in shared memory (one in-place FFT buffer per frame); named barriers
//            (bar.sync id,128) keep the groups independent of each other
//   tables = window / twiddles / banded filterbank are staged once per CTA in shared memory,
//            pass-2 twiddles live in registers
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <type_traits>

#include "bulk_stage.cuh"
#include "fb_core.cuh"
#include "fft_core.cuh"

namespace b2 {

enum InputKind { IN_F32_MONO = 0, IN_F32_STEREO = 1, IN_I16_MONO = 2, IN_I16_STEREO = 3 };
enum KernelMode { MODE_LOGFILT = 0, MODE_SPECTRUM = 1 };

struct FrontParams {
  // input
  const void *sig;
  const long long *clip_off;
  const long long *frame_off;
  int n_clips;
  const int *task_off;  // n_clips + 1 (workspace, written by k_setup_tasks)
  int *task_counter;    // workspace
  int chunk;            // frames per task
  int frame_size;
  double hop;
  int origin;
  // tables (plan-owned, device)
  const float *window;  // F, already * 1/2 (and / 32767 for int16)
  // Hann window evaluated in registers by the pair kernel (win_fly != 0; the table above stays for the
  // clip edges and for any other window):  w[b + n1 BPF] = win_h + A_b C[n1] + B_b S[n1]  with
  // (A_b, B_b) = win_ab[b] = (-H cos t_b, H sin t_b), t_b = 2 pi b / (F-1), (C, S)[n1] = win_cs[n1] =
  // (cos, sin)(2 pi n1 BPF / (F-1)) -- kernel-parameter constants, so the products take them from the
  // constant bank and the per-sample window value costs no load at all
  int win_fly;
  float win_h;
  float2 win_cs[16];
  float2 win_cs32[32];  // the same for the warp kernel (frame 1024, 32 points per lane): (cos, sin)(2 pi 32 a / (F-1))
  const float2 *win_ab; // [BPF of the pair geometry = F / 16]
  const float2 *tw2;    // [16][16]   tw2[k1*16 + n2] = W_256^(n2 k1)
  const float2 *tw3;    // [R3][129]  tw3[n3*129 + q] = W_N^(n3 q)
  const float2 *pt;     // [R3][129]  pt[k3*129 + q]  = -i W_F^(q + 256 k3)
  const float2 *wr;     // [2 R3]     wr[e] = W_(2 R3)^e
  // filterbank (MODE_LOGFILT), slab form: the used bins [fb_kmin, kmax) are cut into slabs of fb_L
  // (odd) consecutive bins; thread t of a group owns slabs t, t + 128, ... (fb_ns of them) and keeps
  // four running sums per frame, one per band index modulo 4 (a slab never touches two bands with
  // the same index modulo 4; bands too narrow for that are "direct" and summed in the band stage)
  int num_bands, nnz, kmax;
  int fb_L, fb_ns, fb_kmin, fb_ndw;
  int fb_w4_global;      // 1: fb_w4 stays in global memory (too large for shared memory; fb_L is 15 then)
  const float4 *fb_w4;   // [fb_ns][fb_L][128]: weights of bin i of thread t's slab, by band index & 3
  const int4 *fb_band;   // per band: slab band {first partial, count, 0, 0}; direct band {dw offset, 0, first bin, taps}
  const float *fb_dw;    // weights of the direct bands
  int log_enabled;
  float mul, add;
  int power;               // 1: filterbank on |X|^2 (the magnitudes are squared as they are read)
  float log_scale, log_floor;
  int diff_frames, positive;
  int num_classes, nproj;   // nproj = proj_off[num_classes]
  const int *proj_off, *proj_band;
  const float *proj_w;
  // outputs (MODE_LOGFILT)
  float *out;
  long long ld_out;
  int col_spec, col_diff;
  float *flux;
  const float *clip_scale;   // per-clip gain on the samples (nullptr = 1): band sums scale with it (its square when power)
  int *clip_status;          // per-clip status word (nullptr = not wanted): bit 0 = a non-finite output value
  float *proj;
  long long ld_proj;
  // outputs (MODE_SPECTRUM)
  float *spec_out;      // (rows, spec_ld) float or float2
  int spec_ld;          // N, or N + 1 with the Nyquist bin (madmom include_nyquist)
  // 1: a task transforms its own frames only -- no warm-up rows for the lagged difference.  The first diff_frames
  // rows of every task (but a clip's first) then hold a difference against a stale ring; k_seam_diff rewrites them
  // from the (log-)filtered rows in global memory once the kernel is done (b200spec.cu: launch_front).
  int seam_fix;
  // tasks of the LAST tail_clips clips are chunk_small frames long instead of chunk: the counter hands out the long
  // tasks first, and workers that find no long task left fill the last round with short ones (0 = one size)
  int chunk_small, tail_clips;
  int spec_complex;
  int circular_shift;   // madmom stft(circular_shift=True) with fft_size == frame_size: the two halves of the windowed
                        // frame are swapped before the transform = bin k times (-1)^k (magnitudes are unchanged)
  // shared-memory carve-up (byte offsets), filled by front_smem_layout()
  int o_win, o_tw3, o_pt, o_wr, o_w4, o_band, o_dw, o_proj, o_groups, group_bytes;
  int g_mags, g_partial, g_hist, g_lrow, g_red, g_task;  // offsets inside a group's block
  int mag_stride;                                        // floats per frame in the magnitude buffer
  int mag_cap;                                           // floats of a magnitude row that are kept (k_front<8192>: kmax + slack)
  int part_stride;                                       // floats per frame in the partial-sum buffer
};

// samples that only the next step touches are requested ahead of pass 1 (tuning: -DB2_PREFETCH_MODE=1 L2 only, 2 none)
#ifndef B2_PREFETCH_MODE
#define B2_PREFETCH_MODE 0
#endif
#if B2_PREFETCH_MODE == 0
#define B2_PREFETCH(ptr) asm volatile("prefetch.global.L1 [%0];" ::"l"(ptr))
#elif B2_PREFETCH_MODE == 1
#define B2_PREFETCH(ptr) asm volatile("prefetch.global.L2 [%0];" ::"l"(ptr))
#else
#define B2_PREFETCH(ptr) ((void)(ptr))
#endif

// frames per task of clip c (FrontParams::tail_clips)
B2_HD int task_chunk(const FrontParams &p, int c) { return c >= p.n_clips - p.tail_clips ? p.chunk_small : p.chunk; }

// the projection (chroma fold / PCP) in CSR-by-class form, staged in shared memory: class offsets, bands, weights
inline size_t proj_table_bytes(const FrontParams &p) {
  return p.num_classes > 0 ? sizeof(int) * (size_t)(p.num_classes + 1) + (sizeof(int) + sizeof(float)) * (size_t)p.nproj : 0;
}

template <int F>
inline size_t front_smem_layout(FrontParams &p, int mode, int G) {
  using C = FftCfg<F>;
  auto al = [](size_t v) { return (v + 15) & ~size_t(15); };
  size_t o = 0;
  p.o_win = (int)o; o = al(o + sizeof(float) * F);
  p.o_tw3 = (int)o; o = al(o + sizeof(float2) * C::TW3);
  p.o_pt = (int)o;  o = al(o + sizeof(float2) * C::PT);
  p.o_wr = (int)o;  o = al(o + sizeof(float2) * C::WR);
  p.o_w4 = p.o_band = p.o_dw = p.o_proj = (int)o;
  p.part_stride = p.fb_ns * kGroupThreads * 4;
  if (mode == MODE_LOGFILT) {
    p.o_w4 = (int)o;   o = al(o + (p.fb_w4_global ? 16 : sizeof(float4) * p.fb_ns * p.fb_L * kGroupThreads));
    p.o_band = (int)o; o = al(o + sizeof(int4) * (p.num_bands > 0 ? p.num_bands : 1));
    p.o_dw = (int)o;   o = al(o + sizeof(float) * (p.fb_ndw > 0 ? p.fb_ndw : 1));
    p.o_proj = (int)o; o = al(o + proj_table_bytes(p));
  }
  p.o_groups = (int)o;
  size_t g = 0;
  g = al(g + sizeof(float2) * C::FPG * C::BUF);  // in-place FFT buffers
  p.g_mags = p.g_partial = p.g_hist = p.g_lrow = (int)g;
  p.mag_stride = C::MS;
  if (mode == MODE_LOGFILT) {
    if (p.mag_cap <= 0 || p.mag_cap > C::MS || C::TB > 1) p.mag_cap = C::MS;   // rows are MS apart when a batch holds several
    p.g_mags = (int)g;    g = al(g + sizeof(float) * (C::TB > 1 ? C::TB * p.mag_stride : p.mag_cap));
    p.g_partial = (int)g; g = al(g + sizeof(float) * C::TBF * p.part_stride);
    p.g_hist = (int)g;    g = al(g + sizeof(float) * (p.diff_frames > 0 ? p.diff_frames : 1) * p.num_bands);
    p.g_lrow = (int)g;    g = al(g + sizeof(float) * C::TBF * p.num_bands);
  }
  p.g_red = (int)g;  g = al(g + sizeof(float) * 4 * C::TBF);
  p.g_task = (int)g; g = al(g + 16);
  p.group_bytes = (int)g;
#ifdef B2_SMEM_PAD   // tuning experiment: shrink the L1 carve-out by this many bytes
  return o + g * G + B2_SMEM_PAD;
#else
  return o + g * G;
#endif
}

#if defined(__CUDACC__)

__device__ __forceinline__ void group_bar(int g) {
  asm volatile("bar.sync %0, %1;" ::"r"(g + 1), "r"(kGroupThreads) : "memory");
}

__device__ __forceinline__ float fast_sqrt(float v) {
  float r;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v));  // one MUFU; max rel. error 2^-23
  return r;
}

__device__ __forceinline__ float fast_lg2(float v) {
  float r;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v));   // one MUFU; the argument is >= add (or the floor) here
  return r;
}

__device__ __forceinline__ float cabs_fast(float2 X) { return fast_sqrt(fmaf(X.x, X.x, X.y * X.y)); }

// ---- sample access: madmom Signal dtype + remix (audio/signal.py) ------------------------------
template <int IN>
struct Samples {
  const void *base;  // first sample of the clip
  __device__ __forceinline__ float at(long long s) const {
    if (IN == IN_F32_MONO) {
      return __ldg(reinterpret_cast<const float *>(base) + s);
    } else if (IN == IN_F32_STEREO) {
      float2 v = __ldg(reinterpret_cast<const float2 *>(base) + s);
      return (v.x + v.y) * 0.5f;                      // np.mean(axis=-1) in float32
    } else if (IN == IN_I16_MONO) {
      return (float)__ldg(reinterpret_cast<const short *>(base) + s);
    } else {
      short2 v = __ldg(reinterpret_cast<const short2 *>(base) + s);
      return (float)(((int)v.x + (int)v.y) / 2);      // float64 mean cast back to int16: truncation
    }
  }
  // Address of sample s, formed ONCE; at_ptr(q, off) then loads sample s + off with `off` folded into
  // the load's immediate field (off is a compile-time constant at every call site) instead of fresh
  // 64-bit address arithmetic per load.
  __device__ __forceinline__ const void *ptr(long long s) const {
    if (IN == IN_F32_MONO) return reinterpret_cast<const float *>(base) + s;
    if (IN == IN_F32_STEREO) return reinterpret_cast<const float2 *>(base) + s;
    if (IN == IN_I16_MONO) return reinterpret_cast<const short *>(base) + s;
    return reinterpret_cast<const short2 *>(base) + s;
  }
  static __device__ __forceinline__ float at_ptr(const void *q, int off) {
    if (IN == IN_F32_MONO) {
      return __ldg(reinterpret_cast<const float *>(q) + off);
    } else if (IN == IN_F32_STEREO) {
      float2 v = __ldg(reinterpret_cast<const float2 *>(q) + off);
      return (v.x + v.y) * 0.5f;
    } else if (IN == IN_I16_MONO) {
      return (float)__ldg(reinterpret_cast<const short *>(q) + off);
    } else {
      short2 v = __ldg(reinterpret_cast<const short2 *>(q) + off);
      return (float)(((int)v.x + (int)v.y) / 2);
    }
  }
};

template <int IN>
__device__ __forceinline__ const void *clip_base(const void *sig, long long off) {
  if (IN == IN_F32_MONO) return reinterpret_cast<const float *>(sig) + off;
  if (IN == IN_F32_STEREO) return reinterpret_cast<const float2 *>(sig) + off;
  if (IN == IN_I16_MONO) return reinterpret_cast<const short *>(sig) + off;
  return reinterpret_cast<const short2 *>(sig) + off;
}

// ---- K2/K3 tail shared by k_front and k_front_pair: slab filterbank, band sums, log, lagged difference,
// stacked stores, flux / projection, for the TB frames whose magnitudes sit in c.s_mags -------------------
struct TailCtx {
  const float4 *s_w4;
  const int4 *s_band;
  const float *s_dw;
  float *s_mags, *s_partial, *s_hist, *s_lrow, *s_red;
  int g, tid;
  int mag_cap;                    // floats of a magnitude row that exist (0: all MS of them)
  // band-stage constants, resolved once per kernel instead of per output element
  float *out_spec, *out_diff;     // p.out + col_spec / + col_diff, or nullptr when that half is not wanted
  bool do_log, positive;
  float lmul, ladd, lfloor, lk;   // lk = log_scale * log10(2): log10(a) = lg2(a) * log10(2)

  // copies the projection tables into shared memory (all threads of the CTA; the caller syncs afterwards)
  static __device__ __forceinline__ void stage_proj(const FrontParams &p, unsigned char *smem) {
    if (p.num_classes <= 0) return;
    int *poff = reinterpret_cast<int *>(smem + p.o_proj);
    int *pband = poff + p.num_classes + 1;
    float *pw = reinterpret_cast<float *>(pband + p.nproj);
    for (int i = threadIdx.x; i <= p.num_classes; i += blockDim.x) poff[i] = p.proj_off[i];
    for (int i = threadIdx.x; i < p.nproj; i += blockDim.x) pband[i] = p.proj_band[i], pw[i] = p.proj_w[i];
  }
  __device__ __forceinline__ void resolve(const FrontParams &p) {
    out_spec = (p.out != nullptr && p.col_spec >= 0) ? p.out + p.col_spec : nullptr;
    out_diff = (p.out != nullptr && p.col_diff >= 0) ? p.out + p.col_diff : nullptr;
    do_log = p.log_enabled != 0;
    positive = p.positive != 0;
    lmul = p.mul;
    ladd = p.add;
    lfloor = p.log_floor;
    lk = p.log_scale * 0.30102999566398120f;
  }
};

template <int TB, int TBF, class MA>
__device__ __forceinline__ int front_tail(const FrontParams &p, const TailCtx &c, int fb, int f0, int f1,
                                          long long row0, int hslot, float cscale, int &nonfinite) {
  const float4 *s_w4 = c.s_w4;
  const int4 *s_band = c.s_band;
  const float *s_dw = c.s_dw;
  float *s_mags = c.s_mags, *s_partial = c.s_partial, *s_hist = c.s_hist, *s_lrow = c.s_lrow, *s_red = c.s_red;
  const int g = c.g, tid = c.tid;
  const int B = p.num_bands, kd = p.diff_frames, pstride = p.part_stride;
  constexpr int MS = MA::MS;
  const int cap = c.mag_cap > 0 ? c.mag_cap : MS;
#pragma unroll 1
  for (int h = 0; h < TB; h += TBF) {
    const int fh = fb + h;
    if (fh >= f1) break;
    if (h > 0) group_bar(g);                     // the previous sub-batch is done with s_partial
    const float *hmags = s_mags + MA::batch_offset(h);
    // ---- K2a: slab filterbank ----
    if (p.fb_w4_global) {
      if (p.power) fb_slabs<15, TBF, MA, true, true>(p.fb_w4, hmags, s_partial, p.fb_ns, p.fb_kmin, pstride, tid, cap);
      else fb_slabs<15, TBF, MA, true, false>(p.fb_w4, hmags, s_partial, p.fb_ns, p.fb_kmin, pstride, tid, cap);
    } else if (p.power) {
      fb_slabs_dispatch<TBF, MA, true>(p.fb_L, s_w4, hmags, s_partial, p.fb_ns, p.fb_kmin, pstride, tid, cap);
    } else {
      fb_slabs_dispatch<TBF, MA, false>(p.fb_L, s_w4, hmags, s_partial, p.fb_ns, p.fb_kmin, pstride, tid, cap);
    }
    group_bar(g);
    // ---- K2b/K3: band sum, log10, lagged difference, stacked store ----
    // lane = band.  Trip counts are made warp-uniform (REDUX max) and lanes past their own count
    // add 0, so the gather loops carry no divergence.
    float fluxacc[TBF];
#pragma unroll
    for (int t = 0; t < TBF; ++t) fluxacc[t] = 0.f;
    const int nvalid = min(TBF, f1 - fh);        // frames of this sub-batch that exist
    const int nskip = max(0, f0 - fh);           // leading warm-up frames: they only feed the difference ring
    // the common case: every frame of the sub-batch exists, none is a warm-up row and none is among the first kd
    // rows of the clip -- the per-frame tests below then fold away
    const bool full = (nvalid == TBF) && (nskip == 0) && (fh >= kd);
    for (int jb = 0; jb < B; jb += kGroupThreads) {
      if (jb + (tid & ~31) >= B) continue;       // no band for any lane of this warp (warp-uniform): nothing to do
      const int j = jb + tid;
      const bool valid = j < B;
      const int4 bd = valid ? s_band[j] : make_int4(0, 0, 0, 0);
      const int nP = __reduce_max_sync(0xffffffffu, bd.y), nD = __reduce_max_sync(0xffffffffu, bd.w);
      float ysum[TBF];
#pragma unroll
      for (int t = 0; t < TBF; ++t) ysum[t] = 0.f;
      if (PartialLayout<TBF>::ADJACENT) {        // slab band: partial sums of the slabs it touches
        const float *pp = s_partial + bd.x * TBF;
#pragma unroll 4
        for (int i = 0; i < nP; ++i)
          if (i < bd.y) FrameVec<TBF>::add(pp + 4 * TBF * i, ysum);
      } else {
        const float *pp = s_partial + bd.x;
#pragma unroll 4
        for (int i = 0; i < nP; ++i) {
          const bool on = i < bd.y;
#pragma unroll
          for (int t = 0; t < TBF; ++t) ysum[t] += on ? pp[t * pstride + 4 * i] : 0.f;
        }
      }
      const float *dwp = s_dw + bd.x;
#pragma unroll 2
      for (int i = 0; i < nD; ++i) {             // direct band: its few taps straight from the magnitudes
        const bool on = i < bd.w;
        const float w = on ? dwp[i] : 0.f;
        float mm[TBF];
#pragma unroll
        for (int t = 0; t < TBF; ++t) mm[t] = 0.f;
        if (on) MA::load(hmags, bd.z + i, mm);
#pragma unroll
        for (int t = 0; t < TBF; ++t) {
          float m = mm[t];
          if (p.power) m *= m;
          ysum[t] = fmaf(w, m, ysum[t]);
        }
      }
      if (valid) {
        // the two output pointers of this band for the first frame of the sub-batch (advanced by ld_out)
        float *ps = c.out_spec + (row0 + fh) * p.ld_out + j;
        float *pd = c.out_diff + (row0 + fh) * p.ld_out + j;
        float *hp = s_hist + hslot * B + j;
        int slot = hslot;
        auto rows = [&](auto full_tag) {
          constexpr bool FULL = decltype(full_tag)::value;
#pragma unroll
          for (int t = 0; t < TBF; ++t) {
            if (FULL || t < nvalid) {
              float L = ysum[t] * cscale;
              if (c.do_log) {
                float a = __fadd_rn(__fmul_rn(c.lmul, L), c.ladd);    // separate multiply and add, as numpy
                if (c.lfloor > 0.f) a = fmaxf(a, c.lfloor);
                L = fast_lg2(a) * c.lk;                               // log_scale * log10(a)
              }
              nonfinite |= !(fabsf(L) <= 3.402823466e38f);           // NaN or Inf (from NaN / Inf samples)
              float D = 0.f;
              if (kd > 0) {
                const float old = *hp;
                *hp = L;
                if (FULL || fh + t >= kd) D = L - old;
                if (c.positive) D = fmaxf(D, 0.f);
                ++slot;
                hp += B;
                if (slot == kd) slot = 0, hp -= kd * B;
              }
              if (p.num_classes > 0) s_lrow[t * B + j] = L;
              if (FULL || t >= nskip) {
                if (c.out_spec != nullptr) *ps = L;
                if (c.out_diff != nullptr) *pd = D;
                fluxacc[t] += D;
              }
            }
            ps += p.ld_out;
            pd += p.ld_out;
          }
        };
        if (full) rows(std::true_type{});
        else rows(std::false_type{});
      }
    }
    if (kd > 0) {                                // ring position of the next step's first frame
      hslot += TBF;
      while (hslot >= kd) hslot -= kd;
    }
    if (p.flux != nullptr || p.num_classes > 0) {
      if (p.flux != nullptr) {
#pragma unroll
        for (int t = 0; t < TBF; ++t) {
          float v = fluxacc[t];
#pragma unroll
          for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
          if ((tid & 31) == 0) s_red[t * 4 + (tid >> 5)] = v;
        }
      }
      group_bar(g);
      if (p.flux != nullptr && tid < TBF) {
        const int frame = fh + tid;
        if (frame >= f0 && frame < f1)
          p.flux[row0 + frame] = (s_red[tid * 4] + s_red[tid * 4 + 1]) + (s_red[tid * 4 + 2] + s_red[tid * 4 + 3]);
      }
      // projection: one thread per (frame, class); tables (stage_proj) and rows in shared memory
      extern __shared__ __align__(16) unsigned char smem_base[];
      const int *s_poff = reinterpret_cast<const int *>(smem_base + p.o_proj);
      const int *s_pband = s_poff + p.num_classes + 1;
      const float *s_pw = reinterpret_cast<const float *>(s_pband + p.nproj);
      for (int i = tid; i < TBF * p.num_classes; i += kGroupThreads) {
        const int t = i / p.num_classes, cls = i - t * p.num_classes;
        const int frame = fh + t;
        if (frame < f0 || frame >= f1) continue;
        float acc = 0.f;
        for (int k = s_poff[cls]; k < s_poff[cls + 1]; ++k) acc = fmaf(s_pw[k], s_lrow[t * B + s_pband[k]], acc);
        p.proj[(row0 + frame) * p.ld_proj + cls] = acc;
      }
    }
  }
  return hslot;
}

// ---- the front-end kernel ----------------------------------------------------------------------
template <int F, int IN, int MODE, int G>
__global__ void __launch_bounds__(kGroupThreads *G, 1) k_front(const FrontParams p) {
  using C = FftCfg<F>;
  constexpr int FPG = C::FPG, N = C::N, R3 = C::R3, S1 = C::S1, TB = C::TB, TBF = C::TBF, MS = C::MS;
  extern __shared__ __align__(16) unsigned char smem[];
  float *s_win = reinterpret_cast<float *>(smem + p.o_win);
  float2 *s_tw3 = reinterpret_cast<float2 *>(smem + p.o_tw3);
  float2 *s_pt = reinterpret_cast<float2 *>(smem + p.o_pt);
  float2 *s_wr = reinterpret_cast<float2 *>(smem + p.o_wr);
  float4 *s_w4 = reinterpret_cast<float4 *>(smem + p.o_w4);
  int4 *s_band = reinterpret_cast<int4 *>(smem + p.o_band);
  float *s_dw = reinterpret_cast<float *>(smem + p.o_dw);

  // plan tables: bulk copies by the TMA unit (bulk_stage.cuh), in flight while the work buffers are cleared
  __shared__ __align__(8) uint64_t s_stage_bar;
  bulk_stage_begin(&s_stage_bar, [&](auto &&table) {
    table(s_win, p.window, (uint32_t)sizeof(float) * F);
    table(s_tw3, p.tw3, (uint32_t)sizeof(float2) * C::TW3);
    table(s_pt, p.pt, (uint32_t)sizeof(float2) * C::PT);
    table(s_wr, p.wr, (uint32_t)sizeof(float2) * C::WR);
    if (MODE == MODE_LOGFILT) {
      if (!p.fb_w4_global) table(s_w4, p.fb_w4, (uint32_t)sizeof(float4) * p.fb_ns * p.fb_L * kGroupThreads);
      table(s_band, p.fb_band, (uint32_t)sizeof(int4) * p.num_bands);
      table(s_dw, p.fb_dw, (uint32_t)sizeof(float) * p.fb_ndw);
    }
  });
  if (MODE == MODE_LOGFILT) {
    TailCtx::stage_proj(p, smem);
    // magnitudes (and their padding, which zero-weight taps may read) start out finite
    float *allmags = reinterpret_cast<float *>(smem + p.o_groups);
    for (int gi = 0; gi < G; ++gi)
      for (int i = threadIdx.x; i < (TB > 1 ? TB * MS : p.mag_cap); i += blockDim.x)
        reinterpret_cast<float *>(reinterpret_cast<unsigned char *>(allmags) + (size_t)gi * p.group_bytes + p.g_mags)[i] = 0.f;
  }
  bulk_stage_wait(&s_stage_bar);

  // Warp w of every group lands on scheduler (SMSP) w.  Roles inside a group are not uniform (the
  // self-paired columns run on virtual warp 0, the band stage uses the low warps), so the virtual
  // warp index is rotated by the group number: the heavy roles of the G groups then sit on
  // different schedulers instead of all on SMSP 0.
#ifdef B2_GROUP_BY_SMSP   // tuning experiment: all four warps of a group on one scheduler
  const int g = (threadIdx.x >> 5) % G;
  const int tid = (((threadIdx.x >> 5) / G) << 5) | (threadIdx.x & 31);
#else
  const int g = threadIdx.x / kGroupThreads;
  const int tid = ((((threadIdx.x >> 5) + g) & 3) << 5) | (threadIdx.x & 31);
#endif
  unsigned char *gmem = smem + p.o_groups + (size_t)g * p.group_bytes;
  float2 *buf = reinterpret_cast<float2 *>(gmem);
  float *s_mags = reinterpret_cast<float *>(gmem + p.g_mags);
  float *s_partial = reinterpret_cast<float *>(gmem + p.g_partial);
  float *s_hist = reinterpret_cast<float *>(gmem + p.g_hist);
  float *s_lrow = reinterpret_cast<float *>(gmem + p.g_lrow);
  float *s_red = reinterpret_cast<float *>(gmem + p.g_red);
  volatile int *s_task = reinterpret_cast<volatile int *>(gmem + p.g_task);
  TailCtx tctx{s_w4, s_band, s_dw, s_mags, s_partial, s_hist, s_lrow, s_red, g, tid};
  tctx.resolve(p);
  tctx.mag_cap = p.mag_cap;
  const int kcap = p.mag_cap;            // magnitude bins >= kcap feed no band: neither formed nor stored (frame 8192)

  // per-thread constants ---------------------------------------------------------------------
  float2 tw2r[16];                       // pass-2 twiddles of this thread's k1
#pragma unroll
  for (int n2 = 0; n2 < 16; ++n2) tw2r[n2] = __ldg(&p.tw2[(tid & 15) * 16 + n2]);
  const int fl12 = (FPG > 1) ? tid / C::BPF : 0;            // frame slot this thread serves in pass 1/2
  const int b12 = (FPG > 1) ? tid % C::BPF : tid;           // butterfly index in pass 1/2 (first iteration)
  float2 *p1 = buf + fl12 * C::BUF + b12;                               // pass-1 store base
  const float *w1 = s_win + 2 * b12;                                    // window pairs of this butterfly
  float2 *p2 = buf + fl12 * C::BUF + (b12 & 15) * S1 + (b12 >> 4);      // pass-2 in-place base (k1, n3)
  const int u = tid;                                                    // pass-3 unit
  const int pa_off = fft_col_offset<F>(u), pb_off = fft_col_offset<F>((256 - u) & 255);

  const int total_tasks = p.task_off[p.n_clips];
  const int kd = p.diff_frames;

  for (;;) {
    if (tid == 0) *s_task = atomicAdd(p.task_counter, 1);
    group_bar(g);
    const int task = *s_task;
    group_bar(g);
    if (task >= total_tasks) break;

    // clip that owns this task: task_off[c] <= task < task_off[c+1]
    int lo = 0, hi = p.n_clips;
    while (hi - lo > 1) {
      int mid = (lo + hi) >> 1;
      if (p.task_off[mid] <= task) lo = mid; else hi = mid;
    }
    const int c = lo;
    const long long samp0 = p.clip_off[c];
    const long long nsamp = p.clip_off[c + 1] - samp0;
    const long long row0 = p.frame_off[c];
    const int T = (int)(p.frame_off[c + 1] - row0);
    const int ch = task_chunk(p, c);
    const int f0 = (task - p.task_off[c]) * ch;
    const int f1 = min(T, f0 + ch);
    const int fs = (MODE == MODE_LOGFILT && kd > 0 && !p.seam_fix) ? max(0, f0 - kd) : f0;  // warm-up rows for the diff
    Samples<IN> S{clip_base<IN>(p.sig, samp0)};
    float cscale = (MODE == MODE_LOGFILT && p.clip_scale != nullptr) ? __ldg(p.clip_scale + c) : 1.f;
    if (p.power) cscale *= cscale;       // |g x|^2 = g^2 |x|^2: a power spectrogram scales with the gain squared
    int nonfinite = 0;

    int hslot = (MODE == MODE_LOGFILT && kd > 0) ? fs % kd : 0;   // difference ring slot of frame f
    for (int fb = fs; fb < f1; fb += TB) {
      // =============== FFT of the TB frames of this tail batch, FPG frames per step ===============
#pragma unroll 1
      for (int sub = 0; sub < TB; sub += FPG) {
        const int f = fb + sub;
        if (f >= f1) break;
        // ---------------- pass 1: frame load * window, DFT16 ----------------
        if (f + fl12 < f1) {
          const long long s0 = (long long)((double)(f + fl12) * p.hop) - (F / 2) - p.origin;
          {
            const bool interior = (s0 >= 0) && (s0 + F <= nsamp);
#pragma unroll 1
            for (int it = 0; it < C::IT12; ++it) {
              const long long sb = s0 + 2 * (b12 + it * kGroupThreads);
              const float *wp = w1 + 2 * it * kGroupThreads;
              if (interior) {
                const void *q = S.ptr(sb);
                fft_pass1<F>([&](int n1) {
                  float2 w = *reinterpret_cast<const float2 *>(wp + 2 * n1 * C::BPF);
                  return emul(w, make_float2(Samples<IN>::at_ptr(q, 2 * n1 * C::BPF), Samples<IN>::at_ptr(q, 2 * n1 * C::BPF + 1)));
                }, p1 + it * kGroupThreads);
              } else {
                fft_pass1<F>([&](int n1) {
                  float2 w = *reinterpret_cast<const float2 *>(wp + 2 * n1 * C::BPF);
                  const long long sa = sb + 2 * n1 * C::BPF, sc = sa + 1;
                  float xa = (sa >= 0 && sa < nsamp) ? S.at(sa) : 0.f;
                  float xb = (sc >= 0 && sc < nsamp) ? S.at(sc) : 0.f;
                  return emul(w, make_float2(xa, xb));
                }, p1 + it * kGroupThreads);
              }
            }
          }
        }
        if (tid < 32) {
          // one warp pulls the samples that only the NEXT step's frames touch into L1 (the rest of
          // their windows overlaps what was just read), so pass 1 does not wait on L2/HBM
          constexpr int ESZ = (IN == IN_F32_MONO) ? 4 : (IN == IN_F32_STEREO) ? 8 : (IN == IN_I16_MONO) ? 2 : 4;
          const long long e0 = ((long long)((double)(f + FPG - 1) * p.hop) + (F / 2) - p.origin) * ESZ;
          long long e1 = ((long long)((double)(f + 2 * FPG - 1) * p.hop) + (F / 2) - p.origin) * ESZ;
          if (e1 > nsamp * ESZ) e1 = nsamp * ESZ;
          const char *bytes = reinterpret_cast<const char *>(S.base);
          for (long long a = (e0 & ~127LL) + tid * 128; a < e1; a += 32 * 128)
            if (a >= 0) B2_PREFETCH(bytes + a);
        }
        group_bar(g);
        // ---------------- pass 2: twiddle, DFT16, in place ----------------
        if (f + fl12 < f1) {
#if !defined(B2_NO_P2_X2)
          if (C::IT12 == 2) fft_pass2_x2<F>(tw2r, p2, kGroupThreads >> 4);   // frame 8192: two butterflies interleaved
          else
#endif
#pragma unroll 1
          for (int it = 0; it < C::IT12; ++it) fft_pass2<F>(tw2r, p2 + it * (kGroupThreads >> 4));
        }
        group_bar(g);
        // ---------------- pass 3: last radix + real split (+ magnitude) ----------------
#pragma unroll 1
        for (int fl = 0; fl < FPG; ++fl) {
          const int frame = f + fl;
          if (frame >= f1) break;
          const float2 *fbuf = buf + fl * C::BUF;
          if (MODE == MODE_LOGFILT) {
            float *mags = s_mags + (sub + fl) * MS;
            if (u != 0)
              fft_pass3_unit<F>(u, fbuf + pa_off, fbuf + pb_off, s_tw3 + u, s_pt + u,
                                [&](int k, float2 X) { if (TB > 1 || k < kcap) mags[k] = cabs_fast(X); });
            if (tid < 2 * R3) {   // the two self-paired columns: one bin per lane of warp 0
              int bin;
              const float2 X = fft_pass3_selfpaired<F>(tid, fbuf, s_wr, s_pt, bin);
              if (TB > 1 || bin < kcap) mags[bin] = cabs_fast(X);
            }
          } else if (frame >= f0) {
            const long long row = row0 + frame;
            if (p.spec_complex) {
              float2 *o = reinterpret_cast<float2 *>(p.spec_out) + row * p.spec_ld;
              if (u != 0)
                fft_pass3_unit<F>(u, fbuf + pa_off, fbuf + pb_off, s_tw3 + u, s_pt + u,
                                  [&](int k, float2 X) { o[k] = (p.circular_shift && (k & 1)) ? make_float2(-X.x, -X.y) : X; });
              if (tid < 2 * R3) {
                int bin;
                float2 Xm;
                const float2 X = fft_pass3_selfpaired<F>(tid, fbuf, s_wr, s_pt, bin, Xm);
                o[bin] = (p.circular_shift && (bin & 1)) ? make_float2(-X.x, -X.y) : X;
                if (tid == 0 && p.spec_ld > N) o[N] = Xm;     // Nyquist bin: the mirror of bin 0 (N is even: no sign flip)
              }
            } else {
              float *o = p.spec_out + row * p.spec_ld;
              if (u != 0)
                fft_pass3_unit<F>(u, fbuf + pa_off, fbuf + pb_off, s_tw3 + u, s_pt + u,
                                  [&](int k, float2 X) { o[k] = cabs_fast(X); });
              if (tid < 2 * R3) {
                int bin;
                float2 Xm;
                const float2 X = fft_pass3_selfpaired<F>(tid, fbuf, s_wr, s_pt, bin, Xm);
                o[bin] = cabs_fast(X);
                if (tid == 0 && p.spec_ld > N) o[N] = cabs_fast(Xm);
              }
            }
          }
        }
        group_bar(g);   // pass-3 reads done: buf is free again; magnitudes visible
      }
      // =============== tail for the TB frames of this batch, TBF at a time ===============
      if (MODE == MODE_LOGFILT) hslot = front_tail<TB, TBF, MagLinear<MS>>(p, tctx, fb, f0, f1, row0, hslot, cscale, nonfinite);
    }
    if (MODE == MODE_LOGFILT && p.clip_status != nullptr && nonfinite) atomicOr(p.clip_status + c, 1);
  }
}

#endif  // __CUDACC__

}  // namespace b2
