// fb_core.cuh -- the two thread-level pieces of the fused filterbank stage (K2): slab sums and the
// per-band gather.  __host__ __device__ so tests/emu can run the exact thread mapping on the CPU
// (test infrastructure only; the product path is the CUDA build).  Packing: fb_pack.h.
//
// Replaces np.dot(spec, filterbank) of madmom.audio.spectrogram.FilteredSpectrogram (reached from
// /root/reference/backend/app/services/grid/beats.py:74).
#pragma once
#include "fft_core.cuh"

#if !defined(__CUDACC__)
struct float4 {
  float x, y, z, w;
};
struct int4 {
  int x, y, z, w;
};
static inline float4 make_float4(float x, float y, float z, float w) {
  float4 r;
  r.x = x; r.y = y; r.z = z; r.w = w;
  return r;
}
#endif

namespace b2 {

// ---- where the magnitude of (bin k, frame t of the batch) lives ------------------------------------
// MagLinear: frame t is a row of MS floats (k_front, and k_front_pair for frames <= 2048).
template <int MS_>
struct MagLinear {
  static constexpr int MS = MS_;
  static B2_HD int batch_offset(int h) { return h * MS_; }
  template <int TBF>
  static B2_HD void load(const float *mags, int k, float (&m)[TBF]) {
#pragma unroll
    for (int t = 0; t < TBF; ++t) m[t] = mags[t * MS + k];
  }
};
// MagInterleaved: the TBF frames of a filterbank call are interleaved per bin (bin k of frame t at k * TBF + t), so
// pass 3 of the pair kernel stores both frames of a pair with one 64-bit store and the filterbank fetches
// all frames of a bin with one vector load (k_front_pair; TBF = 2 or 4).  A tail batch of TB > TBF frames is
// a row of such blocks: sub-batch h (a multiple of TBF) starts h * MS floats in.
template <int TBF_, int MS_>
struct MagInterleaved {
  static constexpr int MS = MS_;
  static B2_HD int batch_offset(int h) { return h * MS_; }
  template <int TBF>
  static B2_HD void load(const float *mags, int k, float (&m)[TBF]) {
    static_assert(TBF == TBF_ && (TBF == 2 || TBF == 4), "one vector per bin");
    if (TBF == 2) {
      const float2 v = *reinterpret_cast<const float2 *>(mags + k * 2);
      m[0] = v.x;
      m[1] = v.y;
    } else {
      const float4 v = *reinterpret_cast<const float4 *>(mags + k * 4);
      m[0] = v.x;
      m[1] = v.y;
      m[TBF > 2 ? 2 : 0] = v.z;
      m[TBF > 2 ? 3 : 0] = v.w;
    }
  }
};

// Layout of the partial sums.  Up to two frames per batch: the frames of one partial sum are adjacent
// (s_part[(4 * slab + r) * TBF + t]) and the band stage fetches them with one vector load; four frames: one
// plane per frame (s_part[t * pstride + 4 * slab + r]) -- measured on B200 the vector form wins 1.3 % at
// frame 4096 (TBF = 2) and loses 5 % at frame 1024 (TBF = 4).
template <int TBF>
struct PartialLayout {
  static constexpr bool ADJACENT = TBF <= 2;
};

// TBF adjacent floats (the frames of one partial sum) added with one vector load
template <int TBF>
struct FrameVec;
template <>
struct FrameVec<1> {
  static B2_HD void add(const float *p, float (&y)[1]) { y[0] += p[0]; }
};
template <>
struct FrameVec<2> {
  static B2_HD void add(const float *p, float (&y)[2]) {
    const float2 v = *reinterpret_cast<const float2 *>(p);
    y[0] += v.x;
    y[1] += v.y;
  }
};
template <>
struct FrameVec<4> {
  static B2_HD void add(const float *p, float (&y)[4]) {
    const float4 v = *reinterpret_cast<const float4 *>(p);
    y[0] += v.x;
    y[1] += v.y;
    y[2] += v.z;
    y[3] += v.w;
  }
};

// Thread `tid` walks the L consecutive bins of each of its slabs once, for TBF frames at a time:
// one 128-bit weight load per bin (shared by the frames), one magnitude load per bin and frame, four
// FMAs per bin and frame.  Neighbouring lanes sit L (odd) bins apart: no bank conflicts, no
// descriptors, no data-dependent control flow.  s_part[(4 * slab + r) * TBF + t] receives the sum of frame t
// for the band with index r modulo 4.
// W4G: the weight table did not fit in shared memory and is read from global memory (read-only path).
// POW: the filterbank works on the power spectrum: every magnitude is squared as it is read.
template <int L, int TBF, class MA, bool W4G = false, bool POW = false>
B2_HD void fb_slabs(const float4 *s_w4, const float *s_mags, float *s_part, int ns, int kmin, int pstride, int tid,
                    int cap = MA::MS) {   // cap: floats of a magnitude row that exist (<= MS)
  for (int s = 0; s < ns; ++s) {
    const int g = s * kGroupThreads + tid;
    int k0 = kmin + g * L;
    if (k0 > cap - L) k0 = cap - L;              // slabs past the kept spectrum carry zero weights
    const float4 *wp = s_w4 + s * L * kGroupThreads + tid;
    float4 acc[TBF];
#pragma unroll
    for (int t = 0; t < TBF; ++t) acc[t] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int i = 0; i < L; ++i) {
#if defined(__CUDA_ARCH__)
      const float4 w = W4G ? __ldg(wp + i * kGroupThreads) : wp[i * kGroupThreads];
#else
      const float4 w = wp[i * kGroupThreads];
#endif
      float mm[TBF];
      MA::load(s_mags, k0 + i, mm);
#pragma unroll
      for (int t = 0; t < TBF; ++t) {
        float m = mm[t];
        if (POW) m *= m;
        acc[t].x = fmaf(w.x, m, acc[t].x);
        acc[t].y = fmaf(w.y, m, acc[t].y);
        acc[t].z = fmaf(w.z, m, acc[t].z);
        acc[t].w = fmaf(w.w, m, acc[t].w);
      }
    }
    if (PartialLayout<TBF>::ADJACENT) {
      float flat[4 * TBF];
#pragma unroll
      for (int t = 0; t < TBF; ++t) {
        flat[0 * TBF + t] = acc[t].x;
        flat[1 * TBF + t] = acc[t].y;
        flat[2 * TBF + t] = acc[t].z;
        flat[3 * TBF + t] = acc[t].w;
      }
      float4 *out = reinterpret_cast<float4 *>(s_part + (size_t)g * 4 * TBF);
#pragma unroll
      for (int c = 0; c < TBF; ++c) out[c] = make_float4(flat[4 * c], flat[4 * c + 1], flat[4 * c + 2], flat[4 * c + 3]);
    } else {
#pragma unroll
      for (int t = 0; t < TBF; ++t) reinterpret_cast<float4 *>(s_part + t * pstride)[g] = acc[t];
    }
  }
}

template <int TBF, class MA, bool POW = false>
B2_HD void fb_slabs_dispatch(int L, const float4 *s_w4, const float *s_mags, float *s_part, int ns, int kmin,
                             int pstride, int tid, int cap = MA::MS) {
  switch (L) {
    case 3: fb_slabs<3, TBF, MA, false, POW>(s_w4, s_mags, s_part, ns, kmin, pstride, tid, cap); break;
    case 5: fb_slabs<5, TBF, MA, false, POW>(s_w4, s_mags, s_part, ns, kmin, pstride, tid, cap); break;
    case 7: fb_slabs<7, TBF, MA, false, POW>(s_w4, s_mags, s_part, ns, kmin, pstride, tid, cap); break;
    case 9: fb_slabs<9, TBF, MA, false, POW>(s_w4, s_mags, s_part, ns, kmin, pstride, tid, cap); break;
    case 11: fb_slabs<11, TBF, MA, false, POW>(s_w4, s_mags, s_part, ns, kmin, pstride, tid, cap); break;
    case 13: fb_slabs<13, TBF, MA, false, POW>(s_w4, s_mags, s_part, ns, kmin, pstride, tid, cap); break;
    default: fb_slabs<15, TBF, MA, false, POW>(s_w4, s_mags, s_part, ns, kmin, pstride, tid, cap); break;
  }
}

// band sums of TBF frames for the band described by bd (FbBand in fb_pack.h)
template <int TBF, class MA>
B2_HD void fb_band_sum(const int4 bd, const float *s_part, int pstride, const float *s_mags, const float *s_dw,
                       float (&ysum)[TBF]) {
#pragma unroll
  for (int t = 0; t < TBF; ++t) ysum[t] = 0.f;
  if (PartialLayout<TBF>::ADJACENT) {            // slab band: partial sums of the slabs it touches
    const float *pp = s_part + bd.x * TBF;
    for (int i = 0; i < bd.y; ++i) FrameVec<TBF>::add(pp + 4 * TBF * i, ysum);
  } else {
    const float *pp = s_part + bd.x;
    for (int i = 0; i < bd.y; ++i) {
#pragma unroll
      for (int t = 0; t < TBF; ++t) ysum[t] += pp[t * pstride + 4 * i];
    }
  }
  for (int i = 0; i < bd.w; ++i) {               // direct band: its few taps straight from the magnitudes
    const float w = s_dw[bd.x + i];
    float mm[TBF];
    MA::load(s_mags, bd.z + i, mm);
#pragma unroll
    for (int t = 0; t < TBF; ++t) ysum[t] = fmaf(w, mm[t], ysum[t]);
  }
}

}  // namespace b2
