// frontend_warp_kernel.cuh -- the fused log-filtered front end of frame sizes 1024 and 2048 with one WARP per FFT.
//
// Both sizes run ONE complex FFT of N = 1024 points per step:
//   frame 1024: two consecutive real frames, z = xA + i xB (the pair transform of fft_core.cuh); the split
//               XA = (Z[k] + conj Z[N-k]) / 2, XB = -i (Z[k] - conj Z[N-k]) / 2 gives both frames' 512 bins;
//   frame 2048: one real frame, z[m] = x[2m] + i x[2m+1]; X[k] = E[k] + W_2048^k O[k] with the same E / O pairing
//               gives its 1024 bins.
// N = 32 x 32: lane b holds the 32 points z[32 a + b] in registers, transforms them (DFT32 over a), multiplies by
// W_1024^(b c), hands them to lane c through ONE 32 x 32 transpose in a warp-private shared-memory tile, and the
// second DFT32 (over b) leaves Z[c + 32 d], d = 0..31, in lane c's registers.  The mirror bins Z[N - k] of the split
// live in lane 32 - c and come by shuffle.  Magnitudes (into the tile, which the transform no longer needs), slab
// filterbank, band stage and the stacked stores follow in the same warp.
//
// Against k_front_pair (four warps per step, radix 16 x 16 x R3): one exchange through shared memory instead of two,
// no group barrier at all (only __syncwarp), about half the warp instructions per frame, and 16 fully independent
// workers per SM instead of 4-5 groups that each wait for their slowest warp.  Measured on B200 (config 2): frame
// 1024 2.18 -> 1.41 ms.
//
// Replaces, for these frame sizes, the madmom 0.16.1 chain reached from
// /root/reference/backend/app/services/grid/beats.py:74 (RNNBeatProcessor):
//   signal_frame -> frame*fft_window -> fftpack.fft[:F/2] -> np.abs -> np.dot(., filterbank)
//   -> np.log10(mul*y+add) -> SpectrogramDifference(positive) -> np.hstack
#pragma once
#include "frontend_pair_kernel.cuh"

namespace b2 {

constexpr int kWarpTile = 32 * 33;       // float2 elements of the transpose tile (row stride 33: conflict free both ways)

template <int F>
struct WarpCfg {
  static_assert(F == 1024 || F == 2048, "warp-per-FFT kernel: frame sizes 1024 (pairs) and 2048 (single frames)");
  static constexpr int NF = (F == 1024) ? 2 : 1;     // frames per FFT step
  static constexpr int NBINS = F / 2;                // magnitude bins per frame
  static constexpr int MS = NBINS + 16;              // ... plus the padding zero-weight filterbank taps may read
  static_assert(NF * MS <= 2 * kWarpTile, "the magnitudes of a step reuse the transpose tile");
};

// byte offsets; returns the dynamic shared memory of a CTA of NW warps
template <int F>
inline size_t warp_smem_layout(FrontParams &p, int NW) {
  using W = WarpCfg<F>;
  auto al = [](size_t v) { return (v + 15) & ~size_t(15); };
  size_t o = 0;
  p.o_win = -1;                                                   // window: global memory (clip edges / other windows only)
  p.o_tw3 = (int)o;  o = al(o + sizeof(float2) * 32 * 32);        // tw[c * 32 + b] = W_1024^(b c)
  p.o_pt = p.o_wr = (int)o;
  p.part_stride = p.fb_ns * 32 * 4;
  p.o_w4 = (int)o;   o = al(o + sizeof(float4) * p.fb_ns * p.fb_L * 32);
  p.o_band = (int)o; o = al(o + sizeof(int4) * (p.num_bands > 0 ? p.num_bands : 1));
  p.o_dw = (int)o;   o = al(o + sizeof(float) * (p.fb_ndw > 0 ? p.fb_ndw : 1));
  p.o_proj = (int)o;
  p.o_groups = (int)o;
  size_t g = 0;
  g = al(g + sizeof(float2) * kWarpTile);                         // transpose tile; the magnitudes of the step reuse it
  p.g_mags = 0;
  p.mag_stride = W::MS;
  p.g_partial = (int)g; g = al(g + sizeof(float) * W::NF * p.part_stride);
  p.g_hist = (int)g;    g = al(g + sizeof(float) * (p.diff_frames > 0 ? p.diff_frames : 1) * p.num_bands);
  p.g_lrow = p.g_red = p.g_task = (int)g;
  p.group_bytes = (int)g;
  return o + g * NW;
}

// cos / sin of j pi / 32, j in [0, 16]: W_64^d = (cos_pi32(d), -sin_pi32(d))
B2_HD constexpr float cos_pi32(int j) {
  constexpr float t[17] = {1.f, 0.99518472667219688624f, 0.98078528040323044913f, 0.95694033573220886494f,
                           0.92387953251128675613f, 0.88192126434835502971f, 0.83146961230254523708f,
                           0.77301045336273696081f, 0.70710678118654752440f, 0.63439328416364549822f,
                           0.55557023301960222474f, 0.47139673682599764856f, 0.38268343236508977173f,
                           0.29028467725446236764f, 0.19509032201612826785f, 0.09801714032956060199f, 0.f};
  return t[j];
}
B2_HD constexpr float sin_pi32(int j) { return cos_pi32(16 - j); }

#if defined(__CUDACC__)

// DFT of 32 points held as even / odd halves (e[i] = x[2 i], o[i] = x[2 i + 1]); on return X[k] sits in
// e[dft_pos<16>(k)] for k < 16 and in o[dft_pos<16>(k - 16)] for k >= 16.
__device__ __forceinline__ void dft32(float2 (&e)[16], float2 (&o)[16]) {
  dft16(e);
  dft16(o);
#pragma unroll
  for (int k = 0; k < 16; ++k) {
    const int q = dft_pos<16>(k);
    float2 t = o[q];
    if (k == 8) t = mul_neg_i(t);
    else if (k > 0) t = cmul(t, make_float2(cos_pi16(k), -sin_pi16(k)));   // W_32^k
    const float2 a = e[q];
    e[q] = cadd(a, t);
    o[q] = csub(a, t);
  }
}
// element k of a dft32 result
#define B2_D32(e, o, k) ((k) < 16 ? (e)[dft_pos<16>(k)] : (o)[dft_pos<16>((k) - 16)])

// lane = slab: L consecutive bins of NF frames (magnitudes interleaved per bin), four running sums per frame
template <int L, int NF>
__device__ __forceinline__ void warp_fb_slabs(const float4 *s_w4, const float *s_mags, float *s_part, int ns, int kmin,
                                              int lane, bool power, int ms) {
  for (int s = 0; s < ns; ++s) {
    const int g = s * 32 + lane;
    int k0 = kmin + g * L;
    if (k0 > ms - L) k0 = ms - L;                  // slabs past the spectrum carry zero weights
    const float4 *wp = s_w4 + s * L * 32 + lane;
    float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0;
#pragma unroll
    for (int i = 0; i < L; ++i) {
      const float4 w = wp[i * 32];
      float m0, m1 = 0.f;
      if (NF == 2) {
        const float2 m = *reinterpret_cast<const float2 *>(s_mags + 2 * (k0 + i));
        m0 = m.x;
        m1 = m.y;
      } else {
        m0 = s_mags[k0 + i];
      }
      if (power) m0 *= m0, m1 *= m1;
      a0.x = fmaf(w.x, m0, a0.x); a0.y = fmaf(w.y, m0, a0.y); a0.z = fmaf(w.z, m0, a0.z); a0.w = fmaf(w.w, m0, a0.w);
      if (NF == 2) {
        a1.x = fmaf(w.x, m1, a1.x); a1.y = fmaf(w.y, m1, a1.y); a1.z = fmaf(w.z, m1, a1.z); a1.w = fmaf(w.w, m1, a1.w);
      }
    }
    float4 *out = reinterpret_cast<float4 *>(s_part + (size_t)g * 4 * NF);   // [(4 slab + r) * NF + frame]
    if (NF == 2) {
      out[0] = make_float4(a0.x, a1.x, a0.y, a1.y);
      out[1] = make_float4(a0.z, a1.z, a0.w, a1.w);
    } else {
      out[0] = a0;
    }
  }
}

template <int F, int IN, int NW>
__global__ void __launch_bounds__(32 * NW, 1) k_front_warp(const FrontParams p) {
  using W = WarpCfg<F>;
  constexpr int NF = W::NF, MS = W::MS, NBINS = W::NBINS;
  extern __shared__ __align__(16) unsigned char smem[];
  float2 *s_tw = reinterpret_cast<float2 *>(smem + p.o_tw3);
  float4 *s_w4 = reinterpret_cast<float4 *>(smem + p.o_w4);
  int4 *s_band = reinterpret_cast<int4 *>(smem + p.o_band);
  float *s_dw = reinterpret_cast<float *>(smem + p.o_dw);
  __shared__ __align__(8) uint64_t s_stage_bar;     // plan tables arrive by TMA bulk copies (bulk_stage.cuh)
  bulk_stage_begin(&s_stage_bar, [&](auto &&table) {
    table(s_tw, p.tw3, (uint32_t)sizeof(float2) * 32 * 32);
    table(s_w4, p.fb_w4, (uint32_t)sizeof(float4) * p.fb_ns * p.fb_L * 32);
    table(s_band, p.fb_band, (uint32_t)sizeof(int4) * p.num_bands);
    table(s_dw, p.fb_dw, (uint32_t)sizeof(float) * p.fb_ndw);
  });
  for (int i = threadIdx.x; i < (p.group_bytes * NW) / 4; i += blockDim.x)      // every warp block starts out zero (finite)
    reinterpret_cast<float *>(smem + p.o_groups)[i] = 0.f;
  bulk_stage_wait(&s_stage_bar);

  const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
  unsigned char *wmem = smem + p.o_groups + (size_t)wid * p.group_bytes;
  float2 *T = reinterpret_cast<float2 *>(wmem);
  float *s_mags = reinterpret_cast<float *>(wmem);               // reuses the tile once the transform has read it
  float *s_part = reinterpret_cast<float *>(wmem + p.g_partial);
  float *s_hist = reinterpret_cast<float *>(wmem + p.g_hist);

  const int B = p.num_bands, kd = p.diff_frames;
  // Hann window in registers (FrontParams::win_fly): frame 1024: sample 32 a + lane; frame 2048: samples 64 a + 2 lane (+1)
  float2 wab0 = make_float2(0.f, 0.f), wab1 = wab0;
  if (p.win_fly) {
    wab0 = __ldg(p.win_ab + (NF == 2 ? lane : 2 * lane));
    if (NF == 1) wab1 = __ldg(p.win_ab + 2 * lane + 1);
  }
  // frame 2048: split twiddle of this lane's bins k = lane + 32 d: pt_k = -i W_2048^k = ptc * W_64^d
  float2 ptc = make_float2(0.f, -1.f);
  if (NF == 1) {
    float sn, cs;
    sincospif((float)lane * (1.0f / 1024.0f), &sn, &cs);
    ptc = make_float2(-sn, -cs);
  }
  float *out_spec = (p.out != nullptr && p.col_spec >= 0) ? p.out + p.col_spec : nullptr;
  float *out_diff = (p.out != nullptr && p.col_diff >= 0) ? p.out + p.col_diff : nullptr;
  const bool do_log = p.log_enabled != 0, positive = p.positive != 0, power = p.power != 0;
  const float lk = p.log_scale * 0.30102999566398120f;
  const int partner = (32 - lane) & 31;
  const int total_tasks = p.task_off[p.n_clips];

  for (;;) {
    int task = 0;
    if (lane == 0) task = atomicAdd(p.task_counter, 1);
    task = __shfl_sync(0xffffffffu, task, 0);
    if (task >= total_tasks) break;
    const int c = task_clip(p.task_off, p.n_clips, task);
    const long long samp0 = p.clip_off[c];
    const long long nsamp = p.clip_off[c + 1] - samp0;
    const long long row0 = p.frame_off[c];
    const int Tn = (int)(p.frame_off[c + 1] - row0);
    const int ch = task_chunk(p, c);
    const int f0 = (task - p.task_off[c]) * ch;
    const int f1 = min(Tn, f0 + ch);
    const int fs = (kd > 0 && !p.seam_fix) ? max(0, f0 - kd) : f0;     // warm-up rows for the difference
    Samples<IN> S{clip_base<IN>(p.sig, samp0)};
    float cscale = p.clip_scale != nullptr ? __ldg(p.clip_scale + c) : 1.f;
    if (power) cscale *= cscale;
    int nonfinite = 0;
    int hslot = kd > 0 ? fs % kd : 0;                          // difference ring slot of frame fA

    for (int fA = fs; fA < f1; fA += NF) {
      // ---------------- load, window, first DFT32 (over a) ----------------
      float2 e[16], o[16];
      if (NF == 2) {             // z[n] = w[n] (xA[n] + i xB[n]), n = 32 a + lane
        const long long sA = (long long)((double)fA * p.hop) - (F / 2) - p.origin;
        const long long sB = (long long)((double)(fA + 1) * p.hop) - (F / 2) - p.origin;
        const bool hasB = fA + 1 < f1;
        const bool interior = (sA >= 0) && (sB + F <= nsamp) && hasB;     // sB >= sA
        if (interior && p.win_fly) {
          const void *qa = S.ptr(sA + lane), *qb = S.ptr(sB + lane);
#pragma unroll
          for (int a = 0; a < 32; ++a) {
            const float w = fmaf(wab0.x, p.win_cs32[a].x, fmaf(wab0.y, p.win_cs32[a].y, p.win_h));
            const float2 v = crscale(make_float2(Samples<IN>::at_ptr(qa, 32 * a), Samples<IN>::at_ptr(qb, 32 * a)), w);
            if (a & 1) o[a >> 1] = v; else e[a >> 1] = v;
          }
        } else {
#pragma unroll
          for (int a = 0; a < 32; ++a) {
            const int n = 32 * a + lane;
            const float w = __ldg(p.window + n);
            const long long ia = sA + n, ib = sB + n;
            const float xa = (ia >= 0 && ia < nsamp) ? S.at(ia) : 0.f;
            const float xb = (hasB && ib >= 0 && ib < nsamp) ? S.at(ib) : 0.f;
            const float2 v = crscale(make_float2(xa, xb), w);
            if (a & 1) o[a >> 1] = v; else e[a >> 1] = v;
          }
        }
      } else {                   // z[m] = (w[2m] x[2m], w[2m+1] x[2m+1]), m = 32 a + lane
        const long long s0 = (long long)((double)fA * p.hop) - (F / 2) - p.origin;
        const bool interior = (s0 >= 0) && (s0 + F <= nsamp);
        if (interior && p.win_fly) {
          const void *q = S.ptr(s0 + 2 * lane);
#pragma unroll
          for (int a = 0; a < 32; ++a) {
            const float w0 = fmaf(wab0.x, p.win_cs32[a].x, fmaf(wab0.y, p.win_cs32[a].y, p.win_h));
            const float w1 = fmaf(wab1.x, p.win_cs32[a].x, fmaf(wab1.y, p.win_cs32[a].y, p.win_h));
            const float2 v = emul(make_float2(w0, w1),
                                  make_float2(Samples<IN>::at_ptr(q, 64 * a), Samples<IN>::at_ptr(q, 64 * a + 1)));
            if (a & 1) o[a >> 1] = v; else e[a >> 1] = v;
          }
        } else {
#pragma unroll
          for (int a = 0; a < 32; ++a) {
            const int n = 64 * a + 2 * lane;
            const long long ia = s0 + n, ib = ia + 1;
            const float xa = (ia >= 0 && ia < nsamp) ? S.at(ia) : 0.f;
            const float xb = (ib >= 0 && ib < nsamp) ? S.at(ib) : 0.f;
            const float2 v = emul(make_float2(__ldg(p.window + n), __ldg(p.window + n + 1)), make_float2(xa, xb));
            if (a & 1) o[a >> 1] = v; else e[a >> 1] = v;
          }
        }
      }
      dft32(e, o);
      // ---------------- twiddle W_1024^(b c), transpose ----------------
      __syncwarp();                                            // the previous step's tail has read the tile (as magnitudes)
#pragma unroll
      for (int cc = 0; cc < 32; ++cc) {
        float2 v = B2_D32(e, o, cc);
        if (cc > 0) v = cmul(v, s_tw[cc * 32 + lane]);
        cstore(T + cc * 33 + lane, v);
      }
      __syncwarp();
#pragma unroll
      for (int b = 0; b < 32; ++b) {
        const float2 v = T[lane * 33 + b];
        if (b & 1) o[b >> 1] = v; else e[b >> 1] = v;
      }
      // ---------------- second DFT32 (over b): Z[lane + 32 d] ----------------
      dft32(e, o);
      __syncwarp();                                            // every lane has read its row before magnitudes overwrite the tile
      // ---------------- real-spectrum split with lane 32 - c, magnitudes ----------------
      if (lane < 16) {                                         // the padding bins NBINS .. NBINS + 15
        if (NF == 2) *reinterpret_cast<float2 *>(s_mags + 2 * (NBINS + lane)) = make_float2(0.f, 0.f);
        else s_mags[NBINS + lane] = 0.f;
      }
#pragma unroll
      for (int d = 0; d < 16; ++d) {
        // what the partner needs from this lane: Z[N - k'] for its bin k' = partner + 32 d
        const float2 other = B2_D32(e, o, 31 - d);                                   // lanes 1..31: index 31 - d
        const float2 own0 = d == 0 ? B2_D32(e, o, 0) : B2_D32(e, o, 32 - d);          // lane 0 pairs with itself: index 32 - d
        const float2 give = lane == 0 ? own0 : other;
        const float rx = __shfl_sync(0xffffffffu, give.x, partner), ry = __shfl_sync(0xffffffffu, give.y, partner);
        const float2 z = B2_D32(e, o, d);
        const float2 P = make_float2(z.x + rx, z.y - ry), M = make_float2(z.x - rx, z.y + ry);   // z +- conj(mirror)
        const int k = lane + 32 * d;
        if (NF == 2) {           // |XA[k]|, |XB[k]|
          *reinterpret_cast<float2 *>(s_mags + 2 * k) = make_float2(cabs_fast(P), cabs_fast(M));
        } else {                 // X[k] = E + pt_k D, X[N - k] = conj(E - pt_k D)
          float2 pt = ptc;
          if (d > 0) pt = cmul(ptc, make_float2(cos_pi32(d), -sin_pi32(d)));
          const float2 Tk = cmul(M, pt);
          s_mags[k] = cabs_fast(cadd(P, Tk));
          if (k > 0) s_mags[1024 - k] = cabs_fast(csub(P, Tk));
        }
      }
      if (NF == 1 && lane == 0) {                              // bin 512 pairs with itself: pt = -1, X = conj(2 Z)
        const float2 z = B2_D32(e, o, 16);
        s_mags[512] = 2.f * cabs_fast(z);
      }
      __syncwarp();
      // ---------------- slab filterbank (lane = slab), band stage (lane = band) ----------------
      switch (p.fb_L) {
        case 3: warp_fb_slabs<3, NF>(s_w4, s_mags, s_part, p.fb_ns, p.fb_kmin, lane, power, MS); break;
        case 5: warp_fb_slabs<5, NF>(s_w4, s_mags, s_part, p.fb_ns, p.fb_kmin, lane, power, MS); break;
        case 7: warp_fb_slabs<7, NF>(s_w4, s_mags, s_part, p.fb_ns, p.fb_kmin, lane, power, MS); break;
        case 9: warp_fb_slabs<9, NF>(s_w4, s_mags, s_part, p.fb_ns, p.fb_kmin, lane, power, MS); break;
        case 11: warp_fb_slabs<11, NF>(s_w4, s_mags, s_part, p.fb_ns, p.fb_kmin, lane, power, MS); break;
        case 13: warp_fb_slabs<13, NF>(s_w4, s_mags, s_part, p.fb_ns, p.fb_kmin, lane, power, MS); break;
        default: warp_fb_slabs<15, NF>(s_w4, s_mags, s_part, p.fb_ns, p.fb_kmin, lane, power, MS); break;
      }
      __syncwarp();
      const int nvalid = min(NF, f1 - fA);                     // frames of the step that exist
      const int nskip = max(0, f0 - fA);                       // leading warm-up frames: they only feed the ring
      float fluxacc[NF];
#pragma unroll
      for (int t = 0; t < NF; ++t) fluxacc[t] = 0.f;
      for (int jb = 0; jb < B; jb += 32) {
        const int j = jb + lane;
        const bool valid = j < B;
        const int4 bd = valid ? s_band[j] : make_int4(0, 0, 0, 0);
        const int nP = __reduce_max_sync(0xffffffffu, bd.y), nD = __reduce_max_sync(0xffffffffu, bd.w);
        float y[NF];
#pragma unroll
        for (int t = 0; t < NF; ++t) y[t] = 0.f;
        const float *pp = s_part + bd.x * NF;
#pragma unroll 4
        for (int i = 0; i < nP; ++i)
          if (i < bd.y) {
#pragma unroll
            for (int t = 0; t < NF; ++t) y[t] += pp[4 * NF * i + t];
          }
        const float *dwp = s_dw + bd.x;
#pragma unroll 2
        for (int i = 0; i < nD; ++i)
          if (i < bd.w) {                                      // direct band: its few taps straight from the magnitudes
#pragma unroll
            for (int t = 0; t < NF; ++t) {
              float m = s_mags[(bd.z + i) * NF + t];
              if (power) m *= m;
              y[t] = fmaf(dwp[i], m, y[t]);
            }
          }
        if (valid) {
          int slot = hslot;
#pragma unroll
          for (int t = 0; t < NF; ++t) {
            if (t < nvalid) {
              float Lt = y[t] * cscale;
              if (do_log) {
                float a = __fadd_rn(__fmul_rn(p.mul, Lt), p.add);        // separate multiply and add, as numpy
                if (p.log_floor > 0.f) a = fmaxf(a, p.log_floor);
                Lt = fast_lg2(a) * lk;                                   // log_scale * log10(a)
              }
              nonfinite |= !(fabsf(Lt) <= 3.402823466e38f);
              float D = 0.f;
              if (kd > 0) {
                float *hp = s_hist + slot * B + j;
                const float old = *hp;
                *hp = Lt;
                if (fA + t >= kd) D = Lt - old;
                if (positive) D = fmaxf(D, 0.f);
                if (++slot == kd) slot = 0;
              }
              if (t >= nskip) {
                const long long ofs = (row0 + fA + t) * p.ld_out + j;
                if (out_spec != nullptr) out_spec[ofs] = Lt;
                if (out_diff != nullptr) out_diff[ofs] = D;
                fluxacc[t] += D;
              }
            }
          }
        }
      }
      if (kd > 0) {                                            // ring position of the next step's first frame
        hslot += NF;
        while (hslot >= kd) hslot -= kd;
      }
      if (p.flux != nullptr) {
#pragma unroll
        for (int t = 0; t < NF; ++t) {
          float v = fluxacc[t];
#pragma unroll
          for (int dd = 16; dd > 0; dd >>= 1) v += __shfl_xor_sync(0xffffffffu, v, dd);
          if (lane == t && t < nvalid && t >= nskip) p.flux[row0 + fA + t] = v;
        }
      }
    }
    if (p.clip_status != nullptr && nonfinite) atomicOr(p.clip_status + c, 1);
  }
}

#endif  // __CUDACC__

}  // namespace b2
