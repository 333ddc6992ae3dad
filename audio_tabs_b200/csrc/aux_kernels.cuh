// aux_kernels.cuh -- small non-template kernels (task table, stand-alone stages).  Included by
// b200spec.cu only.
#pragma once
#include "frontend_kernel.cuh"

namespace b2 {

// ---- task table: tasks per clip = ceil(frames / chunk), exclusive prefix -> task_off -----------
// The last tail_clips clips are cut into tasks of chunk_small frames (FrontParams::tail_clips; 0 = one size).
__global__ void k_setup_tasks(const long long *__restrict__ frame_off, int n_clips, int chunk, int chunk_small, int tail_clips,
                              int *__restrict__ task_off, int *__restrict__ task_counter) {
  __shared__ int s_part[1024];
  const int t = threadIdx.x;
  const int per = (n_clips + 1023) / 1024;
  const int lo = min(t * per, n_clips), hi = min(lo + per, n_clips);
  int sum = 0;
  auto tasks_of = [&](int c) {
    const int ch = c >= n_clips - tail_clips ? chunk_small : chunk;
    return (int)((frame_off[c + 1] - frame_off[c] + ch - 1) / ch);
  };
  for (int c = lo; c < hi; ++c) sum += tasks_of(c);
  s_part[t] = sum;
  __syncthreads();
  for (int d = 1; d < 1024; d <<= 1) {  // Hillis-Steele inclusive scan
    int v = t >= d ? s_part[t - d] : 0;
    __syncthreads();
    s_part[t] += v;
    __syncthreads();
  }
  int run = s_part[t] - sum;
  for (int c = lo; c < hi; ++c) {
    task_off[c] = run;
    run += tasks_of(c);
  }
  if (t == 1023) task_off[n_clips] = s_part[1023];
  if (t == 0) *task_counter = 0;
}

// ---- per-clip peak: max |x| of the down-mixed samples ------------------------------------------------
// grid = kPeakBlocksPerClip * n_clips blocks (clip = blockIdx.x / kPeakBlocksPerClip: no 65535 limit of
// grid.y); the clip's samples are strided over its blocks; |x| >= 0 so the
// float bit pattern orders like an unsigned integer and atomicMax on it is exact.
constexpr int kPeakBlocksPerClip = 32;
template <int IN>
__global__ void k_clip_peak(const void *__restrict__ sig, const long long *__restrict__ clip_off,
                            unsigned int *__restrict__ peak_bits) {
  const int c = blockIdx.x / kPeakBlocksPerClip, part = blockIdx.x % kPeakBlocksPerClip;
  const long long s0 = clip_off[c], n = clip_off[c + 1] - s0;
  Samples<IN> S{clip_base<IN>(sig, s0)};
  float m = 0.f;
  for (long long i = part * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)kPeakBlocksPerClip * blockDim.x)
    m = fmaxf(m, fabsf(S.at(i)));
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, d));
  __shared__ float s_m[32];
  if ((threadIdx.x & 31) == 0) s_m[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x < 32) {
    m = threadIdx.x < (blockDim.x >> 5) ? s_m[threadIdx.x] : 0.f;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, d));
    if (threadIdx.x == 0) atomicMax(peak_bits + c, __float_as_uint(m));
  }
}

__global__ void k_peak_reciprocal(float *__restrict__ peak, int n, float eps, float numer) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) peak[i] = numer / (peak[i] + eps);
}

// ---- onset-strength envelope (librosa.onset.onset_strength on rows of a (T, B) matrix) ------------------
// maximum of each clip's rows (power_to_db's top_db clip is relative to it): grid = kRowmaxBlocksPerClip *
// n_clips blocks, one warp per row at a time, block result merged with an ordered-float atomic max.
// clip_max must hold -inf on entry (k_fill_neg_inf).
constexpr int kRowmaxBlocksPerClip = 16;
__global__ void k_fill_neg_inf(float *__restrict__ x, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) x[i] = -INFINITY;
}
__device__ __forceinline__ void atomic_max_float(float *addr, float v) {
  if (v >= 0.f) atomicMax(reinterpret_cast<int *>(addr), __float_as_int(v));          // non-negative floats order as ints
  else atomicMin(reinterpret_cast<unsigned int *>(addr), __float_as_uint(v));          // negative ones in reverse as uints
}
__global__ void k_clip_rowmax(const float *__restrict__ L, long long ld_L, int B, const long long *__restrict__ frame_off,
                              float *__restrict__ clip_max) {
  const int c = blockIdx.x / kRowmaxBlocksPerClip, part = blockIdx.x % kRowmaxBlocksPerClip;
  const long long r0 = frame_off[c], rows = frame_off[c + 1] - r0;
  const int lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  float m = -INFINITY;
  for (long long r = part * nw + (threadIdx.x >> 5); r < rows; r += (long long)kRowmaxBlocksPerClip * nw) {
    const float *x = L + (r0 + r) * ld_L;
    for (int j = lane; j < B; j += 32) m = fmaxf(m, x[j]);
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, d));
  __shared__ float s_m[32];
  if (lane == 0) s_m[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x < 32) {
    m = threadIdx.x < nw ? s_m[threadIdx.x] : -INFINITY;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, d));
    if (threadIdx.x == 0 && rows > 0) atomic_max_float(clip_max + c, m);
  }
}

// Median of a row held as NPL values per lane (element e = lane * NPL + r; positions >= B hold +inf):
// bitonic sort across the warp -- exchanges with a partner in the same lane are register swaps, the
// others one shuffle per register -- then the two middle order statistics (np.median of an even count
// is their mean).
template <int NPL>
__device__ __forceinline__ float warp_median(float (&v)[NPL], int B, int lane) {
  constexpr int N = 32 * NPL;
#pragma unroll
  for (int k = 2; k <= N; k <<= 1) {
#pragma unroll
    for (int j = k >> 1; j > 0; j >>= 1) {
      if (j >= NPL) {
        const int lj = j / NPL;
#pragma unroll
        for (int r = 0; r < NPL; ++r) {
          const float other = __shfl_xor_sync(0xffffffffu, v[r], lj);
          const bool asc = ((lane * NPL + r) & k) == 0;
          const bool keep_min = ((lane & lj) == 0) == asc;
          v[r] = keep_min ? fminf(v[r], other) : fmaxf(v[r], other);
        }
      } else {
#pragma unroll
        for (int r = 0; r < NPL; ++r) {
          if ((r & j) == 0) {
            const bool asc = ((lane * NPL + r) & k) == 0;
            const float lo = fminf(v[r], v[r ^ j]), hi = fmaxf(v[r], v[r ^ j]);
            v[r] = asc ? lo : hi;
            v[r ^ j] = asc ? hi : lo;
          }
        }
      }
    }
  }
  const int k_hi = B >> 1, k_lo = (B & 1) ? k_hi : k_hi - 1;
  float a = 0.f, b = 0.f;
#pragma unroll
  for (int r = 0; r < NPL; ++r) {
    if (r == k_lo % NPL) a = v[r];
    if (r == k_hi % NPL) b = v[r];
  }
  a = __shfl_sync(0xffffffffu, a, k_lo / NPL);
  b = __shfl_sync(0xffffffffu, b, k_hi / NPL);
  return (a + b) * 0.5f;
}

// One warp per output row; a warp walks a CONTIGUOUS range of rows, so the clip lookup is done once per range and,
// for lag 1, the lagged row is the one the warp has just had in registers (half the loads).
// NPL > 0: the row lives in registers (NPL = values per lane, a power of two with 32 * NPL >= B) and the median is a
// warp bitonic sort -- skipped when more than half of the differences are 0 (they are the smallest values a positive
// difference can take, so both middle order statistics are 0 then); NPL == 0: any B <= 1024, the row goes through
// dynamic shared memory (B floats per warp) and the median is found by rank counting.
template <int NPL>
__global__ void k_onset_env(const float *__restrict__ L, long long ld_L, int B, const long long *__restrict__ frame_off,
                            int n_clips, long long rows, int lag, float top_db, int aggregate, int shift,
                            const float *__restrict__ clip_max, float *__restrict__ env) {
  extern __shared__ float s_rows[];
  constexpr int NV = NPL > 0 ? NPL : 1;
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  float *row = s_rows + (size_t)wib * B;
  const long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  const long long per_warp = (rows + nwarps - 1) / nwarps;
  const long long r0 = warp * per_warp, r1 = min(rows, r0 + per_warp);
  if (r0 >= r1) return;
  int lo = 0, hi = n_clips;
  while (hi - lo > 1) {
    int mid = (lo + hi) >> 1;
    if (frame_off[mid] <= r0) lo = mid; else hi = mid;
  }
  const bool vec4 = NPL == 4 && (ld_L & 3) == 0 && (reinterpret_cast<size_t>(L) & 15) == 0 && lane * 4 + 4 <= B;
  const int k_hi = B >> 1;
  float prev[NV];                       // clamped values of the row before (same clip), lag 1 only
  bool have_prev = false;
  for (long long r = r0; r < r1; ++r) {
    while (lo + 1 < n_clips && r >= frame_off[lo + 1]) { ++lo; have_prev = false; }
    const long long t = r - frame_off[lo];          // output index inside the clip
    const long long src = t - shift;                // row whose difference lands here
    if (src < lag) {                                // leading zeros: lag + centre shift (np.pad)
      if (lane == 0) env[r] = 0.f;
      have_prev = false;
      continue;
    }
    const float floor_db = top_db >= 0.f ? clip_max[lo] - top_db : -INFINITY;
    const float *a = L + (r - shift) * ld_L, *b = a - (long long)lag * ld_L;
    float result;
    if (NPL > 0) {
      float cur[NV], v[NV];
      if (vec4) {
        const float4 q = *reinterpret_cast<const float4 *>(a + lane * 4);
        cur[0] = q.x; cur[1 % NV] = q.y; cur[2 % NV] = q.z; cur[3 % NV] = q.w;
      } else {
#pragma unroll
        for (int i = 0; i < NV; ++i) cur[i] = (lane * NPL + i < B) ? a[lane * NPL + i] : 0.f;
      }
#pragma unroll
      for (int i = 0; i < NV; ++i) cur[i] = fmaxf(cur[i], floor_db);
      if (!(lag == 1 && have_prev)) {
        if (vec4) {
          const float4 q = *reinterpret_cast<const float4 *>(b + lane * 4);
          prev[0] = q.x; prev[1 % NV] = q.y; prev[2 % NV] = q.z; prev[3 % NV] = q.w;
        } else {
#pragma unroll
          for (int i = 0; i < NV; ++i) prev[i] = (lane * NPL + i < B) ? b[lane * NPL + i] : 0.f;
        }
#pragma unroll
        for (int i = 0; i < NV; ++i) prev[i] = fmaxf(prev[i], floor_db);
      }
      float sum = 0.f;
      int zeros = 0;
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        float d = INFINITY;                         // padding sorts to the top
        if (lane * NPL + i < B) {
          d = fmaxf(0.f, cur[i] - prev[i]);
          sum += d;
          zeros += d == 0.f;
        }
        v[i] = d;
        prev[i] = cur[i];                           // the lagged row of the next one
      }
      have_prev = true;
      if (aggregate == 0) {
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, d);
        result = sum / (float)B;
      } else if (__reduce_add_sync(0xffffffffu, zeros) > k_hi) {
        result = 0.f;                               // ranks 0 .. k_hi are all zeros
      } else {
        result = warp_median<NV>(v, B, lane);
      }
    } else {
      float sum = 0.f;
      for (int j = lane; j < B; j += 32) {
        const float d = fmaxf(0.f, fmaxf(a[j], floor_db) - fmaxf(b[j], floor_db));
        row[j] = d;
        sum += d;
      }
      if (aggregate == 0) {
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, d);
        result = sum / (float)B;
      } else {
        __syncwarp();
        // rank of every element (ties broken by index); the two middle order statistics give np.median
        const int k_lo = (B & 1) ? k_hi : k_hi - 1;
        float v_lo = 0.f, v_hi = 0.f;
        for (int j = lane; j < B; j += 32) {
          const float v = row[j];
          int rank = 0;
          for (int i = 0; i < B; ++i) {
            const float w = row[i];
            rank += (w < v) || (w == v && i < j);
          }
          if (rank == k_lo) v_lo = v;
          if (rank == k_hi) v_hi = v;
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {          // exactly one lane holds each of them; the rest hold 0 (values are >= 0)
          v_lo = fmaxf(v_lo, __shfl_xor_sync(0xffffffffu, v_lo, d));
          v_hi = fmaxf(v_hi, __shfl_xor_sync(0xffffffffu, v_hi, d));
        }
        result = (v_lo + v_hi) * 0.5f;
        __syncwarp();
      }
    }
    if (lane == 0) env[r] = result;
  }
}

// ---- DeepChroma context stacking: out[t, c*B + j] = in[t + c - context/2, j] inside the clip, else 0 ------
// SOURCE-centric: a warp takes an input row, loads its B floats ONCE into registers and stores them into the
// `context` places they have in the output (row t = s + context/2 - c, slot c), then zeroes the slots of ITS OWN
// output row whose source lies outside the clip -- every output element is written exactly once, by the warp of its
// source row or by the zero pass.  (ncu on the output-centric loops: first 2270 warp instructions per row on
// "i / B, i % B" and 64-bit row arithmetic -- issue bound at 31 % of the write bandwidth --, then, with that
// hoisted, 83 % of the stall cycles waiting for the `context` x 4 dependent loads of a row.)  A block owns a
// contiguous range of rows and its warps take them in turn, so neighbouring warps complete neighbouring output rows.
__global__ void k_context_stack(const float *__restrict__ in, long long ld_in, int B, const long long *__restrict__ frame_off,
                                int n_clips, long long rows, int context, float *__restrict__ out) {
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  const long long per_block = (rows + gridDim.x - 1) / gridDim.x;
  const long long r0 = blockIdx.x * per_block + wib, r1 = min(rows, (blockIdx.x + 1) * per_block);
  if (r0 >= r1) return;
  const int half = context / 2;
  const long long W = (long long)context * B;
  int lo = 0, hi = n_clips;
  while (hi - lo > 1) {
    int mid = (lo + hi) >> 1;
    if (frame_off[mid] <= r0) lo = mid; else hi = mid;
  }
  for (long long r = r0; r < r1; r += wpb) {
    while (lo + 1 < n_clips && r >= frame_off[lo + 1]) ++lo;
    const long long c0 = frame_off[lo], c1 = frame_off[lo + 1];
    const float *x = in + r * ld_in + lane;
    for (int jb = 0; jb < B; jb += 128) {                 // 128 bands at a time: four values per lane
      float v[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) v[k] = (jb + 32 * k + lane < B) ? x[jb + 32 * k] : 0.f;
      // slot c of output row t = r + half - c takes this row
      const int c_lo = (int)max(0LL, r + half - (c1 - 1)), c_hi = (int)min((long long)context - 1, r + half - c0);
      float *o = out + (r + half - c_lo) * W + (long long)c_lo * B + jb + lane;
      for (int c = c_lo; c <= c_hi; ++c, o += B - W) {
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (jb + 32 * k + lane < B) o[32 * k] = v[k];
      }
    }
    // slots of output row r whose source row r + c - half is outside the clip
    const int z_lo = (int)min((long long)context, max(0LL, c0 - (r - half)));          // slots [0, z_lo)
    const int z_hi = (int)max(0LL, min((long long)context, c1 + half - r));           // slots [z_hi, context)
    float *o = out + r * W;
    for (int i = lane; i < z_lo * B; i += 32) o[i] = 0.f;
    for (long long i = (long long)z_hi * B + lane; i < W; i += 32) o[i] = 0.f;
  }
}

// ---- stand-alone stages -------------------------------------------------------------------------
__global__ void k_magnitude(const float2 *__restrict__ in, long long n, float *__restrict__ out) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    float2 v = in[i];
    out[i] = fast_sqrt(fmaf(v.x, v.x, v.y * v.y));
  }
}

// one warp per row: y[j] = sum_i w[woff_j + i] * spec[row, start_j + i]; optional log10(mul*y+add)
__global__ void k_filter_log(const float *__restrict__ spec, long long ld_spec, long long rows, int num_bins,
                             int num_bands, const int *__restrict__ band_start, const int *__restrict__ band_len,
                             const int *__restrict__ band_woff, const float *__restrict__ weights,
                             int apply_filter, int apply_log, float mul, float add, int power, float log_scale,
                             float log_floor, float *__restrict__ out, long long ld_out) {
  const int lane = threadIdx.x & 31;
  const long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long r = warp; r < rows; r += nwarps) {
    const float *x = spec + r * ld_spec;
    if (apply_filter) {
      for (int j = 0; j < num_bands; ++j) {
        const int st = band_start[j], len = band_len[j], wo = band_woff[j];
        float acc = 0.f;
        for (int i = lane; i < len; i += 32) {
          float m = x[st + i];
          if (power) m *= m;
          acc = fmaf(weights[wo + i], m, acc);
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, d);
        if (lane == 0) {
          if (apply_log) {
            float a = __fadd_rn(__fmul_rn(mul, acc), add);
            if (log_floor > 0.f) a = fmaxf(a, log_floor);
            acc = log10f(a) * log_scale;
          }
          out[r * ld_out + j] = acc;
        }
      }
    } else {
      for (int j = lane; j < num_bins; j += 32) {
        float y = x[j];
        if (power) y *= y;
        if (apply_log) {
          float a = __fadd_rn(__fmul_rn(mul, y), add);
          if (log_floor > 0.f) a = fmaxf(a, log_floor);
          y = log10f(a) * log_scale;
        }
        out[r * ld_out + j] = y;
      }
    }
  }
}

// The fused kernels run without warm-up rows (FrontParams::seam_fix): the first kd rows of every task but a clip's
// first were differenced against a stale ring.  One warp per such row rewrites the difference (and the flux) from
// the (log-)filtered rows L, which are complete in global memory by now: D[r] = L[r] - L[r - kd].
// Tasks: task_off / chunk as k_setup_tasks left them; task i of clip c starts at frame (i - task_off[c]) * chunk.
__global__ void k_seam_diff(const int *__restrict__ task_off, int n_clips, int chunk, int chunk_small, int tail_clips,
                            const long long *__restrict__ frame_off,
                            const float *__restrict__ L, long long ld_L, int B, int kd, int positive,
                            float *__restrict__ out_diff, long long ld_out, float *__restrict__ flux) {
  const int lane = threadIdx.x & 31;
  const long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  const long long total = (long long)task_off[n_clips] * kd;   // the first kd rows of a task have their lagged row in another task
  for (long long idx = warp; idx < total; idx += nwarps) {
    const int task = (int)(idx / kd), t = (int)(idx - (long long)task * kd);
    int lo = 0, hi = n_clips;
    while (hi - lo > 1) {
      int mid = (lo + hi) >> 1;
      if (task_off[mid] <= task) lo = mid; else hi = mid;
    }
    const int ch = lo >= n_clips - tail_clips ? chunk_small : chunk;
    if (t >= ch) continue;                               // that row is the next task's
    const int f0 = (task - task_off[lo]) * ch;
    if (f0 == 0) continue;                               // a clip's first task: rows < kd are 0, written by the fused kernel
    const long long row0 = frame_off[lo];
    const int f = f0 + t;
    if (f >= (int)(frame_off[lo + 1] - row0)) continue;
    const long long r = row0 + f;
    const float *x = L + r * ld_L, *ref = L + (r - kd) * ld_L;
    float fsum = 0.f;
    for (int j = lane; j < B; j += 32) {
      float D = f >= kd ? x[j] - ref[j] : 0.f;
      if (positive) D = fmaxf(D, 0.f);
      if (out_diff != nullptr) out_diff[r * ld_out + j] = D;
      fsum += D;
    }
    if (flux != nullptr) {
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) fsum += __shfl_xor_sync(0xffffffffu, fsum, d);
      if (lane == 0) flux[r] = fsum;
    }
  }
}

// one warp per row: lagged (positive) difference inside each clip, flux row sum, projection
__global__ void k_diff_flux_proj(const float *__restrict__ L, long long ld_L, const long long *__restrict__ frame_off,
                                 int n_clips, long long rows, int B, int kd, int positive, int max_bins, int num_classes,
                                 const int *__restrict__ proj_off, const int *__restrict__ proj_band,
                                 const float *__restrict__ proj_w, float *__restrict__ out, long long ld_out,
                                 int col_spec, int col_diff, float *__restrict__ flux, float *__restrict__ proj,
                                 long long ld_proj) {
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  // a block owns a contiguous range of rows, its warps take them in turn: one clip lookup per warp instead of one
  // per row, and the lagged row is one a neighbouring warp has just pulled into L1
  const long long per_block = (rows + gridDim.x - 1) / gridDim.x;
  const long long r0 = blockIdx.x * per_block + wib, r1 = min(rows, (blockIdx.x + 1) * per_block);
  if (r0 >= r1) return;
  int lo = 0, hi = n_clips;
  while (hi - lo > 1) {
    int mid = (lo + hi) >> 1;
    if (frame_off[mid] <= r0) lo = mid; else hi = mid;
  }
  for (long long r = r0; r < r1; r += wpb) {
    while (lo + 1 < n_clips && r >= frame_off[lo + 1]) ++lo;
    const long long local = r - frame_off[lo];
    const float *x = L + r * ld_L;
    float fsum = 0.f;
    for (int j = lane; j < B; j += 32) {
      const float v = x[j];
      float D = 0.f;
      if (kd > 0) {
        if (local >= kd) {
          const float *ref = L + (r - kd) * ld_L;
          float m = ref[j];
          if (max_bins > 1) {   // scipy maximum_filter, size (1, M), origin 0, mode 'reflect': window [j - M/2, j + (M-1)/2]
            for (int i = j - max_bins / 2; i <= j + (max_bins - 1) / 2; ++i) {
              int q = i < 0 ? -i - 1 : i;
              q = q >= B ? 2 * B - 1 - q : q;
              q = min(max(q, 0), B - 1);
              m = fmaxf(m, ref[q]);
            }
          }
          D = v - m;
        }
        if (positive) D = fmaxf(D, 0.f);
      }
      if (out != nullptr) {
        if (col_spec >= 0) out[r * ld_out + col_spec + j] = v;
        if (col_diff >= 0) out[r * ld_out + col_diff + j] = D;
      }
      fsum += D;
    }
    if (flux != nullptr) {
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) fsum += __shfl_xor_sync(0xffffffffu, fsum, d);
      if (lane == 0) flux[r] = fsum;
    }
    if (proj != nullptr) {
      for (int c = lane; c < num_classes; c += 32) {
        float acc = 0.f;
        for (int i = proj_off[c]; i < proj_off[c + 1]; ++i) acc = fmaf(proj_w[i], x[proj_band[i]], acc);
        proj[r * ld_proj + c] = acc;
      }
    }
  }
}

}  // namespace b2
