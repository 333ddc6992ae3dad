// b200spec.cu -- C ABI of libb200spec.so (see include/b200spec.h): plans, tables, launches.
// No compute happens on the host; every entry point either fails loudly or launches sm_100a kernels.
#include "../../include/b200spec.h"

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include <nvtx3/nvToolsExt.h>

#include "aux_kernels.cuh"
#include "fb_pack.h"
#include "front_inst.cuh"

namespace {

thread_local std::string g_last_error;
std::atomic<long long> g_launches{0};

int fail(int code, const char *fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_last_error = buf;
  return code;
}

#define CU_CHECK(expr)                                                                      \
  do {                                                                                      \
    cudaError_t e__ = (expr);                                                               \
    if (e__ != cudaSuccess)                                                                 \
      return fail(B200SPEC_ERR_CUDA, "%s failed: %s", #expr, cudaGetErrorString(e__));      \
  } while (0)

struct DeviceGuard {
  int prev = -1;
  bool switched = false;
  cudaError_t enter(int dev) {
    cudaError_t e = cudaGetDevice(&prev);
    if (e != cudaSuccess) return e;
    if (prev != dev) {
      e = cudaSetDevice(dev);
      switched = (e == cudaSuccess);
    }
    return e;
  }
  ~DeviceGuard() {
    if (switched) cudaSetDevice(prev);
  }
};

// NVTX range around a launch sequence (header-only NVTX v3: costs one indirect call unless a tool is attached)
struct NvtxRange {
  explicit NvtxRange(const char *name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
};

// the device that owns a device pointer, and its SM count (entry points without a plan)
int pointer_device(const void *p, int *device, int *num_sms) {
  cudaPointerAttributes at;
  CU_CHECK(cudaPointerGetAttributes(&at, p));
  if (at.type != cudaMemoryTypeDevice && at.type != cudaMemoryTypeManaged)
    return fail(B200SPEC_ERR_ARG, "pointer %p is not device memory", p);
  *device = at.device;
  CU_CHECK(cudaDeviceGetAttribute(num_sms, cudaDevAttrMultiProcessorCount, at.device));
  return 0;
}

struct ResPlan {
  int frame_size = 0;
  double hop = 0;
  int origin = 0;
  int num_bands = 0, nnz = 0, kmax = 0;
  int log_enabled = 0;
  float mul = 1.f, add = 1.f;
  int power = 0;
  float log_scale = 1.f, log_floor = 0.f;
  int diff_frames = 0, positive = 0, diff_max_bins = 0, circular_shift = 0, include_nyquist = 0;
  int num_classes = 0, nproj = 0;
  // device tables
  float *d_window = nullptr;
  float2 *d_tw2 = nullptr, *d_tw3 = nullptr, *d_pt = nullptr, *d_wr = nullptr;
  float2 *d_pair_tw3 = nullptr, *d_pair_wr = nullptr;   // F-point tables of the pair transform (F <= 4096)
  int win_fly = 0;                                      // the window is a scaled np.hanning(F): evaluated in registers
  float win_h = 0.f;
  float2 win_cs[16] = {};
  float2 *d_win_ab = nullptr;
  // warp-per-pair kernel (frame 1024): W_1024^(b c) table, window constants, filterbank packed for 32 slabs per round
  float2 *d_w32_tw = nullptr;
  float2 win_cs32[32] = {};
  int w32_L = 0, w32_ns = 0, w32_kmin = 0, w32_ndw = 0;
  float4 *d_w32_w4 = nullptr;
  int4 *d_w32_band = nullptr;
  float *d_w32_dw = nullptr;
  int fb_L = 3, fb_ns = 1, fb_kmin = 0, fb_ndw = 0, fb_ndirect = 0, fb_w4_global = 0;
  float4 *d_fb_w4 = nullptr;
  int4 *d_fb_band = nullptr;
  float *d_fb_dw = nullptr;
  float *d_fbw = nullptr;
  int *d_band_start = nullptr, *d_band_len = nullptr, *d_band_woff = nullptr;
  int *d_proj_off = nullptr, *d_proj_band = nullptr;
  float *d_proj_w = nullptr;
};

}  // namespace

struct b200spec_plan {
  int device = 0;
  int dtype = 0;
  int channels = 1;
  int num_res = 0;
  int num_sms = 0;
  ResPlan res[B200SPEC_MAX_RES];
  std::vector<void *> allocs;
};

namespace {

template <class T>
int upload(b200spec_plan *pl, const T *host, size_t n, T **out) {
  *out = nullptr;
  // every table is padded with zeros to a multiple of 16 bytes: the kernels stage them with 16-byte-granular bulk
  // copies (bulk_stage.cuh); n == 0 still yields a valid pointer
  const size_t bytes = ((n == 0 ? 1 : n) * sizeof(T) + 15) & ~size_t(15);
  void *d = nullptr;
  CU_CHECK(cudaMalloc(&d, bytes));
  pl->allocs.push_back(d);
  CU_CHECK(cudaMemset(d, 0, bytes));
  if (host && n > 0) CU_CHECK(cudaMemcpy(d, host, n * sizeof(T), cudaMemcpyHostToDevice));   // nothing is read from a zero-length host array
  *out = reinterpret_cast<T *>(d);
  return 0;
}

int build_res(b200spec_plan *pl, const b200spec_res_desc &d, ResPlan &r) {
  const int F = d.frame_size;
  if (!(F == 1024 || F == 2048 || F == 4096 || F == 8192))
    return fail(B200SPEC_ERR_UNSUPPORTED, "frame_size %d not supported (1024, 2048, 4096, 8192)", F);
  if (!(d.hop_size > 0.0) || !std::isfinite(d.hop_size)) return fail(B200SPEC_ERR_ARG, "hop_size must be > 0");
  if (d.window == nullptr) return fail(B200SPEC_ERR_ARG, "window is NULL");
  if (d.num_bands < 0 || d.num_classes < 0 || d.num_classes > b2::kGroupThreads)
    return fail(B200SPEC_ERR_ARG, "num_bands/num_classes out of range");
  if (d.diff_frames < 0 || d.diff_frames > B200SPEC_MAX_DIFF_FRAMES)
    return fail(B200SPEC_ERR_UNSUPPORTED, "diff_frames %d outside [0, %d]", d.diff_frames, B200SPEC_MAX_DIFF_FRAMES);
  const int N = F / 2, R3 = N / 256;
  const double PI = 3.14159265358979323846;
  r.frame_size = F;
  r.hop = d.hop_size;
  r.origin = d.origin;
  r.log_enabled = d.log_enabled;
  r.mul = d.mul;
  r.add = d.add;
  r.power = d.power != 0;
  r.log_scale = d.log_scale == 0.f ? 1.f : d.log_scale;   // a zeroed descriptor means madmom's plain log10
  r.log_floor = d.log_floor;
  r.diff_frames = d.diff_frames;
  r.positive = d.positive_diffs;
  if (d.diff_max_bins < 0 || d.diff_max_bins > 64) return fail(B200SPEC_ERR_UNSUPPORTED, "diff_max_bins %d outside [0, 64]", d.diff_max_bins);
  r.diff_max_bins = d.diff_max_bins > 1 ? d.diff_max_bins : 0;
  r.circular_shift = d.circular_shift != 0;
  r.include_nyquist = d.include_nyquist != 0;
  if (r.include_nyquist && d.num_bands > 0)
    return fail(B200SPEC_ERR_UNSUPPORTED, "include_nyquist: a filterbank takes frame_size/2 bins, not frame_size/2 + 1");
  r.num_bands = d.num_bands;
  r.num_classes = d.num_classes;
  if (d.num_classes > 0) {   // validated before anything reads proj_off (the shared-memory probe below does)
    if (!d.proj_off || !d.proj_band || !d.proj_weight) return fail(B200SPEC_ERR_ARG, "projection arrays are NULL");
    if (d.proj_off[0] != 0) return fail(B200SPEC_ERR_ARG, "proj_off[0] must be 0");
    for (int c = 0; c < d.num_classes; ++c)
      if (d.proj_off[c + 1] < d.proj_off[c]) return fail(B200SPEC_ERR_ARG, "proj_off must be non-decreasing");
    if (d.proj_off[d.num_classes] > (1 << 20)) return fail(B200SPEC_ERR_ARG, "projection has too many entries");
  }

  // window: the real frame is packed as z[m] = x[2m] + i x[2m+1]; the 1/2 of the even/odd split
  // E = (Z[k] + conj(Z[N-k])) / 2 is folded into the window (exact in binary floating point)
  std::vector<float> win(F);
  for (int i = 0; i < F; ++i) win[i] = 0.5f * d.window[i];
  std::vector<float2> tw2(256), tw3(129 * R3), pt(129 * R3);
  for (int k1 = 0; k1 < 16; ++k1)
    for (int n2 = 0; n2 < 16; ++n2) {
      double a = -2.0 * PI * (double)(n2 * k1) / 256.0;
      tw2[k1 * 16 + n2] = make_float2((float)cos(a), (float)sin(a));
    }
  for (int q = 0; q <= 128; ++q)
    for (int n3 = 0; n3 < R3; ++n3) {
      double a = -2.0 * PI * (double)(n3 * q) / (double)N;
      tw3[n3 * 129 + q] = make_float2((float)cos(a), (float)sin(a));
    }
  for (int k3 = 0; k3 < R3; ++k3)
    for (int q = 0; q <= 128; ++q) {
      double a = -2.0 * PI * (double)(q + 256 * k3) / (double)F;
      pt[k3 * 129 + q] = make_float2((float)sin(a), (float)-cos(a));  // -i * exp(i a)
    }
  std::vector<float2> wr(2 * R3);
  for (int e = 0; e < 2 * R3; ++e) {
    double a = -2.0 * PI * (double)e / (double)(2 * R3);
    wr[e] = make_float2((float)cos(a), (float)sin(a));
  }
  int rc;
  if ((rc = upload(pl, wr.data(), wr.size(), &r.d_wr))) return rc;
  if ((rc = upload(pl, win.data(), win.size(), &r.d_window))) return rc;
  if ((rc = upload(pl, tw2.data(), tw2.size(), &r.d_tw2))) return rc;
  if ((rc = upload(pl, tw3.data(), tw3.size(), &r.d_tw3))) return rc;
  if ((rc = upload(pl, pt.data(), pt.size(), &r.d_pt))) return rc;
  if (F <= 4096) {   // pair transform (two frames as one complex FFT of F points): N' = F, R3' = F / 256
    const int R3p = F / 256;
    int rows = 0;
    while ((1 << rows) < R3p) ++rows;
    // compact table: row r = W_F^(2^r q) -- the only rows the last pass reads (tw3_rows forms the others as products)
    std::vector<float2> ptw3(129 * rows), pwr(2 * R3p);
    for (int q = 0; q <= 128; ++q)
      for (int r3 = 0; r3 < rows; ++r3) {
        double a = -2.0 * PI * (double)((1 << r3) * q) / (double)F;
        ptw3[r3 * 129 + q] = make_float2((float)cos(a), (float)sin(a));
      }
    for (int e = 0; e < 2 * R3p; ++e) {
      double a = -2.0 * PI * (double)e / (double)(2 * R3p);
      pwr[e] = make_float2((float)cos(a), (float)sin(a));
    }
    if ((rc = upload(pl, ptw3.data(), ptw3.size(), &r.d_pair_tw3))) return rc;
    if ((rc = upload(pl, pwr.data(), pwr.size(), &r.d_pair_wr))) return rc;
    // Is the window s * np.hanning(F) (madmom's default, / 32767 for int16)?  Then the pair kernel forms
    // 0.5 w[n] = H - H cos(2 pi n / (F-1)), H = s / 4, by angle addition instead of loading it.
    double sw = 0.0, shh = 0.0;
    std::vector<double> hann(F);
    for (int i = 0; i < F; ++i) {
      hann[i] = 0.5 - 0.5 * cos(2.0 * PI * (double)i / (double)(F - 1));
      sw += (double)d.window[i] * hann[i];
      shh += hann[i] * hann[i];
    }
    const double sc = sw / shh;
    double dev = 0.0;
    for (int i = 0; i < F; ++i) dev = std::max(dev, fabs((double)d.window[i] - sc * hann[i]));
#ifdef B200SPEC_TUNING   // A/B knob of tuning builds only (-DB200SPEC_TUNING); the product build reads no environment
    const char *wf = getenv("B200SPEC_WINFLY");
    const bool winfly_on = !(wf && wf[0] == '0');
#else
    const bool winfly_on = true;
#endif
    if (sc > 0.0 && dev <= 1.5e-7 * sc && winfly_on) {
      const int bpf = F / 16;
      const double H = 0.25 * sc;
      r.win_fly = 1;
      r.win_h = (float)H;
      for (int n1 = 0; n1 < 16; ++n1) {
        const double a = 2.0 * PI * (double)(n1 * bpf) / (double)(F - 1);
        r.win_cs[n1] = make_float2((float)cos(a), (float)sin(a));
      }
      for (int a32 = 0; a32 < 32; ++a32) {       // warp kernel: 32 points per lane, F / 32 samples apart
        const double a = 2.0 * PI * (double)(a32 * (F / 32)) / (double)(F - 1);
        r.win_cs32[a32] = make_float2((float)cos(a), (float)sin(a));
      }
      std::vector<float2> ab(bpf);
      for (int b = 0; b < bpf; ++b) {
        const double t = 2.0 * PI * (double)b / (double)(F - 1);
        ab[b] = make_float2((float)(-H * cos(t)), (float)(H * sin(t)));
      }
      if ((rc = upload(pl, ab.data(), ab.size(), &r.d_win_ab))) return rc;
    }
  }

  // banded filterbank -> interleaved slices of at most `seg_max` taps
  const int B = d.num_bands;
  int nnz = 0, kmax = 0;
  if (B > 0) {
    if (!d.band_start || !d.band_len || !d.band_woff || !d.weights)
      return fail(B200SPEC_ERR_ARG, "filterbank arrays are NULL");
    for (int j = 0; j < B; ++j) {
      if (d.band_len[j] < 0 || d.band_start[j] < 0 || d.band_start[j] + d.band_len[j] > N)
        return fail(B200SPEC_ERR_ARG, "band %d covers bins outside [0, %d)", j, N);
      if (d.band_woff[j] != nnz) return fail(B200SPEC_ERR_ARG, "band_woff must be the running sum of band_len");
      nnz += d.band_len[j];
      if (d.band_len[j] > 0 && d.band_start[j] + d.band_len[j] > kmax) kmax = d.band_start[j] + d.band_len[j];
    }
  }
  r.nnz = nnz;
  r.kmax = kmax;
  // slab form of the filterbank for the fused kernel (fb_pack.h)
  const int TBF = F == 1024 ? b2::FftCfg<1024>::TBF : F == 2048 ? b2::FftCfg<2048>::TBF : F == 4096 ? b2::FftCfg<4096>::TBF : b2::FftCfg<8192>::TBF;
  b2::FbPack fp = b2::fb_pack(N, B, d.band_start, d.band_len, d.band_woff, d.weights, TBF);
  {  // does the weight table fit next to everything else the kernel keeps in shared memory?
    b2::FrontParams probe{};
    probe.num_bands = B;
    probe.fb_L = fp.L;
    probe.fb_ns = fp.NS;
    probe.fb_ndw = (int)fp.dw.size();
    probe.diff_frames = d.diff_frames;
    probe.num_classes = d.num_classes;
    probe.mag_cap = (kmax + 16 + 3) & ~3;
    probe.nproj = d.num_classes > 0 ? d.proj_off[d.num_classes] : 0;
    size_t need = 0;
    switch (F) {
      case 1024: need = b2::front_smem_layout<1024>(probe, b2::MODE_LOGFILT, b2::GroupsPerCta<1024>::value); break;
      case 2048: need = b2::front_smem_layout<2048>(probe, b2::MODE_LOGFILT, b2::GroupsPerCta<2048>::value); break;
      case 4096: need = b2::front_smem_layout<4096>(probe, b2::MODE_LOGFILT, b2::GroupsPerCta<4096>::value); break;
      default: need = b2::front_smem_layout<8192>(probe, b2::MODE_LOGFILT, b2::GroupsPerCta<8192>::value); break;
    }
#ifdef B200SPEC_TUNING
    if (const char *e = getenv("B200SPEC_W4_GLOBAL"))   // tuning override: force the global-memory table
      if (e[0] == '1') need = b2::kMaxSmemPerCta + 1;
#endif
    if (need > b2::kMaxSmemPerCta) {   // keep the table in global memory (one fixed slab length)
      fp = b2::fb_pack(N, B, d.band_start, d.band_len, d.band_woff, d.weights, TBF, 15);
      r.fb_w4_global = 1;
    }
  }
  r.fb_L = fp.L;
  r.fb_ns = fp.NS;
  r.fb_kmin = fp.kmin;
  r.fb_ndw = (int)fp.dw.size();
  r.fb_ndirect = fp.ndirect;
  static_assert(sizeof(b2::FbBand) == sizeof(int4), "FbBand must match the int4 the kernel reads");
  if ((rc = upload(pl, d.weights, (size_t)nnz, &r.d_fbw))) return rc;
  if ((rc = upload(pl, reinterpret_cast<const float4 *>(fp.w4.data()), fp.w4.size() / 4, &r.d_fb_w4))) return rc;
  if ((rc = upload(pl, reinterpret_cast<const int4 *>(fp.band.data()), fp.band.size(), &r.d_fb_band))) return rc;
  if ((rc = upload(pl, fp.dw.data(), fp.dw.size(), &r.d_fb_dw))) return rc;
  if ((F == 1024 || F == 2048) && B > 0) {   // warp-per-FFT kernel: 32 slabs per round; two frames per call at frame 1024, one at 2048
    b2::FbPack f32p = b2::fb_pack(N, B, d.band_start, d.band_len, d.band_woff, d.weights, F == 1024 ? 2 : 1, 0, 32);
    std::vector<float2> tw(32 * 32);
    for (int cc = 0; cc < 32; ++cc)
      for (int b = 0; b < 32; ++b) {
        const double a = -2.0 * PI * (double)(b * cc) / 1024.0;
        tw[cc * 32 + b] = make_float2((float)cos(a), (float)sin(a));
      }
    r.w32_L = f32p.L;
    r.w32_ns = f32p.NS;
    r.w32_kmin = f32p.kmin;
    r.w32_ndw = (int)f32p.dw.size();
    if ((rc = upload(pl, tw.data(), tw.size(), &r.d_w32_tw))) return rc;
    if ((rc = upload(pl, reinterpret_cast<const float4 *>(f32p.w4.data()), f32p.w4.size() / 4, &r.d_w32_w4))) return rc;
    if ((rc = upload(pl, reinterpret_cast<const int4 *>(f32p.band.data()), f32p.band.size(), &r.d_w32_band))) return rc;
    if ((rc = upload(pl, f32p.dw.data(), f32p.dw.size(), &r.d_w32_dw))) return rc;
  }
  if ((rc = upload(pl, d.band_start, (size_t)B, &r.d_band_start))) return rc;
  if ((rc = upload(pl, d.band_len, (size_t)B, &r.d_band_len))) return rc;
  if ((rc = upload(pl, d.band_woff, (size_t)B, &r.d_band_woff))) return rc;

  if (d.num_classes > 0) {
    const int np = d.proj_off[d.num_classes];
    r.nproj = np;
    for (int i = 0; i < np; ++i)
      if (d.proj_band[i] < 0 || d.proj_band[i] >= B) return fail(B200SPEC_ERR_ARG, "proj_band out of range");
    if ((rc = upload(pl, d.proj_off, (size_t)d.num_classes + 1, &r.d_proj_off))) return rc;
    if ((rc = upload(pl, d.proj_band, (size_t)np, &r.d_proj_band))) return rc;
    if ((rc = upload(pl, d.proj_weight, (size_t)np, &r.d_proj_w))) return rc;
  }
  return 0;
}

// frames per task of the one-launch kernel: its point is that a chunk's samples are still in L2 when the second and third
// resolution read them, so the chunks of all resident groups (444 x chunk x 441 samples) must fit L2 with room to spare
#ifndef B2_MULTI_CHUNK_MAX
#define B2_MULTI_CHUNK_MAX 48
#endif
constexpr long long kMultiChunkMax = B2_MULTI_CHUNK_MAX;

struct Workspace {
  int *counter;
  int *task_off;
};

int carve_workspace(void *ws, size_t bytes, int n_clips, Workspace &w) {
  if (ws == nullptr || bytes < b200spec_workspace_bytes(n_clips))
    return fail(B200SPEC_ERR_ARG, "workspace too small: need %zu bytes", b200spec_workspace_bytes(n_clips));
  w.counter = reinterpret_cast<int *>(ws);
  w.task_off = reinterpret_cast<int *>(reinterpret_cast<char *>(ws) + 16);
  return 0;
}

// Frames per task.  Tasks are pulled from one counter by W workers per GPU (warps of the warp kernels, groups of the
// others), all of about the same length, so a launch runs ceil(tasks / W) rounds of (chunk + warm-up rows + a
// per-task overhead) frames: a chunk that leaves the last round nearly empty wastes up to a whole round (measured on
// B200, 64 x 18000 frames at frame 1024: 48-frame tasks = 10.15 rounds run 1.409 ms, 16- or 64-frame tasks 1.378 ms,
// 96-frame tasks = 5.08 rounds 1.486 ms).  Two plans are priced with that model and the cheaper one is taken:
//   one size   the candidate with the smallest rounds x (chunk + overhead); a small batch (less than one round) gets
//              the smallest tasks so that every SM has work;
//   two sizes  long tasks (little overhead) for all clips but the last few, whose short tasks fill the last round:
//              workers that find no long task left take short ones, and the launch ends within one short task
//              (needs enough clips that the last ones hold a round of long tasks' worth of frames).
// per_sm: workers resident per SM (16 warps of the warp kernel, 2-5 groups of the others); overhead: task fetch and
// the cold L1 of a task's first frames, in frames (0.75 for a warp, 1 for a group: fitted to the measured sweeps).
// kd: warm-up rows per task (0 when the seam fix-up forms the first differences, see launch_front).
struct TaskPlan {
  int chunk, chunk_small, tail_clips;
};

TaskPlan choose_tasks(int num_sms, int per_sm, double overhead, long long total_frames, int n_clips, int kd) {
  const double workers = (double)num_sms * per_sm;
  const double clips = n_clips > 0 ? (double)n_clips : 1.0;
  const double per_clip = (double)total_frames / clips;
  auto aligned = [kd](int c) {
    if (c >= 16) c -= (c + kd) % 4;            // a task transforms c + kd frames: whole tail batches (4 frames / pairs)
    else if ((c + kd) & 1) c += 1;             // whole pairs of frames
    return c;
  };
  static const int cand[] = {2, 4, 8, 12, 16, 20, 24, 28, 32, 40, 48, 56, 64, 72, 80, 96};
  TaskPlan best{16, 16, 0};
  double best_cost = 1e300;
  for (int c0 : cand) {
    const int c = aligned(c0);
    const double tasks = clips * std::ceil(per_clip / c);
    const double rounds = std::max(1.0, std::ceil(tasks / workers));
    const double cost = rounds * (c + kd + overhead);
    if (cost < best_cost * (1.0 - 1e-9) || (cost <= best_cost * (1.0 + 1e-9) && c > best.chunk)) {   // ties: fewer, longer tasks
      best_cost = cost;
      best = TaskPlan{c, c, 0};
    }
  }
#ifndef B2_ONE_TASK_SIZE   // tuning: -DB2_ONE_TASK_SIZE keeps the one-size plan
  const int cs = aligned(16);
  static const int cand_long[] = {48, 64, 80, 96};
  for (int c0 : cand_long) {
    const int cb = aligned(c0);
    const int tail = (int)std::ceil(workers * cb / std::max(per_clip, 1.0));     // clips that hold one round of long tasks
    if (n_clips < 8 || tail < 1 || tail > n_clips / 2) continue;
    const double long_tasks = (clips - tail) * std::ceil(per_clip / cb), short_tasks = tail * std::ceil(per_clip / cs);
    if (long_tasks < 2.0 * workers) continue;                                    // too little work for two phases
    const double cost = (long_tasks * (cb + kd + overhead) + short_tasks * (cs + kd + overhead)) / workers + (cs + kd + overhead);
    if (cost < best_cost * (1.0 - 1e-9)) {
      best_cost = cost;
      best = TaskPlan{cb, cs, tail};
    }
  }
#endif
#ifdef B200SPEC_TUNING
  if (const char *e = getenv("B200SPEC_CHUNK")) {   // tuning override: one size
    const int v = atoi(e);
    if (v > 0) best = TaskPlan{v, v, 0};
  }
#endif
  return best;
}

// everything of FrontParams that comes from the plan (tables, filterbank, constants)
void fill_plan_params(const b200spec_plan *pl, const ResPlan &r, b2::FrontParams &p) {
  p.frame_size = r.frame_size;
  p.hop = r.hop;
  p.origin = r.origin;
  p.window = r.d_window;
  p.win_fly = r.win_fly;
  p.win_h = r.win_h;
  for (int i = 0; i < 16; ++i) p.win_cs[i] = r.win_cs[i];
  for (int i = 0; i < 32; ++i) p.win_cs32[i] = r.win_cs32[i];
  p.win_ab = r.d_win_ab;
  p.tw2 = r.d_tw2;
  p.tw3 = r.d_tw3;
  p.pt = r.d_pt;
  p.wr = r.d_wr;
  p.fb_w4 = r.d_fb_w4;
  p.fb_band = r.d_fb_band;
  p.fb_dw = r.d_fb_dw;
  p.fb_L = r.fb_L;
  p.fb_ns = r.fb_ns;
  p.fb_kmin = r.fb_kmin;
  p.fb_ndw = r.fb_ndw;
  p.fb_w4_global = r.fb_w4_global;
  p.mag_cap = (r.kmax + 16 + 3) & ~3;   // k_front<8192> keeps only the magnitude bins the filterbank reads (+ one slab of slack)
  p.num_bands = r.num_bands;
  p.nnz = r.nnz;
  p.kmax = r.kmax;
  p.log_enabled = r.log_enabled;
  p.mul = r.mul;
  p.add = r.add;
  p.power = r.power;
  p.log_scale = r.log_scale;
  p.log_floor = r.log_floor;
  p.diff_frames = r.diff_frames;
  p.positive = r.positive;
  p.circular_shift = r.circular_shift;
  p.spec_ld = r.frame_size / 2 + (r.include_nyquist ? 1 : 0);
  p.proj_off = r.d_proj_off;
  p.proj_band = r.d_proj_band;
  p.proj_w = r.d_proj_w;
  (void)pl;
}

const char *range_name(int mode, int frame_size) {
  static const char *const kRangeNames[2][4] = {
      {"b200spec front end 1024", "b200spec front end 2048", "b200spec front end 4096", "b200spec front end 8192"},
      {"b200spec stft 1024", "b200spec stft 2048", "b200spec stft 4096", "b200spec stft 8192"}};
  return kRangeNames[mode == b2::MODE_LOGFILT ? 0 : 1][frame_size == 1024 ? 0 : frame_size == 2048 ? 1 : frame_size == 4096 ? 2 : 3];
}

int launch_front(const b200spec_plan *pl, int res, int mode, const void *d_sig, const int64_t *d_clip_off,
                 const int64_t *d_frame_off, int n_clips, int64_t total_frames, b2::FrontParams &p,
                 void *d_workspace, size_t workspace_bytes, void *stream) {
  if (!pl) return fail(B200SPEC_ERR_ARG, "plan is NULL");
  if (res < 0 || res >= pl->num_res) return fail(B200SPEC_ERR_ARG, "resolution %d out of range", res);
  if (n_clips < 0 || total_frames < 0) return fail(B200SPEC_ERR_ARG, "negative sizes");
  if (n_clips == 0 || total_frames == 0) return 0;
  if (!d_sig || !d_clip_off || !d_frame_off) return fail(B200SPEC_ERR_ARG, "NULL device pointer");
  const ResPlan &r = pl->res[res];
  Workspace w;
  int rc = carve_workspace(d_workspace, workspace_bytes, n_clips, w);
  if (rc) return rc;
  DeviceGuard guard;
  CU_CHECK(guard.enter(pl->device));
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  NvtxRange range(range_name(mode, r.frame_size));

  // The log-filtered path of frames <= 4096 runs the pair kernel (two frames per complex FFT); a configuration
  // whose tables do not fit next to the larger FFT buffer keeps the one-frame kernel (as does B200SPEC_PAIR=0 in a
  // tuning build: A/B measurements).
#ifdef B200SPEC_TUNING
  static const bool use_pair = []() { const char *v = getenv("B200SPEC_PAIR"); return !(v && v[0] == '0'); }();
#else
  constexpr bool use_pair = true;
#endif
#ifdef B2_NO_WARP      // tuning: keep the pair kernel at frames 1024 / 2048
  constexpr bool use_warp = false;
#else
  constexpr bool use_warp = true;
#endif
#ifndef B2_WARP_MAX_F  // tuning: -DB2_WARP_MAX_F=1024 keeps the pair kernel at frame 2048
#define B2_WARP_MAX_F 2048
#endif
  // Lagged difference without warm-up rows: a task transforms its own frames only, and the first diff_frames rows of
  // every task -- differenced against a ring that another task left behind -- are rewritten by k_seam_diff from
  // the filtered rows once they are all in global memory (1-2 transforms less per task, and tasks can be as small
  // as the load balance asks for: config 2 on B200 11.45 -> 11.33 ms at equal task sizes).  Needs the filtered rows in the
  // output matrix; a flux-only call keeps the warm-up rows.  No difference wanted: no warm-up either.
  const bool want_diff = mode == b2::MODE_LOGFILT && r.diff_frames > 0 && (p.col_diff >= 0 || p.flux != nullptr);
#ifdef B200SPEC_TUNING   // A/B knob of tuning builds: B200SPEC_SEAM=0 keeps the warm-up rows
  static const bool seam_on = []() { const char *v = getenv("B200SPEC_SEAM"); return !(v && v[0] == '0'); }();
#else
  constexpr bool seam_on = true;
#endif
  const bool seam_fix = seam_on && mode == b2::MODE_LOGFILT && r.diff_frames > 0 &&
                        (!want_diff || (p.out != nullptr && p.col_spec >= 0));
  const bool warp_path = use_warp && use_pair && mode == b2::MODE_LOGFILT && (r.frame_size == 1024 || r.frame_size == B2_WARP_MAX_F) &&
                         r.d_w32_tw != nullptr && p.proj == nullptr && (r.frame_size == 1024 || r.w32_ns * r.w32_L <= 30);
  const int per_sm = warp_path ? 16 : r.frame_size == 1024 ? 5 : r.frame_size == 2048 ? 4 : 3;
  const double task_overhead = warp_path ? 0.75 : 1.0;   // fitted to the task-size sweeps in profiles/r02_variants.txt
  const TaskPlan tp = choose_tasks(pl->num_sms, per_sm, task_overhead, total_frames, n_clips,
                                   (mode == b2::MODE_LOGFILT && !seam_fix) ? r.diff_frames : 0);
  const int chunk = tp.chunk;
  p.seam_fix = seam_fix ? 1 : 0;
  p.chunk_small = tp.chunk_small;
  p.tail_clips = tp.tail_clips;
  b2::k_setup_tasks<<<1, 1024, 0, st>>>(reinterpret_cast<const long long *>(d_frame_off), n_clips, chunk, tp.chunk_small,
                                         tp.tail_clips, w.task_off, w.counter);
  CU_CHECK(cudaGetLastError());
  g_launches++;

  p.sig = d_sig;
  p.clip_off = reinterpret_cast<const long long *>(d_clip_off);
  p.frame_off = reinterpret_cast<const long long *>(d_frame_off);
  p.n_clips = n_clips;
  p.task_off = w.task_off;
  p.task_counter = w.counter;
  p.chunk = chunk;
  fill_plan_params(pl, r, p);

  const int in = (pl->dtype == B200SPEC_I16 ? 2 : 0) + (pl->channels == 2 ? 1 : 0);
  const long long task_bound = total_frames / std::min(chunk, tp.chunk_small) + n_clips;
  cudaError_t e;
  auto seam_fixup = [&]() -> int {
    if (!(seam_fix && want_diff)) return 0;
    long long warps = task_bound * r.diff_frames;
    long long blocks = (warps + 7) / 8;
    if (blocks > pl->num_sms * 8) blocks = pl->num_sms * 8;
    b2::k_seam_diff<<<(int)(blocks < 1 ? 1 : blocks), 256, 0, st>>>(
        w.task_off, n_clips, chunk, tp.chunk_small, tp.tail_clips, p.frame_off, p.out + p.col_spec, p.ld_out, r.num_bands, r.diff_frames, r.positive,
        p.col_diff >= 0 ? p.out + p.col_diff : nullptr, p.ld_out, p.flux);
    CU_CHECK(cudaGetLastError());
    g_launches++;
    return 0;
  };
  // frames 1024 / 2048: one warp per FFT (frontend_warp_kernel.cuh); the projection output keeps the pair kernel
  // (frame 2048 runs one frame per FFT there and walks the filterbank once per frame: with a filterbank that covers the
  //  whole spectrum -- librosa's 128 mel bands: 45 bins per lane -- the pair kernel, which amortises the weights over
  //  four frames, is the faster one: 15.7 against 18.6 ms on the onset-strength workload)
  if (warp_path) {
    b2::FrontParams q = p;
    q.tw3 = r.d_w32_tw;
    q.fb_w4 = r.d_w32_w4;
    q.fb_band = r.d_w32_band;
    q.fb_dw = r.d_w32_dw;
    q.fb_L = r.w32_L;
    q.fb_ns = r.w32_ns;
    q.fb_kmin = r.w32_kmin;
    q.fb_ndw = r.w32_ndw;
    e = r.frame_size == 1024 ? b2_launch_warp_1024(in, q, pl->num_sms, task_bound, st)
                             : b2_launch_warp_2048(in, q, pl->num_sms, task_bound, st);
    if (e == cudaSuccess) {
      g_launches++;
      return seam_fixup();
    }
    if (e != cudaErrorInvalidConfiguration)
      return fail(B200SPEC_ERR_CUDA, "front-end (warp) kernel launch failed: %s", cudaGetErrorString(e));
    (void)cudaGetLastError();   // did not fit: fall through to the pair kernel
  }
  if (use_pair && mode == b2::MODE_LOGFILT && r.frame_size <= 4096 && r.d_pair_tw3 != nullptr) {
    b2::FrontParams q = p;
    q.tw3 = r.d_pair_tw3;
    q.wr = r.d_pair_wr;
    switch (r.frame_size) {
      case 1024: e = b2_launch_pair_1024(in, q, pl->num_sms, task_bound, st); break;
      case 2048: e = b2_launch_pair_2048(in, q, pl->num_sms, task_bound, st); break;
      default: e = b2_launch_pair_4096(in, q, pl->num_sms, task_bound, st); break;
    }
    if (e == cudaSuccess) {
      g_launches++;
      return seam_fixup();
    }
    if (e != cudaErrorInvalidConfiguration)
      return fail(B200SPEC_ERR_CUDA, "front-end (pair) kernel launch failed: %s", cudaGetErrorString(e));
    (void)cudaGetLastError();   // did not fit: fall through to the one-frame kernel
  }
  switch (r.frame_size) {
    case 1024: e = b2_launch_front_1024(in, mode, p, pl->num_sms, task_bound, st); break;
    case 2048: e = b2_launch_front_2048(in, mode, p, pl->num_sms, task_bound, st); break;
    case 4096: e = b2_launch_front_4096(in, mode, p, pl->num_sms, task_bound, st); break;
    default: e = b2_launch_front_8192(in, mode, p, pl->num_sms, task_bound, st); break;
  }
  if (e == cudaErrorInvalidConfiguration)
    return fail(B200SPEC_ERR_UNSUPPORTED, "the fused kernel for frame size %d with this filterbank (%d bands) needs more "
                "shared memory than an SM has; use the unfused stft / filter_log calls", r.frame_size, r.num_bands);
  if (e != cudaSuccess) return fail(B200SPEC_ERR_CUDA, "front-end kernel launch failed: %s", cudaGetErrorString(e));
  g_launches++;
  return seam_fixup();
}

void fill_out_params(const b200spec_out_desc &out, const ResPlan &r, b2::FrontParams &p) {
  p.out = out.d_out;
  p.ld_out = out.ld_out;
  p.col_spec = out.col_spec;
  p.col_diff = out.col_diff;
  p.flux = out.d_flux;
  p.clip_scale = out.d_clip_scale;
  p.clip_status = out.d_clip_status;
  p.proj = out.d_proj;
  p.ld_proj = out.ld_proj;
  p.num_classes = out.d_proj ? r.num_classes : 0;
  p.nproj = out.d_proj ? r.nproj : 0;
}

int check_out_desc(const b200spec_out_desc &out, const ResPlan &r, int res) {
  if (r.num_bands <= 0) return fail(B200SPEC_ERR_ARG, "resolution %d has no filterbank; use b200spec_spectrogram", res);
  if (out.d_proj && r.num_classes <= 0) return fail(B200SPEC_ERR_ARG, "d_proj given but the plan has no projection");
  if (out.d_out && out.ld_out < r.num_bands) return fail(B200SPEC_ERR_ARG, "ld_out smaller than num_bands");
  if (out.col_diff >= 0 && r.diff_frames <= 0) return fail(B200SPEC_ERR_ARG, "col_diff given but diff_frames == 0");
  if (r.diff_max_bins > 1 && (out.col_diff >= 0 || out.d_flux))
    return fail(B200SPEC_ERR_UNSUPPORTED, "diff_max_bins > 1: the fused kernel writes the log-filtered rows only; "
                "compute the SuperFlux difference with b200spec_diff_flux_chroma on them");
  return 0;
}

}  // namespace

extern "C" {

int b200spec_abi_version(void) { return B200SPEC_ABI_VERSION; }

const char *b200spec_last_error(void) { return g_last_error.c_str(); }

int b200spec_num_frames(int64_t n_samples, double hop_size, int end_mode, int64_t *out) {
  if (!out) return fail(B200SPEC_ERR_ARG, "out is NULL");
  if (n_samples < 0 || !(hop_size > 0.0)) return fail(B200SPEC_ERR_ARG, "n_samples < 0 or hop_size <= 0");
  const double q = (double)n_samples / hop_size;  // same float64 division numpy performs
  if (end_mode == B200SPEC_END_NORMAL) *out = (int64_t)std::ceil(q);
  else if (end_mode == B200SPEC_END_EXTEND) *out = (int64_t)std::floor(q + 1.0);
  else return fail(B200SPEC_ERR_ARG, "end of signal handling '%d' unknown", end_mode);
  return 0;
}

int b200spec_frame_start(int64_t index, double hop_size, int32_t frame_size, int32_t origin, int64_t *out) {
  if (!out) return fail(B200SPEC_ERR_ARG, "out is NULL");
  *out = (int64_t)((double)index * hop_size) - frame_size / 2 - origin;  // int() truncates like Python
  return 0;
}

int b200spec_plan_create(const b200spec_plan_desc *desc, b200spec_plan **out) {
  if (!desc || !out) return fail(B200SPEC_ERR_ARG, "desc/out is NULL");
  *out = nullptr;
  if (desc->num_res < 1 || desc->num_res > B200SPEC_MAX_RES)
    return fail(B200SPEC_ERR_ARG, "num_res %d outside [1, %d]", desc->num_res, B200SPEC_MAX_RES);
  if (desc->dtype != B200SPEC_F32 && desc->dtype != B200SPEC_I16) return fail(B200SPEC_ERR_ARG, "unknown dtype");
  if (desc->channels != 1 && desc->channels != 2)
    return fail(B200SPEC_ERR_UNSUPPORTED, "channels must be 1 or 2 (got %d)", desc->channels);
  int ndev = 0;
  CU_CHECK(cudaGetDeviceCount(&ndev));
  if (desc->device < 0 || desc->device >= ndev)
    return fail(B200SPEC_ERR_ARG, "device %d not present (%d devices)", desc->device, ndev);
  cudaDeviceProp prop;
  CU_CHECK(cudaGetDeviceProperties(&prop, desc->device));
  if (prop.major != 10)
    return fail(B200SPEC_ERR_ARCH, "device %d is sm_%d%d; this library only carries sm_100a code", desc->device,
                prop.major, prop.minor);
  DeviceGuard guard;
  CU_CHECK(guard.enter(desc->device));
  b200spec_plan *pl = new b200spec_plan();
  pl->device = desc->device;
  pl->dtype = desc->dtype;
  pl->channels = desc->channels;
  pl->num_res = desc->num_res;
  pl->num_sms = prop.multiProcessorCount;
  for (int i = 0; i < desc->num_res; ++i) {
    int rc = build_res(pl, desc->res[i], pl->res[i]);
    if (rc) {
      b200spec_plan_destroy(pl);
      return rc;
    }
  }
  *out = pl;
  return 0;
}

int b200spec_plan_destroy(b200spec_plan *plan) {
  if (!plan) return 0;
  DeviceGuard guard;
  guard.enter(plan->device);
  for (void *d : plan->allocs) cudaFree(d);
  delete plan;
  return 0;
}

size_t b200spec_workspace_bytes(int32_t n_clips) {
  if (n_clips < 0) n_clips = 0;
  return 16 + sizeof(int) * ((size_t)n_clips + 1) + 16;
}

int b200spec_stft(const b200spec_plan *plan, int32_t res, const void *d_sig, const int64_t *d_clip_off,
                  const int64_t *d_frame_off, int32_t n_clips, int64_t total_frames, float *d_out,
                  void *d_workspace, size_t workspace_bytes, void *stream) {
  if (!d_out && total_frames > 0) return fail(B200SPEC_ERR_ARG, "d_out is NULL");
  b2::FrontParams p{};
  p.spec_out = d_out;
  p.spec_complex = 1;
  return launch_front(plan, res, b2::MODE_SPECTRUM, d_sig, d_clip_off, d_frame_off, n_clips, total_frames, p,
                      d_workspace, workspace_bytes, stream);
}

int b200spec_spectrogram(const b200spec_plan *plan, int32_t res, const void *d_sig, const int64_t *d_clip_off,
                         const int64_t *d_frame_off, int32_t n_clips, int64_t total_frames, float *d_out,
                         void *d_workspace, size_t workspace_bytes, void *stream) {
  if (!d_out && total_frames > 0) return fail(B200SPEC_ERR_ARG, "d_out is NULL");
  b2::FrontParams p{};
  p.spec_out = d_out;
  p.spec_complex = 0;
  return launch_front(plan, res, b2::MODE_SPECTRUM, d_sig, d_clip_off, d_frame_off, n_clips, total_frames, p,
                      d_workspace, workspace_bytes, stream);
}

int b200spec_logfilt(const b200spec_plan *plan, int32_t res, const void *d_sig, const int64_t *d_clip_off,
                     const int64_t *d_frame_off, int32_t n_clips, int64_t total_frames,
                     const b200spec_out_desc *out, void *d_workspace, size_t workspace_bytes, void *stream) {
  if (!plan) return fail(B200SPEC_ERR_ARG, "plan is NULL");
  if (!out) return fail(B200SPEC_ERR_ARG, "out descriptor is NULL");
  if (res < 0 || res >= plan->num_res) return fail(B200SPEC_ERR_ARG, "resolution %d out of range", res);
  const ResPlan &r = plan->res[res];
  if (int rc = check_out_desc(*out, r, res)) return rc;
  b2::FrontParams p{};
  fill_out_params(*out, r, p);
  return launch_front(plan, res, b2::MODE_LOGFILT, d_sig, d_clip_off, d_frame_off, n_clips, total_frames, p,
                      d_workspace, workspace_bytes, stream);
}

int b200spec_logfilt_multi_supported(const b200spec_plan *plan, int32_t n_res, const int32_t *res) {
  if (!plan || !res || n_res < 1 || n_res > b2::kMultiMaxRes) return 0;
  b2::MultiParams m{};
  m.n_res = n_res;
  for (int i = 0; i < n_res; ++i) {
    if (res[i] < 0 || res[i] >= plan->num_res) return 0;
    const ResPlan &r = plan->res[res[i]];
    if (r.frame_size > 4096 || r.num_bands <= 0 || r.d_pair_tw3 == nullptr || r.fb_w4_global) return 0;
    if (r.hop != plan->res[res[0]].hop || r.origin != plan->res[res[0]].origin) return 0;
    fill_plan_params(plan, r, m.r[i]);
    m.r[i].num_classes = r.num_classes;     // worst case for the shared-memory plan: projection tables staged
    m.r[i].nproj = r.nproj;
  }
  return b2::multi_smem_layout(m, b2::kMultiGroups) <= b2::kMaxSmemPerCta ? 1 : 0;
}

int b200spec_logfilt_multi(const b200spec_plan *plan, int32_t n_res, const int32_t *res, const void *d_sig,
                           const int64_t *d_clip_off, const int64_t *d_frame_off, int32_t n_clips,
                           int64_t total_frames, const b200spec_out_desc *outs, void *d_workspace,
                           size_t workspace_bytes, void *stream) {
  if (!plan) return fail(B200SPEC_ERR_ARG, "plan is NULL");
  if (!res || !outs) return fail(B200SPEC_ERR_ARG, "res / outs is NULL");
  if (n_res < 1 || n_res > b2::kMultiMaxRes) return fail(B200SPEC_ERR_ARG, "n_res %d outside [1, %d]", n_res, b2::kMultiMaxRes);
  if (n_clips < 0 || total_frames < 0) return fail(B200SPEC_ERR_ARG, "negative sizes");
  b2::MultiParams m{};
  m.n_res = n_res;
  int kd_max = 0;
  for (int i = 0; i < n_res; ++i) {
    if (res[i] < 0 || res[i] >= plan->num_res) return fail(B200SPEC_ERR_ARG, "resolution %d out of range", res[i]);
    const ResPlan &r = plan->res[res[i]];
    if (int rc = check_out_desc(outs[i], r, res[i])) return rc;
    if (r.frame_size > 4096 || r.d_pair_tw3 == nullptr)
      return fail(B200SPEC_ERR_UNSUPPORTED, "the one-launch kernel runs frame sizes 1024 / 2048 / 4096 (got %d)", r.frame_size);
    if (r.fb_w4_global) return fail(B200SPEC_ERR_UNSUPPORTED, "resolution %d: filterbank table too large for the one-launch kernel", res[i]);
    if (r.hop != plan->res[res[0]].hop || r.origin != plan->res[res[0]].origin)
      return fail(B200SPEC_ERR_ARG, "all resolutions of one launch must share hop_size and origin (rows are per frame)");
    if (r.diff_frames > kd_max) kd_max = r.diff_frames;
  }
  if (n_clips == 0 || total_frames == 0) return 0;
  if (!d_sig || !d_clip_off || !d_frame_off) return fail(B200SPEC_ERR_ARG, "NULL device pointer");
  Workspace w;
  if (int rc = carve_workspace(d_workspace, workspace_bytes, n_clips, w)) return rc;
  DeviceGuard guard;
  CU_CHECK(guard.enter(plan->device));
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  NvtxRange range("b200spec front end (all resolutions, one launch)");

  // one task list for all resolutions: (clip, chunk of frames); every resolution of a task runs on the same group
  const long long slots = (long long)plan->num_sms * b2::kMultiGroups;
  long long chunk = total_frames / (slots * 8);
  if (chunk < 16) chunk = 16;
  if (chunk > kMultiChunkMax) chunk = kMultiChunkMax;
  if (chunk >= 16) chunk -= (chunk + kd_max) % 4;
  b2::k_setup_tasks<<<1, 1024, 0, st>>>(reinterpret_cast<const long long *>(d_frame_off), n_clips, (int)chunk, (int)chunk, 0,
                                         w.task_off, w.counter);
  CU_CHECK(cudaGetLastError());
  g_launches++;
  for (int i = 0; i < n_res; ++i) {
    const ResPlan &r = plan->res[res[i]];
    b2::FrontParams &p = m.r[i];
    fill_out_params(outs[i], r, p);
    p.sig = d_sig;
    p.clip_off = reinterpret_cast<const long long *>(d_clip_off);
    p.frame_off = reinterpret_cast<const long long *>(d_frame_off);
    p.n_clips = n_clips;
    p.task_off = w.task_off;
    p.task_counter = w.counter;
    p.chunk = (int)chunk;
    fill_plan_params(plan, r, p);
    p.tw3 = r.d_pair_tw3;
    p.wr = r.d_pair_wr;
  }
  const int in = (plan->dtype == B200SPEC_I16 ? 2 : 0) + (plan->channels == 2 ? 1 : 0);
  const long long task_bound = total_frames / chunk + n_clips;
  const cudaError_t e = b2_launch_multi(in, m, plan->num_sms, task_bound, st);
  if (e == cudaErrorInvalidConfiguration)
    return fail(B200SPEC_ERR_UNSUPPORTED, "these resolutions do not fit one launch (shared memory); use b200spec_logfilt per resolution");
  if (e != cudaSuccess) return fail(B200SPEC_ERR_CUDA, "front-end (one-launch) kernel launch failed: %s", cudaGetErrorString(e));
  g_launches++;
  return 0;
}

int b200spec_clip_peak(const b200spec_plan *plan, const void *d_sig, const int64_t *d_clip_off, int32_t n_clips,
                       float eps, int32_t reciprocal, float *d_peak, void *stream) {
  if (!plan) return fail(B200SPEC_ERR_ARG, "plan is NULL");
  if (n_clips < 0) return fail(B200SPEC_ERR_ARG, "n_clips < 0");
  if (n_clips == 0) return 0;
  if (!d_sig || !d_clip_off || !d_peak) return fail(B200SPEC_ERR_ARG, "NULL device pointer");
  DeviceGuard guard;
  CU_CHECK(guard.enter(plan->device));
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  CU_CHECK(cudaMemsetAsync(d_peak, 0, sizeof(float) * (size_t)n_clips, st));
  if ((long long)n_clips * b2::kPeakBlocksPerClip > 0x7fffffffLL) return fail(B200SPEC_ERR_ARG, "too many clips");
  const unsigned grid = (unsigned)n_clips * b2::kPeakBlocksPerClip;
  const long long *co = reinterpret_cast<const long long *>(d_clip_off);
  unsigned int *bits = reinterpret_cast<unsigned int *>(d_peak);
  switch ((plan->dtype == B200SPEC_I16 ? 2 : 0) + (plan->channels == 2 ? 1 : 0)) {
    case b2::IN_F32_MONO: b2::k_clip_peak<b2::IN_F32_MONO><<<grid, 256, 0, st>>>(d_sig, co, bits); break;
    case b2::IN_F32_STEREO: b2::k_clip_peak<b2::IN_F32_STEREO><<<grid, 256, 0, st>>>(d_sig, co, bits); break;
    case b2::IN_I16_MONO: b2::k_clip_peak<b2::IN_I16_MONO><<<grid, 256, 0, st>>>(d_sig, co, bits); break;
    default: b2::k_clip_peak<b2::IN_I16_STEREO><<<grid, 256, 0, st>>>(d_sig, co, bits); break;
  }
  CU_CHECK(cudaGetLastError());
  g_launches++;
  if (reciprocal) {
    // the int16 window is pre-divided by 32767 (madmom stft.py); the gain of a peak-normalised clip undoes it
    const float numer = plan->dtype == B200SPEC_I16 ? 32767.f : 1.f;
    b2::k_peak_reciprocal<<<(n_clips + 255) / 256, 256, 0, st>>>(d_peak, n_clips, eps, numer);
    CU_CHECK(cudaGetLastError());
    g_launches++;
  }
  return 0;
}

int b200spec_context_stack(const float *d_in, int64_t ld_in, int32_t num_bands, const int64_t *d_frame_off,
                           int32_t n_clips, int64_t total_frames, int32_t context, float *d_out, void *stream) {
  if (n_clips < 0 || total_frames < 0) return fail(B200SPEC_ERR_ARG, "negative sizes");
  if (n_clips == 0 || total_frames == 0) return 0;
  if (!d_in || !d_frame_off || !d_out) return fail(B200SPEC_ERR_ARG, "NULL device pointer");
  if (num_bands < 1 || ld_in < num_bands || context < 1) return fail(B200SPEC_ERR_ARG, "num_bands / ld_in / context out of range");
  int dev = 0, sms = 0, rc = pointer_device(d_in, &dev, &sms);   // no plan here: launch on the device that owns d_in
  if (rc) return rc;
  DeviceGuard guard;
  CU_CHECK(guard.enter(dev));
  NvtxRange range("b200spec context stack");
  long long blocks = (total_frames + 7) / 8;
  if (blocks > sms * 16) blocks = sms * 16;
  b2::k_context_stack<<<(int)blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      d_in, ld_in, num_bands, reinterpret_cast<const long long *>(d_frame_off), n_clips, total_frames, context, d_out);
  CU_CHECK(cudaGetLastError());
  g_launches++;
  return 0;
}

int b200spec_onset_envelope(const float *d_L, int64_t ld_L, int32_t num_bands, const int64_t *d_frame_off,
                            int32_t n_clips, int64_t total_frames, int32_t lag, float top_db, int32_t aggregate,
                            int32_t shift, float *d_clip_max, float *d_env, void *stream) {
  if (n_clips < 0 || total_frames < 0) return fail(B200SPEC_ERR_ARG, "negative sizes");
  if (n_clips == 0 || total_frames == 0) return 0;
  if (!d_L || !d_frame_off || !d_clip_max || !d_env) return fail(B200SPEC_ERR_ARG, "NULL device pointer");
  if (num_bands < 1 || num_bands > 1024 || ld_L < num_bands) return fail(B200SPEC_ERR_ARG, "num_bands outside [1, 1024] or ld_L < num_bands");
  if (lag < 1 || shift < 0) return fail(B200SPEC_ERR_ARG, "lag must be >= 1 and shift >= 0");
  if (aggregate != 0 && aggregate != 1) return fail(B200SPEC_ERR_ARG, "aggregate must be 0 (mean) or 1 (median)");
  if ((long long)n_clips * b2::kRowmaxBlocksPerClip > 0x7fffffffLL) return fail(B200SPEC_ERR_ARG, "too many clips");
  int dev = 0, sms = 0, rc = pointer_device(d_L, &dev, &sms);    // no plan here: launch on the device that owns d_L
  if (rc) return rc;
  DeviceGuard guard;
  CU_CHECK(guard.enter(dev));
  NvtxRange range("b200spec onset envelope");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const long long *fo = reinterpret_cast<const long long *>(d_frame_off);
  b2::k_fill_neg_inf<<<(n_clips + 255) / 256, 256, 0, st>>>(d_clip_max, n_clips);
  b2::k_clip_rowmax<<<(unsigned)n_clips * b2::kRowmaxBlocksPerClip, 256, 0, st>>>(d_L, ld_L, num_bands, fo, d_clip_max);
  CU_CHECK(cudaGetLastError());
  g_launches += 2;
  long long blocks = (total_frames + 7) / 8;
  if (blocks > sms * 16) blocks = sms * 16;
  // rows of up to 256 bands are sorted in registers (1, 2, 4 or 8 values per lane); wider ones go through shared memory
  const int per_lane = (num_bands + 31) / 32;
#define B2_ONSET_ARGS d_L, ld_L, num_bands, fo, n_clips, total_frames, lag, top_db, aggregate, shift, d_clip_max, d_env
  if (per_lane <= 1) b2::k_onset_env<1><<<(int)blocks, 256, 0, st>>>(B2_ONSET_ARGS);
  else if (per_lane <= 2) b2::k_onset_env<2><<<(int)blocks, 256, 0, st>>>(B2_ONSET_ARGS);
  else if (per_lane <= 4) b2::k_onset_env<4><<<(int)blocks, 256, 0, st>>>(B2_ONSET_ARGS);
  else if (per_lane <= 8) b2::k_onset_env<8><<<(int)blocks, 256, 0, st>>>(B2_ONSET_ARGS);
  else b2::k_onset_env<0><<<(int)blocks, 256, sizeof(float) * 8 * (size_t)num_bands, st>>>(B2_ONSET_ARGS);
#undef B2_ONSET_ARGS
  CU_CHECK(cudaGetLastError());
  g_launches++;
  return 0;
}

int b200spec_magnitude(const float *d_stft, int64_t n_elems, float *d_out, void *stream) {
  if (n_elems < 0) return fail(B200SPEC_ERR_ARG, "n_elems < 0");
  if (n_elems == 0) return 0;
  if (!d_stft || !d_out) return fail(B200SPEC_ERR_ARG, "NULL device pointer");
  int dev = 0, sms = 0, rc = pointer_device(d_stft, &dev, &sms);  // no plan here: launch on the device that owns d_stft
  if (rc) return rc;
  DeviceGuard guard;
  CU_CHECK(guard.enter(dev));
  long long blocks = (n_elems + 255) / 256;
  if (blocks > sms * 16) blocks = sms * 16;
  b2::k_magnitude<<<(int)blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const float2 *>(d_stft), n_elems, d_out);
  CU_CHECK(cudaGetLastError());
  g_launches++;
  return 0;
}

int b200spec_filter_log(const b200spec_plan *plan, int32_t res, const float *d_spec, int64_t ld_spec,
                        int64_t total_frames, int32_t apply_filter, int32_t apply_log, float *d_out,
                        int64_t ld_out, void *stream) {
  if (!plan) return fail(B200SPEC_ERR_ARG, "plan is NULL");
  if (res < 0 || res >= plan->num_res) return fail(B200SPEC_ERR_ARG, "resolution %d out of range", res);
  if (total_frames < 0) return fail(B200SPEC_ERR_ARG, "total_frames < 0");
  if (total_frames == 0) return 0;
  if (!d_spec || !d_out) return fail(B200SPEC_ERR_ARG, "NULL device pointer");
  const ResPlan &r = plan->res[res];
  if (apply_filter && r.num_bands <= 0) return fail(B200SPEC_ERR_ARG, "resolution has no filterbank");
  DeviceGuard guard;
  CU_CHECK(guard.enter(plan->device));
  long long blocks = (total_frames + 7) / 8;
  if (blocks > plan->num_sms * 8) blocks = plan->num_sms * 8;
  const int num_bins = apply_filter ? r.frame_size / 2 : (int)ld_spec;
  b2::k_filter_log<<<(int)blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      d_spec, ld_spec, total_frames, num_bins, r.num_bands, r.d_band_start, r.d_band_len, r.d_band_woff, r.d_fbw,
      apply_filter, apply_log, r.mul, r.add, r.power, r.log_scale, r.log_floor, d_out, ld_out);
  CU_CHECK(cudaGetLastError());
  g_launches++;
  return 0;
}

int b200spec_diff_flux_chroma(const b200spec_plan *plan, int32_t res, const float *d_L, int64_t ld_L,
                              const int64_t *d_frame_off, int32_t n_clips, int64_t total_frames,
                              const b200spec_out_desc *out, void *stream) {
  if (!plan || !out) return fail(B200SPEC_ERR_ARG, "plan/out is NULL");
  if (res < 0 || res >= plan->num_res) return fail(B200SPEC_ERR_ARG, "resolution %d out of range", res);
  if (total_frames < 0 || n_clips < 0) return fail(B200SPEC_ERR_ARG, "negative sizes");
  if (total_frames == 0 || n_clips == 0) return 0;
  if (!d_L || !d_frame_off) return fail(B200SPEC_ERR_ARG, "NULL device pointer");
  const ResPlan &r = plan->res[res];
  if (out->d_proj && r.num_classes <= 0) return fail(B200SPEC_ERR_ARG, "d_proj given but the plan has no projection");
  const int B = r.num_bands > 0 ? r.num_bands : (int)ld_L;
  DeviceGuard guard;
  CU_CHECK(guard.enter(plan->device));
  long long blocks = (total_frames + 7) / 8;
  if (blocks > plan->num_sms * 8) blocks = plan->num_sms * 8;
  b2::k_diff_flux_proj<<<(int)blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      d_L, ld_L, reinterpret_cast<const long long *>(d_frame_off), n_clips, total_frames, B, r.diff_frames,
      r.positive, r.diff_max_bins, out->d_proj ? r.num_classes : 0, r.d_proj_off, r.d_proj_band, r.d_proj_w, out->d_out,
      out->ld_out, out->col_spec, out->col_diff, out->d_flux, out->d_proj, out->ld_proj);
  CU_CHECK(cudaGetLastError());
  g_launches++;
  return 0;
}

int b200spec_plan_num_res(const b200spec_plan *plan) { return plan ? plan->num_res : 0; }

int b200spec_plan_num_bands(const b200spec_plan *plan, int32_t res) {
  if (!plan || res < 0 || res >= plan->num_res) return -1;
  return plan->res[res].num_bands;
}

int b200spec_plan_filterbank_layout(const b200spec_plan *plan, int32_t res, int32_t out[6]) {
  if (!plan || !out || res < 0 || res >= plan->num_res) return fail(B200SPEC_ERR_ARG, "plan/out is NULL or res out of range");
  const ResPlan &r = plan->res[res];
  out[0] = r.fb_L;
  out[1] = r.fb_ns;
  out[2] = r.fb_kmin;
  out[3] = r.kmax;
  out[4] = r.fb_ndirect;
  out[5] = r.fb_ndw;
  return 0;
}

int64_t b200spec_launch_count(void) { return g_launches.load(); }

int b200spec_task_plan(int32_t num_sms, int32_t workers_per_sm, double overhead_frames, int64_t total_frames,
                       int32_t n_clips, int32_t warmup_rows, int32_t out[3]) {
  if (!out) return fail(B200SPEC_ERR_ARG, "out is NULL");
  if (num_sms < 1 || workers_per_sm < 1 || total_frames < 0 || n_clips < 0 || warmup_rows < 0 || !(overhead_frames >= 0.0))
    return fail(B200SPEC_ERR_ARG, "task plan: sizes out of range");
  const TaskPlan tp = choose_tasks(num_sms, workers_per_sm, overhead_frames, total_frames, n_clips, warmup_rows);
  out[0] = tp.chunk;
  out[1] = tp.chunk_small;
  out[2] = tp.tail_clips;
  return 0;
}

}  // extern "C"
