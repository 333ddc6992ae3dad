// front_inst.cuh -- instantiates k_front for one frame size (one translation unit per size so the
// sizes compile in parallel) and exposes a plain launcher to b200spec.cu.
#pragma once
#include "frontend_kernel.cuh"
#include "frontend_pair_kernel.cuh"
#include "frontend_warp_kernel.cuh"

namespace b2 {

// groups of 128 threads per CTA for each frame size (bounded by shared memory and registers)
template <int F>
struct GroupsPerCta {
#ifndef B2_GROUPS
#define B2_GROUPS 4
#endif
  static constexpr int value = (F == 8192) ? 3 : B2_GROUPS;       // 8192: three groups when the magnitude rows are
  static constexpr int fallback = (F == 8192) ? 2 : B2_GROUPS;    // short (band-limited filterbank), else two
};

constexpr size_t kMaxSmemPerCta = 227 * 1024 - kStaticSmemBytes;   // opt-in dynamic shared memory per CTA on sm_100, less the kernels' static words

struct LaunchResult {
  cudaError_t err;
  int launches;
};

template <int F, int IN, int MODE, int G>
static cudaError_t launch_one_g(FrontParams &p, int num_sms, long long task_bound, cudaStream_t st) {
  const size_t smem = front_smem_layout<F>(p, MODE, G);
  if (smem > kMaxSmemPerCta) return cudaErrorInvalidConfiguration;   // reported as "unsupported" by the caller
  auto kern = k_front<F, IN, MODE, G>;
  // per device and cheap; set on every launch so multi-device processes stay correct
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  long long ctas = (task_bound + G - 1) / G;
  int grid = (int)(ctas < num_sms ? (ctas < 1 ? 1 : ctas) : num_sms);
  kern<<<grid, kGroupThreads * G, smem, st>>>(p);
  return cudaGetLastError();
}

template <int F, int IN, int MODE>
static cudaError_t launch_one(FrontParams &p, int num_sms, long long task_bound, cudaStream_t st) {
  cudaError_t e = launch_one_g<F, IN, MODE, GroupsPerCta<F>::value>(p, num_sms, task_bound, st);
  if (e == cudaErrorInvalidConfiguration && GroupsPerCta<F>::fallback != GroupsPerCta<F>::value)
    e = launch_one_g<F, IN, MODE, GroupsPerCta<F>::fallback>(p, num_sms, task_bound, st);
  return e;
}

template <int F>
static cudaError_t launch_front_size(int in, int mode, FrontParams &p, int num_sms, long long task_bound,
                                     cudaStream_t st) {
  switch (in * 2 + mode) {
    case IN_F32_MONO * 2 + MODE_LOGFILT: return launch_one<F, IN_F32_MONO, MODE_LOGFILT>(p, num_sms, task_bound, st);
    case IN_F32_MONO * 2 + MODE_SPECTRUM: return launch_one<F, IN_F32_MONO, MODE_SPECTRUM>(p, num_sms, task_bound, st);
    case IN_F32_STEREO * 2 + MODE_LOGFILT: return launch_one<F, IN_F32_STEREO, MODE_LOGFILT>(p, num_sms, task_bound, st);
    case IN_F32_STEREO * 2 + MODE_SPECTRUM: return launch_one<F, IN_F32_STEREO, MODE_SPECTRUM>(p, num_sms, task_bound, st);
    case IN_I16_MONO * 2 + MODE_LOGFILT: return launch_one<F, IN_I16_MONO, MODE_LOGFILT>(p, num_sms, task_bound, st);
    case IN_I16_MONO * 2 + MODE_SPECTRUM: return launch_one<F, IN_I16_MONO, MODE_SPECTRUM>(p, num_sms, task_bound, st);
    case IN_I16_STEREO * 2 + MODE_LOGFILT: return launch_one<F, IN_I16_STEREO, MODE_LOGFILT>(p, num_sms, task_bound, st);
    case IN_I16_STEREO * 2 + MODE_SPECTRUM: return launch_one<F, IN_I16_STEREO, MODE_SPECTRUM>(p, num_sms, task_bound, st);
  }
  return cudaErrorInvalidValue;
}

// groups per CTA of the pair kernel: its FFT buffer holds two frames, so frame 4096 fits three groups
template <int F>
struct PairGroupsPerCta {
#ifndef B2_PAIR_GROUPS_4096
#define B2_PAIR_GROUPS_4096 3     // 33 KB FFT buffer + 16.5 KB magnitudes per group: shared memory is full at three
#endif
#ifndef B2_PAIR_GROUPS_1024
#define B2_PAIR_GROUPS_1024 5     // 96 registers per thread, no spills: measured 2.14 ms against 2.26 ms with four groups
#endif
#ifndef B2_PAIR_GROUPS_2048
#define B2_PAIR_GROUPS_2048 4
#endif
  static constexpr int value = (F == 4096) ? B2_PAIR_GROUPS_4096 : (F == 1024) ? B2_PAIR_GROUPS_1024 : B2_PAIR_GROUPS_2048;
};

// cudaErrorInvalidConfiguration = does not fit in shared memory (the caller falls back to k_front)
template <int F, int IN>
static cudaError_t launch_pair_one(FrontParams &p, int num_sms, long long task_bound, cudaStream_t st) {
  constexpr int G = PairGroupsPerCta<F>::value;
  const size_t smem = pair_smem_layout<F>(p, G);
  if (smem > kMaxSmemPerCta) return cudaErrorInvalidConfiguration;
  auto kern = k_front_pair<F, IN, G>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  long long ctas = (task_bound + G - 1) / G;
  int grid = (int)(ctas < num_sms ? (ctas < 1 ? 1 : ctas) : num_sms);
  kern<<<grid, kGroupThreads * G, smem, st>>>(p);
  return cudaGetLastError();
}

template <int F>
static cudaError_t launch_pair_size(int in, FrontParams &p, int num_sms, long long task_bound, cudaStream_t st) {
  switch (in) {
    case IN_F32_MONO: return launch_pair_one<F, IN_F32_MONO>(p, num_sms, task_bound, st);
    case IN_F32_STEREO: return launch_pair_one<F, IN_F32_STEREO>(p, num_sms, task_bound, st);
    case IN_I16_MONO: return launch_pair_one<F, IN_I16_MONO>(p, num_sms, task_bound, st);
    case IN_I16_STEREO: return launch_pair_one<F, IN_I16_STEREO>(p, num_sms, task_bound, st);
  }
  return cudaErrorInvalidValue;
}

template <int F>
static cudaError_t launch_warp_size(int in, FrontParams &p, int num_sms, long long task_bound, cudaStream_t st);

// warps per CTA of the warp-per-FFT kernel (frames 1024 / 2048): 16 x 32 threads at <= 128 registers, ~10 KB of shared memory each
#ifndef B2_WARP_NW_1024
#define B2_WARP_NW_1024 16
#endif
#ifndef B2_WARP_NW_2048
#define B2_WARP_NW_2048 16
#endif
template <int F>
struct WarpKernelWarps {
  static constexpr int value = (F == 1024) ? B2_WARP_NW_1024 : B2_WARP_NW_2048;
};

template <int F, int IN>
static cudaError_t launch_warp_one(FrontParams &p, int num_sms, long long task_bound, cudaStream_t st) {
  constexpr int NW = WarpKernelWarps<F>::value;
  const size_t smem = warp_smem_layout<F>(p, NW);
  if (smem > kMaxSmemPerCta) return cudaErrorInvalidConfiguration;
  auto kern = k_front_warp<F, IN, NW>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  long long ctas = (task_bound + NW - 1) / NW;
  int grid = (int)(ctas < num_sms ? (ctas < 1 ? 1 : ctas) : num_sms);
  kern<<<grid, 32 * NW, smem, st>>>(p);
  return cudaGetLastError();
}

template <int F>
static cudaError_t launch_warp_size(int in, FrontParams &p, int num_sms, long long task_bound, cudaStream_t st) {
  switch (in) {
    case IN_F32_MONO: return launch_warp_one<F, IN_F32_MONO>(p, num_sms, task_bound, st);
    case IN_F32_STEREO: return launch_warp_one<F, IN_F32_STEREO>(p, num_sms, task_bound, st);
    case IN_I16_MONO: return launch_warp_one<F, IN_I16_MONO>(p, num_sms, task_bound, st);
    case IN_I16_STEREO: return launch_warp_one<F, IN_I16_STEREO>(p, num_sms, task_bound, st);
  }
  return cudaErrorInvalidValue;
}

// groups per CTA of the one-launch kernel: its group blocks are sized for frame 4096 (33 KB FFT buffer + 16 KB magnitudes)
constexpr int kMultiGroups = 3;

template <int IN>
static cudaError_t launch_multi_one(MultiParams &m, int num_sms, long long task_bound, cudaStream_t st) {
  constexpr int G = kMultiGroups;
  const size_t smem = multi_smem_layout(m, G);
  if (smem > kMaxSmemPerCta) return cudaErrorInvalidConfiguration;
  auto kern = k_front_multi<IN, G>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  long long ctas = (task_bound + G - 1) / G;
  int grid = (int)(ctas < num_sms ? (ctas < 1 ? 1 : ctas) : num_sms);
  kern<<<grid, kGroupThreads * G, smem, st>>>(m);
  return cudaGetLastError();
}

}  // namespace b2

// defined in front_multi.cu
cudaError_t b2_launch_multi(int in, b2::MultiParams &m, int num_sms, long long task_bound, cudaStream_t st);

cudaError_t b2_launch_pair_1024(int in, b2::FrontParams &p, int num_sms, long long task_bound, cudaStream_t st);
cudaError_t b2_launch_warp_1024(int in, b2::FrontParams &p, int num_sms, long long task_bound, cudaStream_t st);
cudaError_t b2_launch_warp_2048(int in, b2::FrontParams &p, int num_sms, long long task_bound, cudaStream_t st);
cudaError_t b2_launch_pair_2048(int in, b2::FrontParams &p, int num_sms, long long task_bound, cudaStream_t st);
cudaError_t b2_launch_pair_4096(int in, b2::FrontParams &p, int num_sms, long long task_bound, cudaStream_t st);

// declared here, defined in front_f<F>.cu
cudaError_t b2_launch_front_1024(int in, int mode, b2::FrontParams &p, int num_sms, long long task_bound, cudaStream_t st);
cudaError_t b2_launch_front_2048(int in, int mode, b2::FrontParams &p, int num_sms, long long task_bound, cudaStream_t st);
cudaError_t b2_launch_front_4096(int in, int mode, b2::FrontParams &p, int num_sms, long long task_bound, cudaStream_t st);
cudaError_t b2_launch_front_8192(int in, int mode, b2::FrontParams &p, int num_sms, long long task_bound, cudaStream_t st);
