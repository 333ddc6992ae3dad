// bulk_stage.cuh -- plan tables (twiddles, window, slab filterbank, band table) are brought into shared memory by
// the TMA unit: ONE elected thread of the CTA issues a one-dimensional bulk copy per table
// (cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes, SASS: UBLKCP), all of them completing on one
// mbarrier that every thread then waits on.  The copies run while the CTA zeroes its work buffers, and the
// prologue no longer costs 50+ KB of LDG / STS round trips per CTA (it is on the critical path of small batches).
// Requirements met by the plan: every table is its own cudaMalloc (256-byte aligned) padded to a multiple of 16
// bytes (upload() in b200spec.cu), every shared-memory slot is 16-byte aligned and padded the same way.
// (The SAMPLES are not staged this way: frame starts int(n * hop) - F/2 are not 16-byte aligned and the measured
// gain of staging them was nil -- DESIGN.md section 4.)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace b2 {

constexpr int kStaticSmemBytes = 16;   // the staging mbarrier (+ k_front_multi's task slot): static shared memory of every kernel

#if defined(__CUDACC__)

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t arrivals) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(arrivals) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");   // visible to the async proxy
}

__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}

// bytes: a multiple of 16; dst and src 16-byte aligned
__device__ __forceinline__ void bulk_load_1d(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}

// bounded wait: a lost transaction traps instead of hanging the GPU
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  for (uint32_t spins = 0; !mbar_try_wait(bar, parity); ++spins)
    if (spins > (1u << 26)) __trap();
}

// `tables(fn)` calls fn(dst_smem, src_gmem, bytes) once per table.  Thread 0 initialises the barrier, announces the
// byte total and starts the copies; the caller runs bulk_stage_wait (all threads) before the first table read.
template <class Tables>
__device__ __forceinline__ void bulk_stage_begin(uint64_t *bar, Tables tables) {
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    uint32_t total = 0;
    tables([&](void *, const void *, uint32_t bytes) { total += (bytes + 15u) & ~15u; });
    mbar_expect_tx(bar, total);
    tables([&](void *dst, const void *src, uint32_t bytes) {
      if (bytes > 0) bulk_load_1d(dst, src, (bytes + 15u) & ~15u, bar);
    });
  }
}

__device__ __forceinline__ void bulk_stage_wait(uint64_t *bar) {
  __syncthreads();        // the barrier's initialisation (thread 0) precedes every wait
  mbar_wait(bar, 0);
}

#endif  // __CUDACC__

}  // namespace b2
