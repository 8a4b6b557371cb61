// pipeline.cuh -- tile / pipeline constants and the PTX helpers every kernel shares: mbarriers, TMA bulk
// copies (cp.async.bulk; SASS UBLKCP), streaming 128-bit accesses, the world staging, fp64 atomics.
#pragma once
#include <cuda_runtime.h>

#include <cstddef>

#include "transition.cuh"

namespace shipenv {

constexpr int kThreads = 256;  // envs per tile
constexpr int kStages = 4;     // input tiles in flight per CTA
constexpr int kStepBlock = kThreads + 32;  // 8 consumer warps + 1 TMA producer warp

struct alignas(16) TileStage {  // one 256-env tile, filled by three TMA bulk copies
  uint4 ships[kThreads];
  uint32_t clock[kThreads];
  int16_t act[kThreads];
};
static_assert(sizeof(TileStage) == 5632, "stage layout");
constexpr size_t kPipeBytes = sizeof(TileStage) * kStages + 2 * 8 * kStages + 4 * kStages;  // slots + full/empty mbarriers + tile ids

// ------------------------------------------------------------------------------------------
// PTX helpers: mbarrier + TMA bulk copies (cp.async.bulk; SASS UBLKCP)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return uint32_t(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void consumer_sync() {  // the 8 consumer warps only (the producer warp has left)
  asm volatile("bar.sync 1, 256;" ::: "memory");
}
__device__ __forceinline__ void mbar_wait_s(uint32_t bar_saddr, uint32_t parity) {   // shared-address forms
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAITS_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONES_%=;\n"
      "bra WAITS_%=;\n"
      "DONES_%=:\n"
      "}\n" ::"r"(bar_saddr), "r"(parity) : "memory");
}
// producer-side wait: sleeps ~1 us between polls (a tile takes a CTA about that long and kStages tiles are
// buffered) so that the waiting warp does not eat issue slots
__device__ __forceinline__ void mbar_wait_backoff(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAITB_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONEB_%=;\n"
      "nanosleep.u32 1024;\n"
      "bra WAITB_%=;\n"
      "DONEB_%=:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void mbar_arrive_s(uint32_t bar_saddr) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar_saddr) : "memory");
}
__device__ __forceinline__ uint4 lds_u4(uint32_t saddr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(saddr));
  return v;
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t saddr) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(saddr));
  return v;
}
__device__ __forceinline__ float lds_f32(uint32_t saddr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(saddr));
  return v;
}
__device__ __forceinline__ double lds_f64(uint32_t saddr) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(saddr));
  return v;
}
__device__ __forceinline__ int lds_s16(uint32_t saddr) {
  short v;
  asm volatile("ld.shared.s16 %0, [%1];" : "=h"(v) : "r"(saddr));
  return int(v);
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// global -> shared bulk copy, completion signalled on an mbarrier (bytes % 16 == 0, 16-B aligned)
__device__ __forceinline__ void tma_load_bulk(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
// shared -> global bulk copy in the current bulk group
__device__ __forceinline__ void tma_store_bulk(void* gmem_dst, const void* smem_src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gmem_dst), "r"(smem_u32(smem_src)),
               "r"(bytes)
               : "memory");
}
// pull a contiguous global range towards L2 (bytes % 16 == 0, 16-B aligned)
__device__ __forceinline__ void tma_prefetch_l2(const void* gmem_src, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(gmem_src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int kPending>
__device__ __forceinline__ void tma_store_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(kPending) : "memory");
}
template <int kPending>
__device__ __forceinline__ void tma_store_wait() {      // all but the newest kPending groups have been written out
  asm volatile("cp.async.bulk.wait_group %0;" ::"n"(kPending) : "memory");
}
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// streaming 128-bit accesses: ship records are touched once per launch, keep them out of L1
__device__ __forceinline__ uint4 ld_stream_u4(const uint4* p) {
  uint4 v;
  asm volatile("ld.global.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ void st_stream_u4(uint4* p, const uint4& v) {
  asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w)
               : "memory");
}

// Stage the world blob: thread 0 issues one bulk copy; everybody waits on the mbarrier.
// Returns after the data is visible to all threads of the CTA.
__device__ __forceinline__ void stage_world(void* smem_world, const void* gmem_world, uint32_t bytes, uint64_t* bar) {
  if (threadIdx.x == 0) mbar_init(bar, 1);
  __syncthreads();
  if (threadIdx.x == 0) {
    mbar_expect_tx(bar, bytes);
    tma_load_bulk(smem_world, gmem_world, bytes, bar);
  }
  mbar_wait(bar, 0);
}

__device__ __forceinline__ void atomic_max_f64(double* addr, double v) {
  unsigned long long* a = reinterpret_cast<unsigned long long*>(addr);
  unsigned long long old = *a, assumed;
  do {
    assumed = old;
    if (__longlong_as_double((long long)assumed) >= v) break;
    old = atomicCAS(a, assumed, (unsigned long long)__double_as_longlong(v));
  } while (assumed != old);
}
__device__ __forceinline__ void atomic_min_f64(double* addr, double v) {
  unsigned long long* a = reinterpret_cast<unsigned long long*>(addr);
  unsigned long long old = *a, assumed;
  do {
    assumed = old;
    if (__longlong_as_double((long long)assumed) <= v) break;
    old = atomicCAS(a, assumed, (unsigned long long)__double_as_longlong(v));
  } while (assumed != old);
}

}  // namespace shipenv
