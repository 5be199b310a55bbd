// Cluster / distributed-shared-memory primitives shared by the LSTM kernels (lstm.cu: exact fp32 FFMA
// variant, lstm_mma.cu: tensor-core variant).
#pragma once
#include "common.cuh"
#include <cooperative_groups.h>
namespace cg = cooperative_groups;

__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }

// ---- DSMEM producer/consumer primitives: st.async + mbarrier complete_tx instead of barrier.cluster.
// (barrier.cluster.arrive.release compiles to MEMBAR.ALL.GPU, which makes every step wait for the drain of the
//  stash stores to global memory; the transaction barrier orders exactly the shared::cluster bytes we exchange.)
__device__ __forceinline__ uint32_t smem_addr_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t mapa_u32(uint32_t local, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local), "r"(rank));
  return r;
}
__device__ __forceinline__ void lbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr_u32(bar)), "r"(count));
}
__device__ __forceinline__ void lbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr_u32(bar)), "r"(bytes) : "memory");
}
// default (acquire.cta) wait: the complete_tx of st.async makes the bytes visible to the waiter, exactly as for
// TMA loads; a cluster-scope acquire would add CCTL.IVALL (L1 invalidate) to every step
__device__ __forceinline__ void lbar_wait_cluster(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "LW_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra LW_DONE;\n"
      "bra LW_LOOP;\n"
      "LW_DONE:\n"
      "}\n" ::"r"(smem_addr_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void rbar_arrive_release(uint32_t remote_bar) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote_bar) : "memory");
}
// no release fence: used where the arrive only guards a write-after-read (the reads completed before the CTA barrier
// that precedes it); the .release form drains every outstanding global store of the thread first
__device__ __forceinline__ void rbar_arrive_relaxed(uint32_t remote_bar) {
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote_bar) : "memory");
}
__device__ __forceinline__ void st_async_f4(uint32_t remote_addr, float4 v, uint32_t remote_bar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];" ::"r"(remote_addr),
               "r"(__float_as_uint(v.x)), "r"(__float_as_uint(v.y)), "r"(__float_as_uint(v.z)), "r"(__float_as_uint(v.w)),
               "r"(remote_bar)
               : "memory");
}

// gate non-linearities on the SFU (ex2 + rcp): absolute error ~1e-7, far inside the 1e-4 parity budget
__device__ __forceinline__ float fast_sigmoid(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }
__device__ __forceinline__ float fast_tanh(float x) { return 1.0f - __fdividef(2.0f, 1.0f + __expf(2.0f * x)); }

__device__ __forceinline__ void st_async_v2(uint32_t remote_addr, uint32_t x, uint32_t y, uint32_t remote_bar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.b32 [%0], {%1, %2}, [%3];" ::"r"(remote_addr), "r"(x),
               "r"(y), "r"(remote_bar)
               : "memory");
}
__device__ __forceinline__ void st_async_b32(uint32_t remote_addr, uint32_t x, uint32_t remote_bar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];" ::"r"(remote_addr), "r"(x),
               "r"(remote_bar)
               : "memory");
}
__device__ __forceinline__ void st_async_u4(uint32_t remote_addr, uint4 v, uint32_t remote_bar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];" ::"r"(remote_addr),
               "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "r"(remote_bar)
               : "memory");
}
