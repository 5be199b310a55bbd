// Shared helpers for the nnr_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

#define NNR_ERR_ARG (-1)
#define NNR_ERR_ALIGN (-2)
#define NNR_ERR_WORKSPACE (-3)
#define NNR_ERR_UNSUPPORTED (-4)

void nnr_set_error(const char* fmt, ...);
void nnr_count_launch(int n);
// kernel-level timing hooks (api_common.cu); tags: 0 gemm_tc_kernel, 1 tc_split_kernel, 2 tc_splitk_reduce, 3 gemm_simt_kernel
void* nnr_prof_begin(int tag, double flops, void* stream);
void nnr_prof_end(void* h, void* stream);

#define NNR_REQUIRE(cond, code, ...)              \
  do {                                            \
    if (!(cond)) {                                \
      nnr_set_error(__VA_ARGS__);                 \
      return (code);                              \
    }                                             \
  } while (0)

// Launch check: returns the (positive) cudaError_t of the launch, if any.
#define NNR_LAUNCH_CHECK(name)                                                   \
  do {                                                                           \
    cudaError_t e__ = cudaGetLastError();                                        \
    nnr_count_launch(1);                                                         \
    if (e__ != cudaSuccess) {                                                    \
      nnr_set_error("%s: launch failed: %s", name, cudaGetErrorString(e__));     \
      return (int)e__;                                                           \
    }                                                                            \
  } while (0)

#define NNR_CUDA(call)                                                           \
  do {                                                                           \
    cudaError_t e__ = (call);                                                    \
    if (e__ != cudaSuccess) {                                                    \
      nnr_set_error("%s failed: %s", #call, cudaGetErrorString(e__));            \
      return (int)e__;                                                           \
    }                                                                            \
  } while (0)

static inline bool nnr_aligned16(const void* p) { return (((uintptr_t)p) & 15) == 0; }

#ifdef __CUDACC__
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }
// GEMM epilogues: the same functions from the hardware exp2 / reciprocal units (2 MUFU + 3 FP32 instructions instead of the
// ~35 of expf + an IEEE division with its slow-path call).  ncu showed the sigmoid-gate epilogue issuing ~50 instructions per
// output element and the K = 400 gate GEMM bound by exactly that (0.245 ms against 0.104 ms for the same product without an
// epilogue).  Absolute error < 2e-7 (sigmoid), < 4e-7 (tanh) over the whole range; saturation and +-inf are handled by the
// units (exp -> 0 / inf, 1 / inf -> 0).
__device__ __forceinline__ float sigmoid_fast(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }
__device__ __forceinline__ float tanh_fast(float x) { return 1.0f - __fdividef(2.0f, 1.0f + __expf(2.0f * x)); }

// Counter-based RNG for dropout masks: one 32-bit hash per element, reproducible from
// (seed, element index) so the backward pass regenerates the forward mask instead of storing it.
__device__ __forceinline__ uint32_t mix32(uint64_t x) {
  x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33;
  return (uint32_t)x;
}
__device__ __forceinline__ uint64_t mix64(uint64_t x) {
  x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33;
  return x;
}
// Dropout seeds cross the C ABI as uint64.  A value below 2^62 is used as is.  With bit 62 set (NNR_SEED_INDIRECT) the
// seed is INDIRECT: bits 0..47 hold the device address of a uint64 base that the caller advances on the device once per
// step, bits 48..61 a site id; the kernels then use hash(base + site).  A step captured in a CUDA graph replays with
// fresh masks this way (kernel arguments are frozen at capture), and the backward of the same step resolves to the same
// value as its forward because the base only moves between steps.
#define NNR_SEED_INDIRECT (1ULL << 62)
__device__ __forceinline__ uint64_t nnr_resolve_seed(uint64_t s) {
  if (s & NNR_SEED_INDIRECT) {
    const uint64_t base = *reinterpret_cast<const uint64_t*>(s & 0x0000FFFFFFFFFFFFULL);
    return mix64(base + ((s >> 48) & 0x3FFFULL) * 0x9E3779B97F4A7C15ULL) >> 2;
  }
  return s;
}
// keep-scale: 0 if dropped, 1/(1-p) if kept.  p == 0 -> always 1.
// ONE 64-bit hash serves four consecutive elements (16 uniform bits each): element i of a tensor uses field (i & 3) of
// hash(i >> 2).  Every mask in the library follows this rule, so a kernel that walks aligned quads (the embedding gather /
// scatter, the GEMM's relu-residual epilogue and its backward pre-pass: there the per-element hash was the bulk of the
// instruction count) computes one hash per float4 with dropout_scale4 and still agrees element for element with a kernel
// that asks for single elements (dropout_scale).
__device__ __forceinline__ void dropout_scale4(uint64_t seed, uint64_t idx4, float p, float inv_keep, float (&s)[4]) {
  const uint64_t r = mix64(seed * 0x9E3779B97F4A7C15ULL + idx4);
  const uint32_t thr = (uint32_t)(p * 65536.0f);
#pragma unroll
  for (int i = 0; i < 4; ++i) s[i] = ((uint32_t)(r >> (16 * i)) & 0xffffu) < thr ? 0.0f : inv_keep;
}
__device__ __forceinline__ float dropout_scale(uint64_t seed, uint64_t idx, float p, float inv_keep) {
  if (p <= 0.0f) return 1.0f;
  const uint64_t r = mix64(seed * 0x9E3779B97F4A7C15ULL + (idx >> 2));
  return ((uint32_t)(r >> (16 * (idx & 3))) & 0xffffu) < (uint32_t)(p * 65536.0f) ? 0.0f : inv_keep;
}
__device__ __forceinline__ float dropout_scale_e4(uint64_t seed, uint64_t idx, float p, float inv_keep) {
  return dropout_scale(seed, idx, p, inv_keep);
}
#endif
