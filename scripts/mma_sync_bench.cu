// Throughput probe: legacy mma.sync.m16n8k8 tf32 (and m16n8k16 bf16) on sm_100a, per SM.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k_tf32(float* out, int iters) {
  float c[8][4];
  for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) c[i][j] = 0.f;
  unsigned a0 = threadIdx.x, a1 = a0 * 3, a2 = a0 * 5, a3 = a0 * 7, b0 = a0 * 11, b1 = a0 * 13;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
      asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
  }
  float s = 0;
  for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) s += c[i][j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_bf16(float* out, int iters) {
  float c[8][4];
  for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) c[i][j] = 0.f;
  unsigned a0 = threadIdx.x, a1 = a0 * 3, a2 = a0 * 5, a3 = a0 * 7, b0 = a0 * 11, b1 = a0 * 13;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
      asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
  }
  float s = 0;
  for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) s += c[i][j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_ffma(float* out, int iters) {
  float c[32];
  for (int i = 0; i < 32; ++i) c[i] = 0.f;
  float a = threadIdx.x * 0.001f, b = 1.0001f;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 32; ++i) c[i] = fmaf(a, b + i, c[i]);
  }
  float s = 0;
  for (int i = 0; i < 32; ++i) s += c[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_ffma2(float* out, int iters) {
  float2 c[16];
  for (int i = 0; i < 16; ++i) c[i] = make_float2(0.f, 0.f);
  float2 a = make_float2(threadIdx.x * 0.001f, threadIdx.x * 0.002f), b = make_float2(1.0001f, 1.0002f);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) c[i] = __ffma2_rn(a, b, c[i]);
  }
  float s = 0;
  for (int i = 0; i < 16; ++i) s += c[i].x + c[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
  int sms = 148, warps = 8, iters = 20000;
  float* out; cudaMalloc(&out, sizeof(float) * sms * 4 * warps * 32);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  int clk_khz; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  for (int kind = 0; kind < 4; ++kind) {
    for (int rep = 0; rep < 2; ++rep) {
      cudaEventRecord(e0);
      if (kind == 0) k_tf32<<<sms * 2, warps * 32>>>(out, iters);
      else if (kind == 1) k_bf16<<<sms * 2, warps * 32>>>(out, iters);
      else if (kind == 2) k_ffma<<<sms * 2, warps * 32>>>(out, iters);
      else k_ffma2<<<sms * 2, warps * 32>>>(out, iters);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      double macs = (kind == 0 ? 1024.0 : kind == 1 ? 2048.0 : 0) * 8.0 * iters * warps * sms * 2;
      if (kind >= 2) macs = 32.0 * 32 * (double)iters * warps * sms * 2;
      if (rep) printf("%s: %.3f ms, %.1f TMAC/s, %.1f MAC/clk/SM (at %d MHz nominal)\n", kind == 0 ? "mma.sync tf32 m16n8k8" : kind == 1 ? "mma.sync bf16 m16n8k16" : kind == 2 ? "ffma" : "ffma2",
                      ms, macs / ms / 1e9, macs / (ms * 1e-3) / sms / (clk_khz * 1e3), clk_khz / 1000);
    }
  }
  return 0;
}
