// Exact-fp32 tiled GEMM (FFMA) with fused epilogues and deterministic split-K.
//
// This is the always-correct arithmetic backend of nnr_gemm (algo NNR_GEMM_SIMT_FP32) and the
// on-device checker for the tcgen05 backend (gemm_tc.cu).  128x128x16 tiles, 256 threads, 8x8
// outputs per thread, register-prefetch double buffering, 16-byte global/shared accesses when the
// operand is aligned and scalar guarded accesses otherwise (K = 225 in SUE is not 16B friendly).
#include "common.cuh"
#include "../../include/nnr_b200.h"

#define BM 128
#define BN 128
#define BK 16
#define PADM 4
#define GT 256

#include "gemm_epilogue.cuh"

// Load 4 consecutive-in-`contiguous dim` elements of an operand tile.
//   KC = true : contiguous along k  (element (r, k) at base[r*ld + k])
//   KC = false: contiguous along r  (element (r, k) at base[k*ld + r])
template <bool KC>
__device__ __forceinline__ float4 load4(const float* __restrict__ base, int64_t ld, int r, int k, int R, int K, bool vec) {
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (KC) {
    if (r < R) {
      const float* p = base + (size_t)r * ld + k;
      if (vec && k + 3 < K) v = __ldg(reinterpret_cast<const float4*>(p));
      else {
        if (k < K) v.x = __ldg(p);
        if (k + 1 < K) v.y = __ldg(p + 1);
        if (k + 2 < K) v.z = __ldg(p + 2);
        if (k + 3 < K) v.w = __ldg(p + 3);
      }
    }
  } else {
    if (k < K) {
      const float* p = base + (size_t)k * ld + r;
      if (vec && r + 3 < R) v = __ldg(reinterpret_cast<const float4*>(p));
      else {
        if (r < R) v.x = __ldg(p);
        if (r + 1 < R) v.y = __ldg(p + 1);
        if (r + 2 < R) v.z = __ldg(p + 2);
        if (r + 3 < R) v.w = __ldg(p + 3);
      }
    }
  }
  return v;
}

// AKC: op(A)[m,k] contiguous in k (transA == 0);  BKC: op(B)[k,n] contiguous in k (transB == 1)
template <bool AKC, bool BKC>
__global__ void __launch_bounds__(GT) gemm_simt_kernel(const float* __restrict__ A, int64_t lda, const float* __restrict__ B,
                                                       int64_t ldb, int M, int N, int K, const int32_t* __restrict__ m_dev,
                                                       const int32_t* __restrict__ k_dev, int splits, bool vecA, bool vecB,
                                                       EpiP epi, float* __restrict__ partial) {
  __shared__ __align__(16) float As[2][BK][BM + PADM];
  __shared__ __align__(16) float Bs[2][BK][BN + PADM];
  if (m_dev) M = min(M, *m_dev);
  if (k_dev) K = min(K, *k_dev);
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  if (m0 >= M) return;
  // split-K over the EFFECTIVE contraction length (device-side token count), BK-aligned chunks
  const int k_chunk = (((K + splits - 1) / splits) + BK - 1) / BK * BK;
  const int kb = blockIdx.z * k_chunk;
  const int ke = min(K, kb + k_chunk);
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;

  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  // per-thread global->smem assignments: 2 float4 per operand per tile
  float4 ra[2], rb[2];
  auto gload = [&](int k0) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      if (AKC) { int r = (tid >> 2) + 64 * i, kq = (tid & 3) * 4; ra[i] = load4<true>(A, lda, m0 + r, k0 + kq, M, ke, vecA); }
      else     { int k = (tid >> 5) + 8 * i, rq = (tid & 31) * 4; ra[i] = load4<false>(A, lda, m0 + rq, k0 + k, M, ke, vecA); }
      if (BKC) { int r = (tid >> 2) + 64 * i, kq = (tid & 3) * 4; rb[i] = load4<true>(B, ldb, n0 + r, k0 + kq, N, ke, vecB); }
      else     { int k = (tid >> 5) + 8 * i, rq = (tid & 31) * 4; rb[i] = load4<false>(B, ldb, n0 + rq, k0 + k, N, ke, vecB); }
    }
  };
  auto sstore = [&](int buf) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      if (AKC) { int r = (tid >> 2) + 64 * i, kq = (tid & 3) * 4;
        As[buf][kq + 0][r] = ra[i].x; As[buf][kq + 1][r] = ra[i].y; As[buf][kq + 2][r] = ra[i].z; As[buf][kq + 3][r] = ra[i].w; }
      else     { int k = (tid >> 5) + 8 * i, rq = (tid & 31) * 4; *reinterpret_cast<float4*>(&As[buf][k][rq]) = ra[i]; }
      if (BKC) { int r = (tid >> 2) + 64 * i, kq = (tid & 3) * 4;
        Bs[buf][kq + 0][r] = rb[i].x; Bs[buf][kq + 1][r] = rb[i].y; Bs[buf][kq + 2][r] = rb[i].z; Bs[buf][kq + 3][r] = rb[i].w; }
      else     { int k = (tid >> 5) + 8 * i, rq = (tid & 31) * 4; *reinterpret_cast<float4*>(&Bs[buf][k][rq]) = rb[i]; }
    }
  };

  const int ntiles = (ke - kb + BK - 1) / BK;
  if (ntiles > 0) {
    gload(kb);
    sstore(0);
  }
  __syncthreads();
  for (int t = 0; t < ntiles; ++t) {
    const int buf = t & 1;
    if (t + 1 < ntiles) gload(kb + (t + 1) * BK);
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
      float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][64 + ty * 4]);
      float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
      float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][k][64 + tx * 4]);
      float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    if (t + 1 < ntiles) sstore(buf ^ 1);
    __syncthreads();
  }

#pragma unroll
  for (int i = 0; i < 8; ++i) {
    int m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      int n = n0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4));
      if (n >= N) continue;
      if (partial) partial[((size_t)blockIdx.z * M + m) * N + n] = acc[i][j];   // note: M here is the effective M
      else epi_store(epi, m, n, acc[i][j]);
    }
  }
}

static int simt_splits(const nnr_gemm_args* a) {
  // split-K only when the output grid cannot fill the machine and the contraction is long
  long tiles = (long)((a->M + BM - 1) / BM) * ((a->N + BN - 1) / BN);
  if (tiles >= 148 || a->K < 2048) return 1;
  long s = (296 + tiles - 1) / tiles;
  long maxs = a->K / 512;
  if (s > maxs) s = maxs;
  if (s > 64) s = 64;
  return s < 1 ? 1 : (int)s;
}

size_t nnr_gemm_simt_workspace_bytes(const nnr_gemm_args* a) {
  int s = simt_splits(a);
  return s > 1 ? (size_t)s * a->M * a->N * sizeof(float) : 0;
}

int nnr_gemm_simt(const nnr_gemm_args* a, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  EpiP e = make_epi(a);
  int splits = simt_splits(a);
  float* partial = nullptr;
  if (splits > 1) {
    size_t need = (size_t)splits * a->M * a->N * sizeof(float);
    NNR_REQUIRE(a->workspace && a->workspace_bytes >= need, NNR_ERR_WORKSPACE,
                "nnr_gemm: split-K needs %zu workspace bytes (got %zu)", need, a->workspace_bytes);
    partial = (float*)a->workspace;
  }
  bool akc = a->transA == 0, bkc = a->transB != 0;
  bool vecA = nnr_aligned16(a->A) && (a->lda % 4 == 0);
  bool vecB = nnr_aligned16(a->B) && (a->ldb % 4 == 0);
  dim3 grid((a->N + BN - 1) / BN, (a->M + BM - 1) / BM, splits);
#define LAUNCH(AK, BKc)                                                                                          \
  gemm_simt_kernel<AK, BKc><<<grid, GT, 0, st>>>(a->A, a->lda, a->B, a->ldb, a->M, a->N, a->K, a->m_dev, a->k_dev, \
                                                 splits, vecA, vecB, e, partial)
  void* ph = nnr_prof_begin(3, 0.0, st);
  if (akc && bkc) LAUNCH(true, true);
  else if (akc && !bkc) LAUNCH(true, false);
  else if (!akc && bkc) LAUNCH(false, true);
  else LAUNCH(false, false);
#undef LAUNCH
  nnr_prof_end(ph, st);
  NNR_LAUNCH_CHECK("gemm_simt_kernel");
  if (splits > 1) {
    size_t tot = (size_t)a->M * a->N;
    gemm_splitk_reduce_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(partial, splits, a->M, a->N, a->m_dev, e);
    NNR_LAUNCH_CHECK("gemm_splitk_reduce_kernel");
  }
  return 0;
}
