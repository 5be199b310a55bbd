// tcgen05 / TMEM / TMA GEMM backend of nnr_gemm (sm_100a).
//
//   C[M,N] = epi( op(A)[M,K] * op(B)[K,N] )      fp32 in, fp32 out, fp32-grade accuracy
//
// Arithmetic: 3xTF32.  Every fp32 operand x is split into hi = rna_tf32(x) and lo = rna_tf32(x - hi);
// D += A_hi*B_hi + A_hi*B_lo + A_lo*B_hi on the 5th-generation tensor cores (tcgen05.mma kind::tf32,
// fp32 accumulation in TMEM).  The dropped lo*lo term and the rounding of lo are ~2^-21 relative, i.e.
// the result is within a few fp32 ulps of an exact-fp32 product and passes the same parity tests as
// the FFMA backend.  algo NNR_GEMM_TC_BF16 runs one pass of kind::f16 on bf16-rounded operands.
//
// Pipeline (one 128 x BLOCK_N output tile per CTA, two CTAs co-resident per SM so one CTA's
// epilogue overlaps the other's main loop):
//   pre-pass   tc_split_kernel: op(X) -> K-major planes [2][rows][Kp] (hi, lo) in the workspace; handles
//              the transposed operands of dgrad/wgrad, zero-fills the K tail, honours m_dev/k_dev.
//   warp 0     TMA producer: cp.async.bulk.tensor (3-D map: k, row, plane; SWIZZLE_128B) into a ring of
//              smem stages, completion on mbarriers (expect_tx).
//   warp 1     allocates TMEM, then one elected lane issues tcgen05.mma (M=128, N=BLOCK_N, K=8 per
//              instruction, 4 per 128-byte swizzle atom) and tcgen05.commit to free stages / signal the
//              epilogue.
//   warps 2-5  epilogue: tcgen05.ld the accumulator rows (one row per thread), apply the fused epilogue
//              (bias / tanh / relu+residual / sigmoid gate / add) and store.
// Split-K (wgrad: K = tokens) writes partials and reuses the deterministic reduce of the FFMA backend.
#include "common.cuh"
#include "gemm_epilogue.cuh"
#include <cuda.h>
#include <cuda_bf16.h>
#include <stdlib.h>

#define TC_BM 128
#define TC_BK 32          // fp32 elements per k-block = one 128-byte swizzle atom row
#define TC_THREADS 192
#define TC_TMEM_COLS 256

// ------------------------------------------------------------------------------------------------
// device helpers (raw PTX)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra WAIT_DONE;\n"
      "bra WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
template <bool BF16>
__device__ __forceinline__ void tc_mma(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  if (BF16) {
    asm volatile(
        "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc),
        "r"(accumulate)
        : "memory");
  } else {
    asm volatile(
        "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc),
        "r"(accumulate)
        : "memory");
  }
}
// 32 lanes x 16 consecutive fp32 columns: thread i of the warp receives row (lane base + i)
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// K-major, SWIZZLE_128B shared-memory matrix descriptor (sm_100 format): rows of 128 bytes, 8-row
// groups 1024 bytes apart (SBO), LBO unused, version 1, layout type 2 (= 128B swizzle).
__device__ __forceinline__ uint64_t make_kmajor_sw128_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);        // start address, bits [0,14)
  d |= (uint64_t)0 << 16;                              // leading byte offset (ignored for swizzled K-major)
  d |= (uint64_t)(1024 >> 4) << 32;                    // stride byte offset, bits [32,46)
  d |= (uint64_t)1 << 46;                              // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;                              // SWIZZLE_128B
  return d;
}

struct TcParams {
  int M, N, K;                 // capacities
  const int32_t* m_dev;
  const int32_t* k_dev;
  int block_n, stages, splits, passes;
  uint32_t idesc;
  float* partial;
  EpiP epi;
};

template <bool BF16>
__global__ void __launch_bounds__(TC_THREADS) gemm_tc_kernel(const __grid_constant__ CUtensorMap map_a,
                                                             const __grid_constant__ CUtensorMap map_b, TcParams p) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  int M = p.M, K = p.K;
  if (p.m_dev) M = min(M, *p.m_dev);
  if (p.k_dev) K = min(K, *p.k_dev);
  const int m0 = blockIdx.y * TC_BM, n0 = blockIdx.x * p.block_n;
  if (m0 >= M) return;                                   // whole CTA leaves before any barrier / TMEM use
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // smem carve-up: [stages][A 16 KB | B block_n*128 B] (1024-aligned), then barriers
  const uint32_t a_bytes = TC_BM * 128, b_bytes = (uint32_t)p.block_n * 128;
  const uint32_t stage_bytes = (a_bytes + b_bytes + 1023) & ~1023u;
  unsigned char* tiles = (unsigned char*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(tiles + (size_t)p.stages * stage_bytes);
  uint64_t* empty_bar = full_bar + p.stages;
  uint64_t* tmem_full_bar = empty_bar + p.stages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

  // k-block range of this split
  constexpr int KELEM = BF16 ? 2 * TC_BK : TC_BK;        // elements per 128-byte k-block
  const int nkb_total = (K + KELEM - 1) / KELEM;
  const int kb_per = (nkb_total + p.splits - 1) / p.splits;
  const int kb0 = blockIdx.z * kb_per;
  const int kb1 = min(nkb_total, kb0 + kb_per);
  const int iters = max(0, kb1 - kb0) * p.passes;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
    for (int s = 0; s < p.stages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(tmem_full_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(TC_TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      for (int it = 0; it < iters; ++it) {
        const int s = it % p.stages;
        const uint32_t ph = (it / p.stages) & 1;
        mbar_wait(&empty_bar[s], ph ^ 1);
        const int kb = kb0 + it / p.passes;
        const int pass = it % p.passes;
        // 3xTF32 passes per k-block: (A_hi,B_hi) (A_hi,B_lo) (A_lo,B_hi)
        const int plane_a = (pass == 2) ? 1 : 0;
        const int plane_b = (pass == 1) ? 1 : 0;
        unsigned char* sa = tiles + (size_t)s * stage_bytes;
        unsigned char* sb = sa + a_bytes;
        mbar_expect_tx(&full_bar[s], a_bytes + b_bytes);
        tma_load_3d(&map_a, &full_bar[s], sa, kb * KELEM, m0, plane_a);
        tma_load_3d(&map_b, &full_bar[s], sb, kb * KELEM, n0, plane_b);
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      for (int it = 0; it < iters; ++it) {
        const int s = it % p.stages;
        const uint32_t ph = (it / p.stages) & 1;
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        const uint32_t sa = smem_u32(tiles + (size_t)s * stage_bytes);
        const uint64_t da = make_kmajor_sw128_desc(sa);
        const uint64_t db = make_kmajor_sw128_desc(sa + a_bytes);
#pragma unroll
        for (int k = 0; k < 4; ++k) {       // 4 MMAs of K = 32 bytes inside the 128-byte swizzle atom
          tc_mma<BF16>(tmem_base, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), p.idesc, (it > 0 || k > 0) ? 1u : 0u);
        }
        tc_commit(&empty_bar[s]);           // frees the smem stage once these MMAs have read it
      }
      tc_commit(tmem_full_bar);             // accumulator complete
    }
  } else {
    // ===== epilogue: warps 2..5 own TMEM lane quarters (warp % 4) =====
    const int q = warp & 3;
    const int m = m0 + q * 32 + lane;
    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
    for (int c0 = 0; c0 < p.block_n; c0 += 16) {
      float v[16];
      if (iters > 0) tmem_ld16(lane_addr + (uint32_t)c0, v);
      else {
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = 0.f;
      }
      if (m < M) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const int n = n0 + c0 + i;
          if (n < p.N) {
            if (p.partial) p.partial[((size_t)blockIdx.z * M + m) * p.N + n] = v[i];
            else epi_store(p.epi, m, n, v[i]);
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TC_TMEM_COLS) : "memory");
  }
}

// ------------------------------------------------------------------------------------------------
// pre-pass: op(X)[R,K] -> K-major planes.  TF32: planes (hi, lo) fp32 [2][R][Kp].  BF16: one bf16 plane.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float rna_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}
// contiguous-in-k source (src[r*ld + k]); one thread per 4 k's.  Kz: k >= Kz is written as zero.
template <bool BF16>
__global__ void tc_split_rowmajor_kernel(const float* __restrict__ src, int64_t ld, int R, int K, int Kp,
                                         const int32_t* __restrict__ r_dev, const int32_t* __restrict__ k_dev, bool vec,
                                         void* __restrict__ out, size_t plane_stride) {
  if (r_dev) R = min(R, *r_dev);
  int Kv = k_dev ? min(K, *k_dev) : K;                      // valid k
  int Kw = k_dev ? min(Kp, (Kv + 2 * TC_BK - 1) / (2 * TC_BK) * (2 * TC_BK)) : Kp;   // written k (zero tail)
  const int kq = (blockIdx.y * blockDim.x + threadIdx.x) * 4;
  const int r = blockIdx.x;
  if (r >= R || kq >= Kw) return;
  float x[4] = {0.f, 0.f, 0.f, 0.f};
  const float* s = src + (size_t)r * ld + kq;
  if (vec && kq + 3 < Kv) { float4 v = __ldg(reinterpret_cast<const float4*>(s)); x[0] = v.x; x[1] = v.y; x[2] = v.z; x[3] = v.w; }
  else {
#pragma unroll
    for (int i = 0; i < 4; ++i) if (kq + i < Kv) x[i] = __ldg(s + i);
  }
  if (BF16) {
    __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(out) + (size_t)r * Kp + kq;
#pragma unroll
    for (int i = 0; i < 4; ++i) if (kq + i < Kp) o[i] = __float2bfloat16_rn(x[i]);
  } else {
    float* hi = reinterpret_cast<float*>(out) + (size_t)r * Kp + kq;
    float* lo = hi + plane_stride;
    float h[4], l[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) { h[i] = rna_tf32(x[i]); l[i] = rna_tf32(x[i] - h[i]); }
    *reinterpret_cast<float4*>(hi) = make_float4(h[0], h[1], h[2], h[3]);     // Kp % 4 == 0, 16B aligned planes
    *reinterpret_cast<float4*>(lo) = make_float4(l[0], l[1], l[2], l[3]);
  }
}
// transposed source (src[k*ld + r]) -> planes[r][k] through a 32x32 smem tile
template <bool BF16>
__global__ void tc_split_transpose_kernel(const float* __restrict__ src, int64_t ld, int R, int K, int Kp,
                                          const int32_t* __restrict__ r_dev, const int32_t* __restrict__ k_dev,
                                          void* __restrict__ out, size_t plane_stride) {
  __shared__ float tile[32][33];
  if (r_dev) R = min(R, *r_dev);
  int Kv = k_dev ? min(K, *k_dev) : K;
  int Kw = k_dev ? min(Kp, (Kv + 2 * TC_BK - 1) / (2 * TC_BK) * (2 * TC_BK)) : Kp;
  const int k0 = blockIdx.y * 32, r0 = blockIdx.x * 32;
  if (k0 >= Kw || r0 >= R) return;
  const int tx = threadIdx.x, ty = threadIdx.y;           // 32 x 8
  for (int i = ty; i < 32; i += 8) {
    int k = k0 + i, r = r0 + tx;
    tile[i][tx] = (k < Kv && r < R) ? __ldg(src + (size_t)k * ld + r) : 0.f;
  }
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    int r = r0 + i, k = k0 + tx;
    if (r < R && k < Kw) {
      float x = tile[tx][i];
      if (BF16) reinterpret_cast<__nv_bfloat16*>(out)[(size_t)r * Kp + k] = __float2bfloat16_rn(x);
      else {
        float h = rna_tf32(x);
        float* hi = reinterpret_cast<float*>(out) + (size_t)r * Kp + k;
        hi[0] = h;
        hi[plane_stride] = rna_tf32(x - h);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
    else (void)cudaGetLastError();
  }
  return fn;
}

static size_t up(size_t x, size_t a) { return (x + a - 1) / a * a; }

struct TcPlan {
  bool bf16;
  int block_n, stages, splits, passes;
  int Kp;                       // plane pitch in elements
  size_t a_plane, b_plane;      // elements per plane
  size_t a_off, b_off, partial_off, total;
  size_t smem;
};

static int pick_block_n(int N) {
  int best = 64;
  double best_cost = 1e30;
  for (int bn = 256; bn >= 32; bn -= 16) {
    double padded = (double)((N + bn - 1) / bn) * bn;
    double cost = padded * (1.0 + 40.0 / bn);
    if (cost < best_cost - 1e-9) { best_cost = cost; best = bn; }
  }
  return best;
}

static TcPlan make_plan(const nnr_gemm_args* a, bool bf16) {
  TcPlan pl;
  pl.bf16 = bf16;
  pl.block_n = pick_block_n(a->N);
  pl.passes = bf16 ? 1 : 3;
  const int kelem = bf16 ? 2 * TC_BK : TC_BK;            // elements per 128-byte row
  pl.Kp = (int)up((size_t)a->K, bf16 ? 8 : 4);           // 16-byte row pitch
  // split-K when the output grid is small and the contraction long (wgrad)
  long tiles = (long)((a->M + TC_BM - 1) / TC_BM) * ((a->N + pl.block_n - 1) / pl.block_n);
  int splits = 1;
  if (tiles < 148 && a->K >= 4096) {
    splits = (int)((296 + tiles - 1) / tiles);
    int maxs = a->K / (kelem * 8);
    if (splits > maxs) splits = maxs;
    if (splits < 1) splits = 1;
  }
  // bound the length of one TMEM accumulation chain: the tensor core truncates when it adds into the fp32
  // accumulator, which biases long sums (~2^-24 per MMA); partial sums are combined in exact fp32 order instead
  // (only for the small-output / long-contraction GEMMs = wgrad; K <= 1600 elsewhere on this path)
  const int max_chain = 1024;
  if (tiles < 148 && (a->K + splits - 1) / splits > max_chain) splits = (a->K + max_chain - 1) / max_chain;
  if (splits > 1024) splits = 1024;
  pl.splits = splits;
  size_t stage = up((size_t)TC_BM * 128 + (size_t)pl.block_n * 128, 1024);
  int stages = (int)((100 * 1024) / stage);              // <= ~100 KB -> two CTAs per SM
  if (stages < 2) stages = 2;
  if (stages > 8) stages = 8;
  pl.stages = stages;
  pl.smem = 1024 + (size_t)stages * stage + (2 * stages + 1) * 8 + 16;
  const size_t esz = bf16 ? 2 : 4;
  const int nplanes = bf16 ? 1 : 2;
  pl.a_plane = (size_t)a->M * pl.Kp;
  pl.b_plane = (size_t)a->N * pl.Kp;
  size_t o = 0;
  pl.a_off = o; o = up(o + pl.a_plane * esz * nplanes, 1024);
  pl.b_off = o; o = up(o + pl.b_plane * esz * nplanes, 1024);
  pl.partial_off = o;
  if (splits > 1) o = up(o + (size_t)splits * a->M * a->N * sizeof(float), 1024);
  pl.total = o;
  return pl;
}

int nnr_gemm_tc_supported(const nnr_gemm_args* a) {
  static int disabled = -1;
  if (disabled < 0) { const char* e = getenv("NNR_DISABLE_TC"); disabled = (e && e[0] == '1') ? 1 : 0; }
  if (disabled) return 0;
  if (!get_encode()) return 0;
  // tiny problems are launch-bound either way; the FFMA kernel handles them exactly
  if ((double)a->M * a->N * a->K < 2.0e6) return 0;
  if (a->K < 8) return 0;
  return 1;
}

size_t nnr_gemm_tc_workspace_bytes(const nnr_gemm_args* a) {
  int algo = a->algo;
  return make_plan(a, algo == NNR_GEMM_TC_BF16).total;
}

static int encode_map(CUtensorMap* map, void* base, bool bf16, int Kp, int rows, int nplanes, int box_rows) {
  cuuint64_t gdim[3] = {(cuuint64_t)Kp, (cuuint64_t)rows, (cuuint64_t)nplanes};
  const size_t esz = bf16 ? 2 : 4;
  cuuint64_t gstr[2] = {(cuuint64_t)Kp * esz, (cuuint64_t)Kp * esz * (cuuint64_t)rows};
  cuuint32_t box[3] = {(cuuint32_t)(bf16 ? 2 * TC_BK : TC_BK), (cuuint32_t)box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = get_encode()(map, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, base, gdim, gstr,
                            box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { nnr_set_error("nnr_gemm(tc): cuTensorMapEncodeTiled failed (%d)", (int)r); return NNR_ERR_ARG; }
  return 0;
}

template <bool BF16>
static int split_operand(const float* X, int64_t ld, bool k_contig, int R, int K, int Kp, const int32_t* r_dev,
                         const int32_t* k_dev, void* out, size_t plane_stride, cudaStream_t st) {
  if (k_contig) {
    bool vec = nnr_aligned16(X) && (ld % 4 == 0);
    dim3 grid(R, (Kp / 4 + 63) / 64);
    tc_split_rowmajor_kernel<BF16><<<grid, 64, 0, st>>>(X, ld, R, K, Kp, r_dev, k_dev, vec, out, plane_stride);
  } else {
    dim3 grid((R + 31) / 32, (Kp + 31) / 32);
    tc_split_transpose_kernel<BF16><<<grid, dim3(32, 8), 0, st>>>(X, ld, R, K, Kp, r_dev, k_dev, out, plane_stride);
  }
  NNR_LAUNCH_CHECK("tc_split_kernel");
  return 0;
}

template <bool BF16>
static int run_tc(const nnr_gemm_args* a, cudaStream_t st) {
  TcPlan pl = make_plan(a, BF16);
  NNR_REQUIRE(a->workspace && a->workspace_bytes >= pl.total, NNR_ERR_WORKSPACE, "nnr_gemm(tc): workspace %zu < %zu",
              a->workspace_bytes, pl.total);
  NNR_REQUIRE(nnr_aligned16(a->workspace), NNR_ERR_ALIGN, "nnr_gemm(tc): workspace must be 16B aligned");
  char* ws = (char*)a->workspace;
  void* pa = ws + pl.a_off;
  void* pb = ws + pl.b_off;
  // op(A)[m,k]: contiguous in k when transA == 0.  op(B)[k,n] as rows n: contiguous in k when transB != 0.
  int rc = split_operand<BF16>(a->A, a->lda, a->transA == 0, a->M, a->K, pl.Kp, a->m_dev, a->k_dev, pa, pl.a_plane, st);
  if (rc) return rc;
  rc = split_operand<BF16>(a->B, a->ldb, a->transB != 0, a->N, a->K, pl.Kp, nullptr, a->k_dev, pb, pl.b_plane, st);
  if (rc) return rc;
  CUtensorMap map_a, map_b;
  const int nplanes = BF16 ? 1 : 2;
  rc = encode_map(&map_a, pa, BF16, pl.Kp, a->M, nplanes, TC_BM);
  if (rc) return rc;
  rc = encode_map(&map_b, pb, BF16, pl.Kp, a->N, nplanes, pl.block_n);
  if (rc) return rc;
  TcParams p;
  p.M = a->M; p.N = a->N; p.K = a->K; p.m_dev = a->m_dev; p.k_dev = a->k_dev;
  p.block_n = pl.block_n; p.stages = pl.stages; p.splits = pl.splits; p.passes = pl.passes;
  // instruction descriptor: D=f32 (bits 4-5 = 1), A/B format (tf32 = 2, bf16 = 1) at bits 7-9 / 10-12, K-major A and B,
  // N >> 3 at bits 17-22, M >> 4 at bits 24-28
  const uint32_t fmt = BF16 ? 1u : 2u;
  p.idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(pl.block_n >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
  p.partial = pl.splits > 1 ? (float*)(ws + pl.partial_off) : nullptr;
  p.epi = make_epi(a);
  static bool attr_set[2] = {false, false};
  if (!attr_set[BF16 ? 1 : 0]) {
    NNR_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<BF16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024));
    attr_set[BF16 ? 1 : 0] = true;
  }
  NNR_REQUIRE(pl.smem <= 112 * 1024, NNR_ERR_UNSUPPORTED, "nnr_gemm(tc): smem plan too large");
  dim3 grid((a->N + pl.block_n - 1) / pl.block_n, (a->M + TC_BM - 1) / TC_BM, pl.splits);
  gemm_tc_kernel<BF16><<<grid, TC_THREADS, pl.smem, st>>>(map_a, map_b, p);
  NNR_LAUNCH_CHECK("gemm_tc_kernel");
  if (pl.splits > 1) {
    size_t tot = (size_t)a->M * a->N;
    gemm_splitk_reduce_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(p.partial, pl.splits, a->M, a->N, a->m_dev, p.epi);
    NNR_LAUNCH_CHECK("gemm_splitk_reduce_kernel");
  }
  return 0;
}

int nnr_gemm_tc(const nnr_gemm_args* a, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (a->algo == NNR_GEMM_TC_BF16) return run_tc<true>(a, st);
  return run_tc<false>(a, st);
}
