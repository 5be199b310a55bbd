// Bidirectional LSTM recurrence, forward and backward-through-time, as persistent thread-block
// cluster kernels (sm_100a).  Replaces cuDNN's RNN behind nn.LSTM at newsEncoders.py:66-67,119-127.
//
// Design
//   * One cluster of CL = 4 CTAs owns a tile of sequences and one direction.  CTA r of the cluster owns
//     hidden units [r*H/4, (r+1)*H/4): the i,f,g,o rows of those units, i.e. a [4*H/4, H] slice of
//     W_hh = 160 KB fp32 for H = 200, which stays in shared memory for every time step of every tile the
//     cluster processes (the kernel is persistent over tiles).
//   * forward step (32-row tiles): z = gx_t + h_{t-1} W_slice^T, register tiled FFMA (2 units x 4 gates x
//     4 rows per thread), fused sigmoid/tanh/cell update with c_t kept in registers, then the new h slice is
//     broadcast into all four CTAs' (double-buffered) h tiles through distributed shared memory and one
//     barrier.cluster per step publishes it.  (A 64-row single-buffer variant was measured: better FFMA
//     efficiency but a longer per-step latency, which loses on MIND-like length distributions where the
//     longest sequences set the critical path.)
//   * backward step (32-row tiles): each CTA turns dh_t (+ recurrent part) into d(pre-activations) for its
//     own units, multiplies them with its W slice to get a PARTIAL dh_{t-1} over all H units, and
//     reduce-scatters the partials to the owning CTAs through DSMEM (fixed summation order ->
//     deterministic).
//   * the per-step exchange uses st.async into the peers' shared memory with mbarrier transaction counts
//     (and, in BPTT, a "recv is free" barrier fed by remote arrives) - no barrier.cluster inside the time loop.
//   * variable lengths: rows are tiled in length-sorted order; a row is active for its own `len` steps only
//     (packed-sequence semantics); the reverse direction walks t = len-1 .. 0.
//   * the input projection gx (a dense GEMM) is computed outside (nnr_gemm); the activated gates
//     overwrite gx in place as the stash for BPTT, and BPTT overwrites them with dL/dgx.
#include "common.cuh"
#include "../../include/nnr_b200.h"
#include <cooperative_groups.h>
namespace cg = cooperative_groups;

template <int HID_, int CL_, int MT_, int UPT_>   // hidden, cluster size, rows per tile, units per thread (phase 1)
struct LCfg {
  static constexpr int HID = HID_, CL = CL_, MT = MT_, UPT = UPT_;
  static constexpr int UPC = HID / CL;        // hidden units per CTA
  static constexpr int COLS = 4 * UPC;        // gate columns per CTA
  static constexpr int UG = UPC / UPT;        // unit groups
  static constexpr int RG = MT / 4;           // row groups of 4
  static constexpr int NWORK = UG * RG;       // working threads
  static constexpr int NT = ((NWORK + 31) / 32) * 32;
  static_assert(HID % CL == 0 && UPC % UPT == 0 && MT % 4 == 0 && HID % 4 == 0, "unsupported LSTM geometry");
  static constexpr size_t FWD_SMEM = sizeof(float) * ((size_t)HID * COLS + 2 * (size_t)HID * MT) + 3 * MT * sizeof(int);
  static constexpr size_t BWD_SMEM = sizeof(float) * ((size_t)COLS * HID + (size_t)COLS * MT + (size_t)CL * UPC * MT) + 3 * MT * sizeof(int);
};

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
__device__ __forceinline__ void lbar_wait_cluster(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "LW_LOOP:\n"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%0], %1;\n"
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
__device__ __forceinline__ void st_async_f4(uint32_t remote_addr, float4 v, uint32_t remote_bar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];" ::"r"(remote_addr),
               "r"(__float_as_uint(v.x)), "r"(__float_as_uint(v.y)), "r"(__float_as_uint(v.z)), "r"(__float_as_uint(v.w)),
               "r"(remote_bar)
               : "memory");
}

// gate non-linearities on the SFU (ex2 + rcp): absolute error ~1e-7, far inside the 1e-4 parity budget
__device__ __forceinline__ float fast_sigmoid(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }
__device__ __forceinline__ float fast_tanh(float x) { return 1.0f - __fdividef(2.0f, 1.0f + __expf(2.0f * x)); }

// ------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------
template <class C>
__global__ void __launch_bounds__(C::NT, 1)
lstm_fwd_kernel(float* __restrict__ gx, const float* __restrict__ w_hh, const int32_t* __restrict__ len,
                const int32_t* __restrict__ off, const int32_t* __restrict__ order, int N, int ntiles,
                float* __restrict__ h_out, float* __restrict__ c_stash, float* __restrict__ c_n, int* __restrict__ tile_counter) {
  constexpr int HID = C::HID, CL = C::CL, MT = C::MT, UPC = C::UPC, COLS = C::COLS, UP = C::UG;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* Wt = reinterpret_cast<float*>(smem_raw);                 // [HID][COLS]
  float* hT = Wt + (size_t)HID * COLS;                            // [2][HID][MT]
  int* s_row = reinterpret_cast<int*>(hT + 2 * (size_t)HID * MT); // [MT]
  int* s_len = s_row + MT;
  int* s_off = s_len + MT;

  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int cluster_id = blockIdx.x / CL;
  const int nclusters = gridDim.x / CL;
  const int dir = cluster_id & 1;
  const int tid = threadIdx.x;
  const bool worker = tid < C::NWORK;
  const int up = tid % UP, rg = tid / UP;
  const int j0 = 2 * up;                       // local unit of this thread (and j0+1)
  const int unit0 = rank * UPC + j0;           // global hidden unit

  // W slice -> smem, transposed to k-major: Wt[k][g*UPC + j] = W_hh[dir][g*HID + rank*UPC + j][k]
  {
    const float* W = w_hh + (size_t)dir * 4 * HID * HID;
    for (int idx = tid; idx < COLS * HID; idx += C::NT) {
      int c = idx / HID, k = idx - c * HID;
      int g = c / UPC, j = c - g * UPC;
      Wt[(size_t)k * COLS + c] = W[(size_t)(g * HID + rank * UPC + j) * HID + k];
    }
  }
  __shared__ __align__(8) uint64_t hfull[2];        // "all of h for the next step has landed in buffer b"
  uint32_t remote_hT[CL], remote_bar[CL];
#pragma unroll
  for (int d = 0; d < CL; ++d) {
    remote_hT[d] = mapa_u32(smem_addr_u32(hT), d);
    remote_bar[d] = mapa_u32(smem_addr_u32(&hfull[0]), d);
  }
  if (tid == 0) {
    lbar_init(&hfull[0], 1);
    lbar_init(&hfull[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  uint32_t hph[2] = {0u, 0u};                       // phase parity per buffer (identical in every thread)

  const size_t GS = (size_t)2 * 4 * HID;  // gx row stride (both directions)
  // dynamic tile scheduling: tiles are sorted longest-first, so handing the next tile to whichever cluster
  // becomes free is longest-processing-time-first list scheduling (one atomic per tile, broadcast via DSMEM)
  __shared__ int s_tile;
  int* remote_tile[CL];
#pragma unroll
  for (int d = 0; d < CL; ++d) remote_tile[d] = cluster.map_shared_rank(&s_tile, d);
  (void)nclusters;
  for (;;) {
    if (rank == 0 && tid == 0) {
      int t = atomicAdd(&tile_counter[dir], 1);
#pragma unroll
      for (int d = 0; d < CL; ++d) *remote_tile[d] = t;
    }
    cluster_arrive();
    cluster_wait();
    const int tile = s_tile;
    if (tile >= ntiles) break;
    __syncthreads();
    if (tid < MT) {
      int i = tile * MT + tid;
      int r = (i < N) ? order[i] : -1;
      s_row[tid] = r;
      s_len[tid] = (r >= 0) ? len[r] : 0;
      s_off[tid] = (r >= 0) ? off[r] : 0;
    }
    for (int idx = tid; idx < HID * MT; idx += C::NT) hT[idx] = 0.f;   // h_0 = 0 in buffer 0
    __syncthreads();
    int maxlen = 0;
    for (int i = 0; i < MT; ++i) maxlen = max(maxlen, s_len[i]);

    int rlen[4], roff[4], rrow[4];
    float cst[2][4], hst[2][4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      rlen[i] = worker ? s_len[4 * rg + i] : 0;
      roff[i] = worker ? s_off[4 * rg + i] : 0;
      rrow[i] = worker ? s_row[4 * rg + i] : -1;
      cst[0][i] = cst[1][i] = hst[0][i] = hst[1][i] = 0.f;
    }
    // prefetch gx for step 0
    float2 gxr[4][4];  // [gate][row]
    auto load_gx = [&](int s) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        if (s < rlen[i]) {
          int t = dir ? (rlen[i] - 1 - s) : s;
          const float* p = gx + ((size_t)roff[i] + t) * GS + (size_t)dir * 4 * HID + unit0;
#pragma unroll
          for (int g = 0; g < 4; ++g) gxr[g][i] = *reinterpret_cast<const float2*>(p + g * HID);
        } else {
#pragma unroll
          for (int g = 0; g < 4; ++g) gxr[g][i] = make_float2(0.f, 0.f);
        }
      }
    };
    if (worker) load_gx(0);

    for (int s = 0; s < maxlen; ++s) {
      const int cur = s & 1, nxt = cur ^ 1;
      if (tid == 0 && s + 1 < maxlen) lbar_expect_tx(&hfull[nxt], (uint32_t)(HID * MT * sizeof(float)));  // arm for h_{s+1}
      if (s > 0) { lbar_wait_cluster(&hfull[cur], hph[cur]); hph[cur] ^= 1u; }                          // h_s landed
      if (worker) {
        // accumulators as packed pairs (unit j0, unit j0+1): one FFMA2 (fma.rn.f32x2) per gate and row
        float2 acc2[4][4];
#pragma unroll
        for (int g = 0; g < 4; ++g)
#pragma unroll
          for (int i = 0; i < 4; ++i) acc2[g][i] = gxr[g][i];
        if (s + 1 < maxlen) load_gx(s + 1);   // in flight during the k loop
        const float* hb = hT + (size_t)cur * HID * MT + 4 * rg;
        const float* wb = Wt + j0;
#pragma unroll 4
        for (int k = 0; k < HID; ++k) {
          const float4 hv = *reinterpret_cast<const float4*>(hb + (size_t)k * MT);
          const float2 h0 = make_float2(hv.x, hv.x), h1 = make_float2(hv.y, hv.y);
          const float2 h2 = make_float2(hv.z, hv.z), h3 = make_float2(hv.w, hv.w);
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            const float2 w = *reinterpret_cast<const float2*>(wb + (size_t)k * COLS + g * UPC);
            acc2[g][0] = __ffma2_rn(w, h0, acc2[g][0]);
            acc2[g][1] = __ffma2_rn(w, h1, acc2[g][1]);
            acc2[g][2] = __ffma2_rn(w, h2, acc2[g][2]);
            acc2[g][3] = __ffma2_rn(w, h3, acc2[g][3]);
          }
        }
        float acc[4][2][4];
#pragma unroll
        for (int g = 0; g < 4; ++g)
#pragma unroll
          for (int i = 0; i < 4; ++i) { acc[g][0][i] = acc2[g][i].x; acc[g][1][i] = acc2[g][i].y; }
        // gates, state update, stash
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          if (s < rlen[i]) {
            int t = dir ? (rlen[i] - 1 - s) : s;
            size_t p = (size_t)roff[i] + t;
            float ig[2], fg[2], gg[2], og[2];
#pragma unroll
            for (int u = 0; u < 2; ++u) {
              ig[u] = fast_sigmoid(acc[0][u][i]);
              fg[u] = fast_sigmoid(acc[1][u][i]);
              gg[u] = fast_tanh(acc[2][u][i]);
              og[u] = fast_sigmoid(acc[3][u][i]);
              cst[u][i] = fg[u] * cst[u][i] + ig[u] * gg[u];
              hst[u][i] = og[u] * fast_tanh(cst[u][i]);
            }
            float* gp = gx + p * GS + (size_t)dir * 4 * HID + unit0;
            *reinterpret_cast<float2*>(gp) = make_float2(ig[0], ig[1]);
            *reinterpret_cast<float2*>(gp + HID) = make_float2(fg[0], fg[1]);
            *reinterpret_cast<float2*>(gp + 2 * HID) = make_float2(gg[0], gg[1]);
            *reinterpret_cast<float2*>(gp + 3 * HID) = make_float2(og[0], og[1]);
            *reinterpret_cast<float2*>(c_stash + (p * 2 + dir) * HID + unit0) = make_float2(cst[0][i], cst[1][i]);
            *reinterpret_cast<float2*>(h_out + p * 2 * HID + (size_t)dir * HID + unit0) = make_float2(hst[0][i], hst[1][i]);
            if (s == rlen[i] - 1)
              *reinterpret_cast<float2*>(c_n + (size_t)rrow[i] * 2 * HID + (size_t)dir * HID + unit0) =
                  make_float2(cst[0][i], cst[1][i]);
          }
        }
        // send the new h slice into every CTA's next buffer (frozen rows re-send their state); the bytes are
        // counted on the receiver's transaction barrier.  Double buffering + the data dependency of the
        // recurrence make the write-after-read safe without a second barrier.
        if (s + 1 < maxlen) {
          float4 h0 = make_float4(hst[0][0], hst[0][1], hst[0][2], hst[0][3]);
          float4 h1 = make_float4(hst[1][0], hst[1][1], hst[1][2], hst[1][3]);
          const uint32_t o = (uint32_t)(((size_t)nxt * HID * MT + (size_t)unit0 * MT + 4 * rg) * sizeof(float));
#pragma unroll
          for (int d = 0; d < CL; ++d) {
            st_async_f4(remote_hT[d] + o, h0, remote_bar[d] + nxt * 8);
            st_async_f4(remote_hT[d] + o + MT * sizeof(float), h1, remote_bar[d] + nxt * 8);
          }
        }
      }
    }
  }
  // no CTA may exit while a peer can still write into its shared memory
  cluster_arrive();
  cluster_wait();
}

// ------------------------------------------------------------------------------------------------
// backward through time
// ------------------------------------------------------------------------------------------------
template <class C>
__global__ void __launch_bounds__(C::NT, 1)
lstm_bwd_kernel(float* __restrict__ gates, const float* __restrict__ c_stash, const float* __restrict__ w_hh,
                const int32_t* __restrict__ len, const int32_t* __restrict__ off, const int32_t* __restrict__ order,
                int N, int ntiles, const float* __restrict__ dh, const float* __restrict__ dcn, int* __restrict__ tile_counter) {
  constexpr int HID = C::HID, CL = C::CL, MT = C::MT, UPC = C::UPC, COLS = C::COLS, UG = C::UG;
  static_assert(C::UPT == 1, "backward kernel is written for 1 unit per thread");
  constexpr int KG = HID / 4;                  // phase 2: groups of 4 k's
  static_assert(KG * C::RG <= C::NT, "backward phase-2 mapping does not fit");
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* Wc = reinterpret_cast<float*>(smem_raw);                 // [COLS][HID]  (c-major)
  float* dzT = Wc + (size_t)COLS * HID;                           // [COLS][MT]
  float* recv = dzT + (size_t)COLS * MT;                          // [CL][UPC][MT] partial dh from each peer
  int* s_row = reinterpret_cast<int*>(recv + (size_t)CL * UPC * MT);
  int* s_len = s_row + MT;
  int* s_off = s_len + MT;

  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int cluster_id = blockIdx.x / CL;
  const int nclusters = gridDim.x / CL;
  const int dir = cluster_id & 1;
  const int tid = threadIdx.x;
  const bool worker = tid < C::NWORK;
  const int j = tid % UG, rg = tid / UG;       // phase 1: one hidden unit x 4 rows
  const int unit = rank * UPC + j;
  const bool worker2 = tid < KG * C::RG;       // phase 2: 4 k's x 4 rows
  const int kg = tid % KG, rg2 = tid / KG;

  {
    const float* W = w_hh + (size_t)dir * 4 * HID * HID;
    for (int idx = tid; idx < COLS * HID; idx += C::NT) {
      int c = idx / HID, k = idx - c * HID;
      int g = c / UPC, jj = c - g * UPC;
      Wc[idx] = W[(size_t)(g * HID + rank * UPC + jj) * HID + k];
    }
  }
  // rfull: the CL partial slices of dh for the next iteration have landed in my recv (transaction barrier)
  // rfree: every CTA of the cluster has finished reading ITS recv, so mine may send (CL arrivals)
  __shared__ __align__(8) uint64_t rbar[2];
  uint64_t* rfull = &rbar[0];
  uint64_t* rfree = &rbar[1];
  uint32_t remote_recv[CL], remote_rfull[CL], remote_rfree[CL];
#pragma unroll
  for (int d = 0; d < CL; ++d) {
    remote_recv[d] = mapa_u32(smem_addr_u32(recv), d);
    remote_rfull[d] = mapa_u32(smem_addr_u32(rfull), d);
    remote_rfree[d] = mapa_u32(smem_addr_u32(rfree), d);
  }
  if (tid == 0) {
    lbar_init(rfull, 1);
    lbar_init(rfree, CL);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  uint32_t ph_full = 0u, ph_free = 0u;

  const size_t GS = (size_t)2 * 4 * HID;
  // dynamic tile scheduling: tiles are sorted longest-first, so handing the next tile to whichever cluster
  // becomes free is longest-processing-time-first list scheduling (one atomic per tile, broadcast via DSMEM)
  __shared__ int s_tile;
  int* remote_tile[CL];
#pragma unroll
  for (int d = 0; d < CL; ++d) remote_tile[d] = cluster.map_shared_rank(&s_tile, d);
  (void)nclusters;
  for (;;) {
    if (rank == 0 && tid == 0) {
      int t = atomicAdd(&tile_counter[dir], 1);
#pragma unroll
      for (int d = 0; d < CL; ++d) *remote_tile[d] = t;
    }
    cluster_arrive();
    cluster_wait();
    const int tile = s_tile;
    if (tile >= ntiles) break;
    __syncthreads();
    if (tid < MT) {
      int i = tile * MT + tid;
      int r = (i < N) ? order[i] : -1;
      s_row[tid] = r;
      s_len[tid] = (r >= 0) ? len[r] : 0;
      s_off[tid] = (r >= 0) ? off[r] : 0;
    }
    __syncthreads();
    int maxlen = 0;
    for (int i = 0; i < MT; ++i) maxlen = max(maxlen, s_len[i]);
    int rlen[4], roff[4], rrow[4];
    float dcc[4];      // dL/dc carried to the previous step
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      rlen[i] = worker ? s_len[4 * rg + i] : 0;
      roff[i] = worker ? s_off[4 * rg + i] : 0;
      rrow[i] = worker ? s_row[4 * rg + i] : -1;
      dcc[i] = 0.f;
    }
    for (int s = maxlen - 1; s >= 0; --s) {
      if (s < maxlen - 1) { lbar_wait_cluster(rfull, ph_full); ph_full ^= 1u; }   // partials of iteration s+1 landed
      if (worker) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          float dz[4];
          if (s < rlen[i]) {
            int t = dir ? (rlen[i] - 1 - s) : s;
            size_t p = (size_t)roff[i] + t;
            float* gp = gates + p * GS + (size_t)dir * 4 * HID + unit;
            float iv = gp[0], fv = gp[HID], gv = gp[2 * HID], ov = gp[3 * HID];
            float cv = c_stash[(p * 2 + dir) * HID + unit];
            float cpv = 0.f;
            if (s > 0) {
              size_t pp = dir ? p + 1 : p - 1;
              cpv = c_stash[(pp * 2 + dir) * HID + unit];
            }
            float dht = dh[p * 2 * HID + (size_t)dir * HID + unit];
            if (s == rlen[i] - 1) {   // the row's last step: start of its backward recursion
              dcc[i] = dcn[(size_t)rrow[i] * 2 * HID + (size_t)dir * HID + unit];
            } else {
              // recurrent part: sum of the CL partials in fixed order
              float r = 0.f;
#pragma unroll
              for (int src = 0; src < CL; ++src) r += recv[((size_t)src * UPC + j) * MT + 4 * rg + i];
              dht += r;
            }
            float tc = fast_tanh(cv);
            float dc = dcc[i] + dht * ov * (1.f - tc * tc);
            dz[3] = dht * tc * ov * (1.f - ov);
            dz[0] = dc * gv * iv * (1.f - iv);
            dz[2] = dc * iv * (1.f - gv * gv);
            dz[1] = dc * cpv * fv * (1.f - fv);
            dcc[i] = dc * fv;
#pragma unroll
            for (int g = 0; g < 4; ++g) gp[g * HID] = dz[g];
          } else {
#pragma unroll
            for (int g = 0; g < 4; ++g) dz[g] = 0.f;
          }
#pragma unroll
          for (int g = 0; g < 4; ++g) dzT[(size_t)(g * UPC + j) * MT + 4 * rg + i] = dz[g];
        }
      }
      __syncthreads();    // dzT complete; every thread of this CTA has consumed its recv values
      if (tid == 0 && s > 0) {
        lbar_expect_tx(rfull, (uint32_t)(CL * UPC * MT * sizeof(float)));           // arm for the partials of iteration s
#pragma unroll
        for (int d = 0; d < CL; ++d) rbar_arrive_release(remote_rfree[d]);           // "my recv may be overwritten"
      }
      float acc[4][4];
      if (worker2 && s > 0) {
        // packed pairs over rows (0,1) and (2,3): one FFMA2 per k and row pair
        float2 a2[4][2];
#pragma unroll
        for (int a = 0; a < 4; ++a) { a2[a][0] = make_float2(0.f, 0.f); a2[a][1] = make_float2(0.f, 0.f); }
        const float* wb = Wc + 4 * kg;
        const float* zb = dzT + 4 * rg2;
#pragma unroll 8
        for (int c = 0; c < COLS; ++c) {
          const float4 z = *reinterpret_cast<const float4*>(zb + (size_t)c * MT);
          const float4 w = *reinterpret_cast<const float4*>(wb + (size_t)c * HID);
          const float2 z01 = make_float2(z.x, z.y), z23 = make_float2(z.z, z.w);
          const float2 w0 = make_float2(w.x, w.x), w1 = make_float2(w.y, w.y), w2 = make_float2(w.z, w.z), w3 = make_float2(w.w, w.w);
          a2[0][0] = __ffma2_rn(w0, z01, a2[0][0]); a2[0][1] = __ffma2_rn(w0, z23, a2[0][1]);
          a2[1][0] = __ffma2_rn(w1, z01, a2[1][0]); a2[1][1] = __ffma2_rn(w1, z23, a2[1][1]);
          a2[2][0] = __ffma2_rn(w2, z01, a2[2][0]); a2[2][1] = __ffma2_rn(w2, z23, a2[2][1]);
          a2[3][0] = __ffma2_rn(w3, z01, a2[3][0]); a2[3][1] = __ffma2_rn(w3, z23, a2[3][1]);
        }
#pragma unroll
        for (int a = 0; a < 4; ++a) { acc[a][0] = a2[a][0].x; acc[a][1] = a2[a][0].y; acc[a][2] = a2[a][1].x; acc[a][3] = a2[a][1].y; }
      }
      if (s > 0) {
        lbar_wait_cluster(rfree, ph_free);   // every CTA has finished reading its recv -> safe to overwrite
        ph_free ^= 1u;
        if (worker2) {
#pragma unroll
          for (int a = 0; a < 4; ++a) {
            int k = 4 * kg + a;
            int owner = k / UPC, kl = k - owner * UPC;
            const uint32_t o = (uint32_t)((((size_t)rank * UPC + kl) * MT + 4 * rg2) * sizeof(float));
            st_async_f4(remote_recv[owner] + o, make_float4(acc[a][0], acc[a][1], acc[a][2], acc[a][3]), remote_rfull[owner]);
          }
        }
      }
      __syncthreads();    // dzT is rewritten by the next iteration's phase 1
    }
  }
  cluster_arrive();
  cluster_wait();
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
typedef LCfg<200, 4, 32, 2> FwdCfg200;     // 25 unit pairs x 8 row groups = 200 threads, double-buffered h
typedef LCfg<200, 4, 32, 1> BwdCfg200;     // 50 units x 8 row groups = 400 threads

template <class C, class K>
static int launch_cluster(K kernel, size_t smem, int ntiles, cudaStream_t st, void** args, const char* name) {
  static int max_clusters_cache = 0;
  static bool attr_set = false;
  if (!attr_set) {
    NNR_CUDA(cudaFuncSetAttribute((const void*)kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set = true;
  }
  cudaLaunchConfig_t cfg = {};
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = C::CL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  cfg.blockDim = dim3(C::NT); cfg.dynamicSmemBytes = smem; cfg.stream = st;
  if (max_clusters_cache == 0) {
    cfg.gridDim = dim3(C::CL * 2);
    int nc = 0;
    cudaError_t e = cudaOccupancyMaxActiveClusters(&nc, (const void*)kernel, &cfg);
    if (e != cudaSuccess || nc < 2) { (void)cudaGetLastError(); nc = 32; }
    max_clusters_cache = nc & ~1;
  }
  int want = 2 * ntiles;                      // (tile, direction) pairs
  int nclusters = want < max_clusters_cache ? want : max_clusters_cache;
  if (nclusters < 2) nclusters = 2;
  cfg.gridDim = dim3(nclusters * C::CL);
  cudaError_t e = cudaLaunchKernelExC(&cfg, (const void*)kernel, args);
  nnr_count_launch(1);
  if (e != cudaSuccess) { nnr_set_error("%s: launch failed: %s", name, cudaGetErrorString(e)); return (int)e; }
  return 0;
}

extern "C" int nnr_lstm_fwd(float* gx, const float* w_hh, const int32_t* len, const int32_t* off, const int32_t* order,
                            int N, int L, int H, float* h_out, float* c_stash, float* c_n, int32_t* tile_counters,
                            void* stream) {
  NNR_REQUIRE(gx && w_hh && len && off && order && h_out && c_stash && c_n && tile_counters && N > 0 && L > 0, NNR_ERR_ARG,
              "nnr_lstm_fwd: bad arguments");
  NNR_REQUIRE(H == 200, NNR_ERR_UNSUPPORTED, "nnr_lstm_fwd: hidden_dim %d not instantiated (200 only)", H);
  NNR_REQUIRE(nnr_aligned16(gx) && nnr_aligned16(h_out) && nnr_aligned16(c_stash) && nnr_aligned16(c_n), NNR_ERR_ALIGN,
              "nnr_lstm_fwd: buffers must be 16B aligned");
  typedef FwdCfg200 C;
  int ntiles = (N + C::MT - 1) / C::MT;
  NNR_CUDA(cudaMemsetAsync(tile_counters, 0, 2 * sizeof(int32_t), (cudaStream_t)stream));
  void* args[] = {&gx, &w_hh, &len, &off, &order, &N, &ntiles, &h_out, &c_stash, &c_n, &tile_counters};
  return launch_cluster<C>(lstm_fwd_kernel<C>, C::FWD_SMEM, ntiles, (cudaStream_t)stream, args, "lstm_fwd_kernel");
}

extern "C" int nnr_lstm_bwd(float* gates, const float* c_stash, const float* w_hh, const int32_t* len, const int32_t* off,
                            const int32_t* order, int N, int L, int H, const float* dh, const float* dcn,
                            int32_t* tile_counters, void* stream) {
  NNR_REQUIRE(gates && c_stash && w_hh && len && off && order && dh && dcn && tile_counters && N > 0 && L > 0, NNR_ERR_ARG,
              "nnr_lstm_bwd: bad arguments");
  NNR_REQUIRE(H == 200, NNR_ERR_UNSUPPORTED, "nnr_lstm_bwd: hidden_dim %d not instantiated (200 only)", H);
  NNR_REQUIRE(nnr_aligned16(gates) && nnr_aligned16(c_stash) && nnr_aligned16(dh) && nnr_aligned16(dcn), NNR_ERR_ALIGN,
              "nnr_lstm_bwd: buffers must be 16B aligned");
  typedef BwdCfg200 C;
  int ntiles = (N + C::MT - 1) / C::MT;
  NNR_CUDA(cudaMemsetAsync(tile_counters, 0, 2 * sizeof(int32_t), (cudaStream_t)stream));
  void* args[] = {&gates, &c_stash, &w_hh, &len, &off, &order, &N, &ntiles, &dh, &dcn, &tile_counters};
  return launch_cluster<C>(lstm_bwd_kernel<C>, C::BWD_SMEM, ntiles, (cudaStream_t)stream, args, "lstm_bwd_kernel");
}
