// Sequence bookkeeping, word-embedding gather and its deterministic scatter-add backward.
//
// Replaces (reference, PyTorch library calls): newsEncoders.py:106-111 (mask -> lengths),
// :117-118 (nn.Embedding + in-place nn.Dropout) and ATen's embedding_dense_backward.
//
// HBM-bound integer/byte work.  Gather: one warp per token row, 16-byte loads/stores
// (E = 300 floats = 75 float4), ids read once per token.  Backward: radix sort of the token
// slots by word id (CUB), then a two-level segment reduce whose summation order is a pure
// function of the sorted order -> bit-deterministic dense [V,E] gradient.
#include "common.cuh"
#include "../../include/nnr_b200.h"
#include <cub/cub.cuh>

// ------------------------------------------------------------------------------------------
// nnr_seq_prepare
// ------------------------------------------------------------------------------------------
__global__ void seq_len_kernel(uint8_t* __restrict__ mask, int N, int L, int32_t* __restrict__ len) {
  int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (warp >= N) return;
  uint8_t* m = mask + (size_t)warp * L;
  if (lane == 0) m[0] = 1;  // newsEncoders.py:108-109 "To avoid empty input of LSTM"
  __syncwarp();
  int c = 0;
  for (int t = lane; t < L; t += 32) c += (m[t] != 0) ? 1 : 0;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if (lane == 0) len[warp] = c;
}

// single-block exclusive scan (N up to a few hundred thousand rows; loops in chunks of 1024)
__global__ void seq_scan_kernel(const int32_t* __restrict__ len, int N, int32_t* __restrict__ off) {
  __shared__ int32_t warp_tot[32];
  __shared__ int32_t carry_s;
  int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  if (tid == 0) carry_s = 0;
  __syncthreads();
  for (int base = 0; base < N; base += 1024) {
    int i = base + tid;
    int v = (i < N) ? len[i] : 0;
    int x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int y = __shfl_up_sync(0xffffffffu, x, o);
      if (lane >= o) x += y;
    }
    if (lane == 31) warp_tot[w] = x;
    __syncthreads();
    if (w == 0) {
      int t = warp_tot[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        int y = __shfl_up_sync(0xffffffffu, t, o);
        if (lane >= o) t += y;
      }
      warp_tot[lane] = t;
    }
    __syncthreads();
    int carry = carry_s;
    int incl = x + (w > 0 ? warp_tot[w - 1] : 0) + carry;
    if (i < N) off[i] = incl - v;
    __syncthreads();
    if (tid == 1023) carry_s = incl;
    __syncthreads();
  }
  if (tid == 0) off[N] = carry_s;
}

__global__ void seq_tokrow_kernel(const int32_t* __restrict__ len, const int32_t* __restrict__ off, int N,
                                  int32_t* __restrict__ tok_row) {
  int r = blockIdx.x;
  if (r >= N) return;
  int o = off[r], l = len[r];
  for (int t = threadIdx.x; t < l; t += blockDim.x) tok_row[o + t] = r;
}

extern "C" int nnr_seq_prepare(uint8_t* mask, int N, int L, int32_t* len, int32_t* off, int32_t* tok_row,
                               void* stream) {
  NNR_REQUIRE(mask && len && off && N > 0 && L > 0, NNR_ERR_ARG, "nnr_seq_prepare: bad arguments (N=%d L=%d)", N, L);
  cudaStream_t st = (cudaStream_t)stream;
  int warps_per_block = 8;
  seq_len_kernel<<<(N + warps_per_block - 1) / warps_per_block, warps_per_block * 32, 0, st>>>(mask, N, L, len);
  NNR_LAUNCH_CHECK("seq_len_kernel");
  seq_scan_kernel<<<1, 1024, 0, st>>>(len, N, off);
  NNR_LAUNCH_CHECK("seq_scan_kernel");
  if (tok_row) {
    seq_tokrow_kernel<<<N, 32, 0, st>>>(len, off, N, tok_row);
    NNR_LAUNCH_CHECK("seq_tokrow_kernel");
  }
  return 0;
}

// ------------------------------------------------------------------------------------------
// nnr_embed_gather_fwd : one warp per (row, t) token slot; packed output
// ------------------------------------------------------------------------------------------
template <bool VEC>
__global__ void embed_gather_kernel(const float* __restrict__ table, const int32_t* __restrict__ ids,
                                    const int32_t* __restrict__ len, const int32_t* __restrict__ off, int N, int L,
                                    int E, int V, float* __restrict__ out, float p, float inv_keep, uint64_t seed) {
  int slot = (int)(((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  int lane = threadIdx.x & 31;
  if (slot >= N * L) return;
  int r = slot / L, t = slot - r * L;
  if (t >= len[r]) return;
  if (p > 0.0f) seed = nnr_resolve_seed(seed);
  int id = ids[slot];
  id = min(max(id, 0), V - 1);
  const float* src = table + (size_t)id * E;
  float* dst = out + ((size_t)off[r] + t) * E;
  uint64_t ebase = (uint64_t)slot * (uint64_t)E;
  if (VEC) {
    int E4 = E >> 2;
    for (int q = lane; q < E4; q += 32) {
      float4 v = __ldg(reinterpret_cast<const float4*>(src) + q);
      if (p > 0.0f) {                 // E % 4 == 0 on this path, so ebase is a multiple of 4
        float ks[4];
        dropout_scale4(seed, (ebase >> 2) + q, p, inv_keep, ks);
        v.x *= ks[0]; v.y *= ks[1]; v.z *= ks[2]; v.w *= ks[3];
      }
      reinterpret_cast<float4*>(dst)[q] = v;
    }
  } else {
    for (int e = lane; e < E; e += 32) dst[e] = __ldg(src + e) * dropout_scale_e4(seed, ebase + e, p, inv_keep);
  }
}

extern "C" int nnr_embed_gather_fwd(const float* table, const int32_t* ids, const int32_t* len, const int32_t* off,
                                    int N, int L, int E, int V, float* out, float p_drop, uint64_t seed,
                                    void* stream) {
  NNR_REQUIRE(table && ids && len && off && out && N > 0 && L > 0 && E > 0 && V > 0, NNR_ERR_ARG,
              "nnr_embed_gather_fwd: bad arguments");
  NNR_REQUIRE(p_drop >= 0.0f && p_drop < 1.0f, NNR_ERR_ARG, "nnr_embed_gather_fwd: p_drop=%f", p_drop);
  cudaStream_t st = (cudaStream_t)stream;
  size_t slots = (size_t)N * L;
  int threads = 256;
  size_t blocks = (slots * 32 + threads - 1) / threads;
  float inv_keep = 1.0f / (1.0f - p_drop);
  bool vec = (E % 4 == 0) && nnr_aligned16(table) && nnr_aligned16(out);
  if (vec)
    embed_gather_kernel<true><<<(unsigned)blocks, threads, 0, st>>>(table, ids, len, off, N, L, E, V, out, p_drop, inv_keep, seed);
  else
    embed_gather_kernel<false><<<(unsigned)blocks, threads, 0, st>>>(table, ids, len, off, N, L, E, V, out, p_drop, inv_keep, seed);
  NNR_LAUNCH_CHECK("embed_gather_kernel");
  return 0;
}

// ------------------------------------------------------------------------------------------
// nnr_embed_gather_bwd : sort slots by id, then chunked segment reduce (deterministic)
// ------------------------------------------------------------------------------------------
#ifndef EB_CHUNK
#define EB_CHUNK 128   // sorted slots per CTA = threads per CTA: small chunks = more CTAs in flight (the per-run work is latency
                       // bound; only the first ~18 % of the sorted slots are valid tokens).  scripts/sweep_eb_chunk.sh: 256 -> 0.424 ms,
                       // 128 / 64 -> 0.358 ms per step
#endif
#define EB_WARPS (EB_CHUNK / 32)
struct EbMeta { int32_t head_cross, tail_cross, covering, head_key, tail_key, pad0, pad1, pad2; };

__global__ void eb_keys_kernel(const int32_t* __restrict__ ids, const int32_t* __restrict__ len, int N, int L, int V,
                               int32_t* __restrict__ keys, int32_t* __restrict__ vals) {
  int slot = blockIdx.x * blockDim.x + threadIdx.x;
  if (slot >= N * L) return;
  int r = slot / L, t = slot - r * L;
  int id = ids[slot];
  id = min(max(id, 0), V - 1);
  keys[slot] = (t < len[r]) ? id : V;  // invalid slots sort to the end
  vals[slot] = slot;
}

// Sum the dout rows of sorted entries first, first+stride, ... (< end) of a chunk into acc (one warp; lane l
// owns float4 columns l, l+32, l+64).  Slot ids / row offsets are fetched 32 at a time by the lanes and
// broadcast with shuffles; rows are loaded four at a time so 12 independent 16-byte loads are in flight per
// lane, and added in entry order (the summation order is a pure function of the sorted order).
__device__ __forceinline__ void eb_warp_accumulate(float4 (&acc)[3], const float* __restrict__ dout,
                                                   const int32_t* __restrict__ s_slot, const unsigned long long* __restrict__ s_src,
                                                   int E, int E4, int first, int end, int stride, int lane, float p,
                                                   float inv_keep, uint64_t seed) {
  // s_slot / s_src: token slot and source offset of every sorted entry of the chunk, resolved once per CTA (the two
  // dependent global loads per entry used to sit in front of every run's row loads)
  for (int base = first; base < end; base += 32 * stride) {
    const int e = base + lane * stride;
    int slot_l = 0;
    unsigned long long src_l = 0;
    if (e < end) {
      slot_l = s_slot[e];
      src_l = s_src[e];
    }
    const int gn = min(32, (end - base + stride - 1) / stride);
    int j = 0;
    for (; j + 4 <= gn; j += 4) {
      float4 v[4][3];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int slot = __shfl_sync(0xffffffffu, slot_l, j + u);
        const unsigned long long so = __shfl_sync(0xffffffffu, src_l, j + u);
        const float4* src = reinterpret_cast<const float4*>(dout + so);
        const uint64_t ebase = (uint64_t)slot * (uint64_t)E;
#pragma unroll
        for (int q = 0; q < 3; ++q) {
          const int c4 = lane + 32 * q;
          v[u][q] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (c4 < E4) {
            float4 x = __ldg(src + c4);
            if (p > 0.0f) {
              float ks[4];
              dropout_scale4(seed, (ebase >> 2) + c4, p, inv_keep, ks);
              x.x *= ks[0]; x.y *= ks[1]; x.z *= ks[2]; x.w *= ks[3];
            }
            v[u][q] = x;
          }
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int q = 0; q < 3; ++q) { acc[q].x += v[u][q].x; acc[q].y += v[u][q].y; acc[q].z += v[u][q].z; acc[q].w += v[u][q].w; }
    }
    for (; j < gn; ++j) {
      const int slot = __shfl_sync(0xffffffffu, slot_l, j);
      const unsigned long long so = __shfl_sync(0xffffffffu, src_l, j);
      const float4* src = reinterpret_cast<const float4*>(dout + so);
      const uint64_t ebase = (uint64_t)slot * (uint64_t)E;
#pragma unroll
      for (int q = 0; q < 3; ++q) {
        const int c4 = lane + 32 * q;
        if (c4 < E4) {
          float4 x = __ldg(src + c4);
          if (p > 0.0f) {
            float ks[4];
            dropout_scale4(seed, (ebase >> 2) + c4, p, inv_keep, ks);
            x.x *= ks[0]; x.y *= ks[1]; x.z *= ks[2]; x.w *= ks[3];
          }
          acc[q].x += x.x; acc[q].y += x.y; acc[q].z += x.z; acc[q].w += x.w;
        }
      }
    }
  }
}

#define EB_LONG 48   // runs at least this long are summed by all warps of the CTA

// one CTA per chunk of EB_CHUNK sorted slots.  Short runs of equal keys: one warp each.  Long runs (frequent
// words, the <PAD> token): the 8 warps take every 8th entry and their partial sums are combined in warp order.
__global__ void __launch_bounds__(EB_CHUNK) eb_chunk_kernel(const float* __restrict__ dout, const int32_t* __restrict__ keys,
                                                       const int32_t* __restrict__ vals,
                                                       const int32_t* __restrict__ off, int N, int L, int E, float p,
                                                       float inv_keep, uint64_t seed, float* __restrict__ dtable,
                                                       int accumulate, float* __restrict__ slots_ws,
                                                       EbMeta* __restrict__ meta) {
  __shared__ int32_t s_key[EB_CHUNK + 2];
  __shared__ int32_t s_start[EB_CHUNK + 1];
  __shared__ int32_t s_wcnt[EB_WARPS];
  __shared__ int32_t s_nruns;
  __shared__ __align__(16) float s_part[EB_WARPS][384];
  const int n_valid = off[N];
  const int chunk = blockIdx.x;
  const int cs = chunk * EB_CHUNK;
  if (cs >= n_valid) return;
  if (p > 0.0f) seed = nnr_resolve_seed(seed);
  const int ce = min(cs + EB_CHUNK, n_valid);
  const int cnt = ce - cs;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  __shared__ int32_t s_slot[EB_CHUNK];
  __shared__ unsigned long long s_src[EB_CHUNK];
  if (tid < cnt) {                                   // sorted entry -> (token slot, offset of its dL/dout row)
    const int slot = vals[cs + tid];
    const int row = slot / L, t = slot - row * L;
    s_slot[tid] = slot;
    s_src[tid] = ((unsigned long long)off[row] + t) * (unsigned long long)E;
  }
  // keys of [cs-1, ce]  (s_key[i+1] = key[cs+i])
  for (int i = tid; i < cnt + 2; i += EB_CHUNK) {
    int pos = cs - 1 + i;
    s_key[i] = (pos >= 0 && pos < n_valid) ? keys[pos] : -1;
  }
  __syncthreads();
  // run-start flags -> compact list (each warp owns 32 consecutive entries)
  static_assert(EB_CHUNK % 32 == 0 && EB_CHUNK >= 32 && EB_CHUNK <= 1024, "one entry per thread");
  int f0 = 0;
  const int i0 = tid;
  if (i0 < cnt) f0 = (i0 == 0) || (s_key[i0 + 1] != s_key[i0]);
  const unsigned b0 = __ballot_sync(0xffffffffu, f0);
  if (lane == 0) s_wcnt[w] = __popc(b0);
  __syncthreads();
  int wbase = 0;
  for (int j = 0; j < w; ++j) wbase += s_wcnt[j];
  if (f0) s_start[wbase + __popc(b0 & ((1u << lane) - 1))] = i0;
  if (tid == 0) {
    int tot = 0;
    for (int j = 0; j < EB_WARPS; ++j) tot += s_wcnt[j];
    s_nruns = tot;
    s_start[tot] = cnt;
  }
  __syncthreads();
  const int nruns = s_nruns;
  const bool head_cross = (cs > 0) && (s_key[0] == s_key[1]);
  const bool tail_cross = (ce < n_valid) && (s_key[cnt + 1] == s_key[cnt]);
  if (tid == 0) {
    EbMeta m;
    m.head_cross = head_cross; m.tail_cross = tail_cross;
    m.covering = (nruns == 1) && head_cross && tail_cross ? 1 : ((nruns == 1) && tail_cross && !head_cross ? 2 : 0);
    m.head_key = s_key[1]; m.tail_key = s_key[cnt];
    m.pad0 = m.pad1 = m.pad2 = 0;
    meta[chunk] = m;
  }
  const int E4 = E >> 2;  // host guarantees E % 4 == 0
  // destination of run r: the table row, or a boundary slot when the run crosses the chunk edge
  auto run_dst = [&](int r, bool& add) -> float* {
    const bool hc = (r == 0) && head_cross;
    const bool tc = (r == nruns - 1) && tail_cross;
    add = false;
    if (!hc && !tc) { add = accumulate != 0; return dtable + (size_t)s_key[s_start[r] + 1] * E; }
    if (hc) return slots_ws + (size_t)(2 * chunk) * E;          // head (or covering) piece
    return slots_ws + (size_t)(2 * chunk + 1) * E;              // tail piece
  };
  // phase A: short runs, one warp each
  int ri = 0;
  for (int r = 0; r < nruns; ++r) {
    const int a = s_start[r], b = s_start[r + 1];
    if (b - a >= EB_LONG) continue;
    if ((ri++ % EB_WARPS) != w) continue;
    float4 acc[3];
#pragma unroll
    for (int q = 0; q < 3; ++q) acc[q] = make_float4(0.f, 0.f, 0.f, 0.f);
    eb_warp_accumulate(acc, dout, s_slot, s_src, E, E4, a, b, 1, lane, p, inv_keep, seed);
    bool add;
    float* dst = run_dst(r, add);
#pragma unroll
    for (int q = 0; q < 3; ++q) {
      int c4 = lane + 32 * q;
      if (c4 < E4) {
        float4* d4 = reinterpret_cast<float4*>(dst) + c4;
        float4 v = acc[q];
        if (add) { float4 o = *d4; v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w; }
        *d4 = v;
      }
    }
  }
  // phase B: long runs, all warps cooperate (uniform control flow: every thread sees the same run list)
  for (int r = 0; r < nruns; ++r) {
    const int a = s_start[r], b = s_start[r + 1];
    if (b - a < EB_LONG) continue;
    float4 acc[3];
#pragma unroll
    for (int q = 0; q < 3; ++q) acc[q] = make_float4(0.f, 0.f, 0.f, 0.f);
    eb_warp_accumulate(acc, dout, s_slot, s_src, E, E4, a + w, b, EB_WARPS, lane, p, inv_keep, seed);
#pragma unroll
    for (int q = 0; q < 3; ++q) {
      int c4 = lane + 32 * q;
      if (c4 < E4) reinterpret_cast<float4*>(s_part[w])[c4] = acc[q];
    }
    __syncthreads();
    bool add;
    float* dst = run_dst(r, add);
    for (int e = tid; e < E; e += EB_CHUNK) {
      float v = 0.f;
#pragma unroll
      for (int ww = 0; ww < EB_WARPS; ++ww) v += s_part[ww][e];
      dst[e] = add ? dst[e] + v : v;
    }
    __syncthreads();
  }
}

// one CTA per chunk whose tail piece BEGINS a boundary-crossing run: add the following pieces in order
__global__ void eb_fix_kernel(const int32_t* __restrict__ off, int N, int E, const float* __restrict__ slots_ws,
                              const EbMeta* __restrict__ meta, float* __restrict__ dtable, int accumulate) {
  const int n_valid = off[N];
  const int j = blockIdx.x;
  if (j * EB_CHUNK >= n_valid) return;
  const EbMeta m = meta[j];
  if (!m.tail_cross) return;
  if (m.covering == 1) return;  // middle of a run that started in an earlier chunk
  const int key = m.tail_key;
  // the run continues through every following chunk it covers entirely and ends in the head piece of the next one:
  // find that chunk first (warp-parallel scan of the flags), then add the pieces in order with independent loads
  __shared__ int s_end;
  if (threadIdx.x < 32) {
    int k = j + 1;
    const int nchunks = (n_valid + EB_CHUNK - 1) / EB_CHUNK;
    while (true) {
      const int kk = k + (int)threadIdx.x;
      const bool cover = (kk < nchunks) && (meta[kk].covering == 1);
      const unsigned m32 = __ballot_sync(0xffffffffu, cover);
      if (m32 != 0xffffffffu) { k += __ffs(~m32) - 1; break; }
      k += 32;
    }
    if (threadIdx.x == 0) s_end = k;           // last piece = head slot of chunk k
  }
  __syncthreads();
  const int kend = s_end;
  for (int e = threadIdx.x; e < E; e += blockDim.x) {
    float acc = slots_ws[(size_t)(2 * j + 1) * E + e];
    int k = j + 1;
    for (; k + 4 <= kend + 1; k += 4) {
      const float v0 = slots_ws[(size_t)(2 * k) * E + e], v1 = slots_ws[(size_t)(2 * k + 2) * E + e];
      const float v2 = slots_ws[(size_t)(2 * k + 4) * E + e], v3 = slots_ws[(size_t)(2 * k + 6) * E + e];
      acc += v0; acc += v1; acc += v2; acc += v3;
    }
    for (; k <= kend; ++k) acc += slots_ws[(size_t)(2 * k) * E + e];
    float* d = dtable + (size_t)key * E + e;
    *d = accumulate ? (*d + acc) : acc;
  }
}

static size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

struct EbLayout { size_t keys_in, keys_out, vals_in, vals_out, slots, meta, cub, total, cub_bytes; };
static EbLayout eb_layout(int N, int L, int E) {
  EbLayout l;
  size_t n = (size_t)N * L;
  size_t nchunks = (n + EB_CHUNK - 1) / EB_CHUNK;
  size_t o = 0;
  l.keys_in = o; o = align_up(o + n * 4, 256);
  l.keys_out = o; o = align_up(o + n * 4, 256);
  l.vals_in = o; o = align_up(o + n * 4, 256);
  l.vals_out = o; o = align_up(o + n * 4, 256);
  l.slots = o; o = align_up(o + 2 * nchunks * (size_t)E * 4, 256);
  l.meta = o; o = align_up(o + nchunks * sizeof(EbMeta), 256);
  size_t cub_bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, cub_bytes, (const int32_t*)nullptr, (int32_t*)nullptr,
                                  (const int32_t*)nullptr, (int32_t*)nullptr, (int)n, 0, 32);
  l.cub = o; l.cub_bytes = cub_bytes; o = align_up(o + cub_bytes, 256);
  l.total = o;
  return l;
}

extern "C" size_t nnr_embed_gather_bwd_workspace_bytes(int N, int L) {
  if (N <= 0 || L <= 0) return 0;
  return eb_layout(N, L, 512).total;  // E-independent upper bound for E <= 512
}

extern "C" int nnr_embed_gather_bwd(const float* dout, const int32_t* ids, const int32_t* len, const int32_t* off,
                                    int N, int L, int E, int V, float p_drop, uint64_t seed, float* dtable,
                                    int accumulate, void* workspace, size_t workspace_bytes, void* stream) {
  NNR_REQUIRE(dout && ids && len && off && dtable && workspace && N > 0 && L > 0 && V > 0, NNR_ERR_ARG,
              "nnr_embed_gather_bwd: bad arguments");
  NNR_REQUIRE(E % 4 == 0 && E <= 384 && nnr_aligned16(dout) && nnr_aligned16(dtable) && nnr_aligned16(workspace),
              NNR_ERR_ALIGN, "nnr_embed_gather_bwd: E=%d must be a multiple of 4 (<=384), pointers 16B aligned", E);
  EbLayout l = eb_layout(N, L, E);
  NNR_REQUIRE(workspace_bytes >= l.total, NNR_ERR_WORKSPACE, "nnr_embed_gather_bwd: workspace %zu < %zu",
              workspace_bytes, l.total);
  cudaStream_t st = (cudaStream_t)stream;
  char* ws = (char*)workspace;
  int32_t* keys_in = (int32_t*)(ws + l.keys_in);
  int32_t* keys_out = (int32_t*)(ws + l.keys_out);
  int32_t* vals_in = (int32_t*)(ws + l.vals_in);
  int32_t* vals_out = (int32_t*)(ws + l.vals_out);
  float* slots = (float*)(ws + l.slots);
  EbMeta* meta = (EbMeta*)(ws + l.meta);
  int n = N * L;
  if (!accumulate) NNR_CUDA(cudaMemsetAsync(dtable, 0, (size_t)V * E * sizeof(float), st));
  eb_keys_kernel<<<(n + 255) / 256, 256, 0, st>>>(ids, len, N, L, V, keys_in, vals_in);
  NNR_LAUNCH_CHECK("eb_keys_kernel");
  int end_bit = 1;
  while ((1 << end_bit) <= V) ++end_bit;  // keys are in [0, V]
  size_t cub_bytes = l.cub_bytes;
  NNR_CUDA(cub::DeviceRadixSort::SortPairs(ws + l.cub, cub_bytes, keys_in, keys_out, vals_in, vals_out, n, 0,
                                           end_bit, st));
  nnr_count_launch(3);
  int nchunks = (n + EB_CHUNK - 1) / EB_CHUNK;
  float inv_keep = 1.0f / (1.0f - p_drop);
  eb_chunk_kernel<<<nchunks, EB_CHUNK, 0, st>>>(dout, keys_out, vals_out, off, N, L, E, p_drop, inv_keep, seed, dtable,
                                           accumulate, slots, meta);
  NNR_LAUNCH_CHECK("eb_chunk_kernel");
  eb_fix_kernel<<<nchunks, 128, 0, st>>>(off, N, E, slots, meta, dtable, accumulate);
  NNR_LAUNCH_CHECK("eb_fix_kernel");
  return 0;
}
