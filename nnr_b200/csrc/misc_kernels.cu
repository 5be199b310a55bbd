// Small HBM-bound kernels: column sums, segment sums, LSTM hidden-state shift, gate backward
// prologue, news-vector assembly, row dot products, dropout, fused clip+Adam.
#include "common.cuh"
#include "../../include/nnr_b200.h"

// ------------------------------------------------------------------------------------------
// column sums (bias gradients): two-stage, fixed partition -> deterministic
// ------------------------------------------------------------------------------------------
#define CS_ROWS 64
__global__ void colsum_stage1(const float* __restrict__ X, int64_t ldx, int M, int N, const int32_t* __restrict__ m_dev,
                              float* __restrict__ part) {
  int Me = m_dev ? min(M, *m_dev) : M;
  int n = blockIdx.x * blockDim.x + threadIdx.x;
  int r0 = blockIdx.y * CS_ROWS;
  if (n >= N) return;
  float acc = 0.f;
  int r1 = min(r0 + CS_ROWS, Me);
  int r = r0;
  for (; r + 8 <= r1; r += 8) {             // eight rows in flight, added in row order
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = X[(size_t)(r + j) * ldx + n];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc += v[j];
  }
  for (; r < r1; ++r) acc += X[(size_t)r * ldx + n];
  part[(size_t)blockIdx.y * N + n] = acc;
}
__global__ void colsum_stage2(const float* __restrict__ part, int nparts, int N, int M, const int32_t* __restrict__ m_dev,
                              float* __restrict__ out, int accumulate) {
  int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  if (m_dev) nparts = min(nparts, (min(M, *m_dev) + CS_ROWS - 1) / CS_ROWS);   // partials beyond the valid rows are zero
  float acc = 0.f;
  for (int p = 0; p < nparts; ++p) acc += part[(size_t)p * N + n];
  out[n] = accumulate ? out[n] + acc : acc;
}
extern "C" size_t nnr_colsum_workspace_bytes(int M, int N) {
  if (M <= 0 || N <= 0) return 0;
  return (size_t)((M + CS_ROWS - 1) / CS_ROWS) * N * sizeof(float);
}
extern "C" int nnr_colsum(const float* X, int64_t ldx, int M, int N, const int32_t* m_dev, float* out, int accumulate,
                          void* workspace, size_t workspace_bytes, void* stream) {
  NNR_REQUIRE(X && out && workspace && M > 0 && N > 0, NNR_ERR_ARG, "nnr_colsum: bad arguments");
  NNR_REQUIRE(workspace_bytes >= nnr_colsum_workspace_bytes(M, N), NNR_ERR_WORKSPACE, "nnr_colsum: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  int nparts = (M + CS_ROWS - 1) / CS_ROWS;
  dim3 g1((N + 127) / 128, nparts);
  colsum_stage1<<<g1, 128, 0, st>>>(X, ldx, M, N, m_dev, (float*)workspace);
  NNR_LAUNCH_CHECK("colsum_stage1");
  colsum_stage2<<<(N + 127) / 128, 128, 0, st>>>((const float*)workspace, nparts, N, M, m_dev, out, accumulate);
  NNR_LAUNCH_CHECK("colsum_stage2");
  return 0;
}

// out[r,:] = sum_t X[off[r]+t,:]
__global__ void segment_colsum_kernel(const float* __restrict__ X, int64_t ldx, const int32_t* __restrict__ off, int D,
                                      float* __restrict__ out, int64_t ldo) {
  int r = blockIdx.x;
  int a = off[r], b = off[r + 1];
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    float acc = 0.f;
    for (int p = a; p < b; ++p) acc += X[(size_t)p * ldx + d];
    out[(size_t)r * ldo + d] = acc;
  }
}
extern "C" int nnr_segment_colsum(const float* X, int64_t ldx, const int32_t* off, int N, int D, float* out, int64_t ldo,
                                  void* stream) {
  NNR_REQUIRE(X && off && out && N > 0 && D > 0, NNR_ERR_ARG, "nnr_segment_colsum: bad arguments");
  segment_colsum_kernel<<<N, 128, 0, (cudaStream_t)stream>>>(X, ldx, off, D, out, ldo);
  NNR_LAUNCH_CHECK("segment_colsum_kernel");
  return 0;
}

// ------------------------------------------------------------------------------------------
// hprev for dW_hh
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) lstm_shift_h_kernel(const float* __restrict__ h, const int32_t* __restrict__ len,
                                                           const int32_t* __restrict__ off, const int32_t* __restrict__ tok_row,
                                                           int N, int H, float* __restrict__ hprev) {
  // grid-stride over (token, float4 column group); the token count is read on the device
  const int ntok = off[N];
  const int H2 = 2 * H, q_per_tok = H2 >> 2, hq = H >> 2;
  const long long total = (long long)ntok * q_per_tok;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int p = (int)(i / q_per_tok), q = (int)(i - (long long)p * q_per_tok);
    const int r = tok_row[p];
    const int t = p - off[r];
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (q < hq) { if (t > 0) v = __ldg(reinterpret_cast<const float4*>(h + (size_t)(p - 1) * H2) + q); }
    else { if (t < len[r] - 1) v = __ldg(reinterpret_cast<const float4*>(h + (size_t)(p + 1) * H2) + q); }
    reinterpret_cast<float4*>(hprev + (size_t)p * H2)[q] = v;
  }
}
extern "C" int nnr_lstm_shift_h(const float* h, const int32_t* len, const int32_t* off, const int32_t* tok_row, int N,
                                int L, int H, float* hprev, void* stream) {
  NNR_REQUIRE(h && len && off && tok_row && hprev && N > 0 && L > 0 && H > 0, NNR_ERR_ARG, "nnr_lstm_shift_h: bad arguments");
  NNR_REQUIRE(H % 4 == 0 && nnr_aligned16(h) && nnr_aligned16(hprev), NNR_ERR_ALIGN, "nnr_lstm_shift_h: H %% 4 == 0, 16B alignment");
  lstm_shift_h_kernel<<<148 * 8, 256, 0, (cudaStream_t)stream>>>(h, len, off, tok_row, N, H, hprev);
  NNR_LAUNCH_CHECK("lstm_shift_h_kernel");
  return 0;
}

// ------------------------------------------------------------------------------------------
// selective gate backward prologue
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) gate_bwd_pre_kernel(const float* __restrict__ dhg, const float* __restrict__ h,
                                                           const float* __restrict__ g, int64_t n_max,
                                                           const int32_t* __restrict__ n_dev, int D, float* __restrict__ dz,
                                                           float* __restrict__ dh0) {
  const int64_t n = n_dev ? min(n_max, (int64_t)(*n_dev) * D) : n_max;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x * 4;
  for (int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < n; i += stride) {
    float4 a = *reinterpret_cast<const float4*>(dhg + i);
    float4 hh = *reinterpret_cast<const float4*>(h + i);
    float4 gg = *reinterpret_cast<const float4*>(g + i);
    float4 z, d0;
    z.x = a.x * hh.x * gg.x * (1.f - gg.x); d0.x = a.x * gg.x;
    z.y = a.y * hh.y * gg.y * (1.f - gg.y); d0.y = a.y * gg.y;
    z.z = a.z * hh.z * gg.z * (1.f - gg.z); d0.z = a.z * gg.z;
    z.w = a.w * hh.w * gg.w * (1.f - gg.w); d0.w = a.w * gg.w;
    *reinterpret_cast<float4*>(dz + i) = z;
    *reinterpret_cast<float4*>(dh0 + i) = d0;
  }
}
extern "C" int nnr_gate_bwd_pre(const float* dhg, const float* h, const float* g, int64_t n_max, const int32_t* n_dev,
                                int D, float* dz, float* dh0, void* stream) {
  NNR_REQUIRE(dhg && h && g && dz && dh0 && n_max > 0 && D > 0, NNR_ERR_ARG, "nnr_gate_bwd_pre: bad arguments");
  NNR_REQUIRE(n_max % 4 == 0 && D % 4 == 0 && nnr_aligned16(dhg) && nnr_aligned16(h) && nnr_aligned16(g) &&
                  nnr_aligned16(dz) && nnr_aligned16(dh0),
              NNR_ERR_ALIGN, "nnr_gate_bwd_pre: needs 16B alignment and D %% 4 == 0");
  gate_bwd_pre_kernel<<<148 * 8, 256, 0, (cudaStream_t)stream>>>(dhg, h, g, n_max, n_dev, D, dz, dh0);
  NNR_LAUNCH_CHECK("gate_bwd_pre_kernel");
  return 0;
}

// ------------------------------------------------------------------------------------------
// news vector assembly (newsEncoders.py:50-54,138)
// ------------------------------------------------------------------------------------------
__global__ void news_fuse_fwd_kernel(const float* __restrict__ ts, const float* __restrict__ tc, const float* __restrict__ cs,
                                     const float* __restrict__ cc, const float* __restrict__ cat_table,
                                     const float* __restrict__ sub_table, const int32_t* __restrict__ cat,
                                     const int32_t* __restrict__ sub, int D2, int Ec, int Es, float p, float inv_keep,
                                     uint64_t seed, float* __restrict__ out) {
  int r = blockIdx.x;
  if (p > 0.f) seed = nnr_resolve_seed(seed);
  const int DM = cs ? 2 * D2 : D2;                  // one or two modalities (CNE_Title / CNE_Content pass c_self = NULL)
  int Dout = DM + Ec + Es;
  float* o = out + (size_t)r * Dout;
  for (int d = threadIdx.x; d < Dout; d += blockDim.x) {
    float v;
    if (d < D2) v = ts[(size_t)r * D2 + d] + (tc ? tc[(size_t)r * D2 + d] : 0.f);
    else if (d < DM) v = cs[(size_t)r * D2 + d - D2] + (cc ? cc[(size_t)r * D2 + d - D2] : 0.f);
    else if (d < DM + Ec) {
      int e = d - DM;
      v = cat_table[(size_t)cat[r] * Ec + e] * dropout_scale(seed, (uint64_t)r * (Ec + Es) + e, p, inv_keep);
    } else {
      int e = d - DM - Ec;
      v = sub_table[(size_t)sub[r] * Es + e] * dropout_scale(seed, (uint64_t)r * (Ec + Es) + Ec + e, p, inv_keep);
    }
    o[d] = v;
  }
}
extern "C" int nnr_news_fuse_fwd(const float* t_self, const float* t_cross, const float* c_self, const float* c_cross,
                                 const float* cat_table, const float* sub_table, const int32_t* cat, const int32_t* sub,
                                 int N, int D2, int Ec, int Es, float p_drop, uint64_t seed, float* out, void* stream) {
  NNR_REQUIRE(t_self && cat_table && sub_table && cat && sub && out && N > 0, NNR_ERR_ARG,
              "nnr_news_fuse_fwd: bad arguments");
  float inv_keep = 1.0f / (1.0f - p_drop);
  news_fuse_fwd_kernel<<<N, 256, 0, (cudaStream_t)stream>>>(t_self, t_cross, c_self, c_cross, cat_table, sub_table, cat,
                                                           sub, D2, Ec, Es, p_drop, inv_keep, seed, out);
  NNR_LAUNCH_CHECK("news_fuse_fwd_kernel");
  return 0;
}

__global__ void news_fuse_split_kernel(const float* __restrict__ dout, int D2, int Dout, float* __restrict__ d_a,
                                       float* __restrict__ d_b) {
  int r = blockIdx.x;
  for (int d = threadIdx.x; d < (d_b ? 2 * D2 : D2); d += blockDim.x) {
    float v = dout[(size_t)r * Dout + d];
    if (d < D2) d_a[(size_t)r * D2 + d] = v;
    else d_b[(size_t)r * D2 + d - D2] = v;
  }
}
// one block per table row, 32 warps.  Warp w owns news rows [w*span, (w+1)*span): a category that matches a large share
// of the rows (e.g. the padding category of the history) costs pipelined loads rather than one exposed latency per match,
// and rows that do not match cost one vote per 32.
// Lane = embedding column; rows are accumulated in order and the 32 per-warp partials are combined in warp order:
// deterministic, no atomics.
#define NF_MAXE 64
#define NF_WARPS 32
__global__ void __launch_bounds__(NF_WARPS * 32) news_fuse_table_bwd_kernel(const float* __restrict__ dout, const int32_t* __restrict__ idx,
                                                                          int N, int Dout, int col0, int Edim, int Etot, int eoff,
                                                                          float p, float inv_keep, uint64_t seed,
                                                                          float* __restrict__ dtable, int accumulate) {
  __shared__ float s_part[NF_WARPS][NF_MAXE];
  if (p > 0.f) seed = nnr_resolve_seed(seed);
  const int row = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int span = (N + NF_WARPS - 1) / NF_WARPS;
  const int r_begin = w * span, r_end = min(N, r_begin + span);
  const bool c0 = lane < Edim, c1 = lane + 32 < Edim;
  float acc0 = 0.f, acc1 = 0.f;                       // columns lane and lane + 32
  // the lanes fetch 32 indices at a time (coalesced) and vote; only the matching rows are visited, four at a time with
  // their row loads issued together, in ascending row order
  for (int rb = r_begin; rb < r_end; rb += 32) {
    const int rl = rb + lane;
    unsigned m = __ballot_sync(0xffffffffu, rl < r_end && __ldg(idx + rl) == row);
    while (m) {
      int rr[4];
      float v0[4], v1[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        rr[j] = -1;
        if (m) { const int b = __ffs(m) - 1; m &= m - 1; rr[j] = rb + b; }
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        v0[j] = 0.f; v1[j] = 0.f;
        if (rr[j] >= 0) {
          const float* src = dout + (size_t)rr[j] * Dout + col0;
          if (c0) v0[j] = src[lane];
          if (c1) v1[j] = src[lane + 32];
        }
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (rr[j] >= 0) {
          const uint64_t base = (uint64_t)rr[j] * Etot + eoff;
          if (c0) acc0 += v0[j] * dropout_scale(seed, base + lane, p, inv_keep);
          if (c1) acc1 += v1[j] * dropout_scale(seed, base + lane + 32, p, inv_keep);
        }
      }
    }
  }
  s_part[w][lane] = acc0;
  s_part[w][lane + 32] = acc1;
  __syncthreads();
  if (tid < Edim) {
    float v = 0.f;
#pragma unroll
    for (int ww = 0; ww < NF_WARPS; ++ww) v += s_part[ww][tid];
    float* d = dtable + (size_t)row * Edim + tid;
    *d = accumulate ? (*d + v) : v;
  }
}
static int news_fuse_tables_launch(const float* dout, const int32_t* cat, const int32_t* sub, int N, int Dout, int col0, int Ec, int Es,
                                   int n_cat, int n_sub, float p_drop, uint64_t seed, float* dcat_table, float* dsub_table,
                                   int accumulate, cudaStream_t st) {
  float inv_keep = 1.0f / (1.0f - p_drop);
  news_fuse_table_bwd_kernel<<<n_cat, NF_WARPS * 32, 0, st>>>(dout, cat, N, Dout, col0, Ec, Ec + Es, 0, p_drop,
                                                                     inv_keep, seed, dcat_table, accumulate);
  NNR_LAUNCH_CHECK("news_fuse_table_bwd_kernel(cat)");
  news_fuse_table_bwd_kernel<<<n_sub, NF_WARPS * 32, 0, st>>>(dout, sub, N, Dout, col0 + Ec, Es, Ec + Es, Ec,
                                                                     p_drop, inv_keep, seed, dsub_table, accumulate);
  NNR_LAUNCH_CHECK("news_fuse_table_bwd_kernel(sub)");
  return 0;
}
extern "C" int nnr_news_fuse_bwd(const float* dout, const int32_t* cat, const int32_t* sub, int N, int D2, int Ec, int Es,
                                 int n_cat, int n_sub, float p_drop, uint64_t seed, float* d_a, float* d_b,
                                 float* dcat_table, float* dsub_table, int accumulate, void* stream) {
  const bool tables = dcat_table || dsub_table;     // both NULL: only the activation split (the table gradients are then
                                                    // taken by nnr_news_fuse_tables_bwd, possibly on another stream)
  NNR_REQUIRE(dout && d_a && N > 0 && (!tables || (cat && sub && dcat_table && dsub_table)), NNR_ERR_ARG,
              "nnr_news_fuse_bwd: bad arguments");
  NNR_REQUIRE(Ec <= NF_MAXE && Es <= NF_MAXE, NNR_ERR_UNSUPPORTED, "nnr_news_fuse_bwd: embedding dim > %d", NF_MAXE);
  cudaStream_t st = (cudaStream_t)stream;
  const int DM = d_b ? 2 * D2 : D2;                 // d_b = NULL: single-modality encoders (CNE_Title / CNE_Content)
  int Dout = DM + Ec + Es;
  news_fuse_split_kernel<<<N, 256, 0, st>>>(dout, D2, Dout, d_a, d_b);
  NNR_LAUNCH_CHECK("news_fuse_split_kernel");
  if (!tables) return 0;
  return news_fuse_tables_launch(dout, cat, sub, N, Dout, DM, Ec, Es, n_cat, n_sub, p_drop, seed, dcat_table, dsub_table, accumulate, st);
}
extern "C" int nnr_news_fuse_tables_bwd(const float* dout, const int32_t* cat, const int32_t* sub, int N, int Dout, int col0, int Ec,
                                        int Es, int n_cat, int n_sub, float p_drop, uint64_t seed, float* dcat_table,
                                        float* dsub_table, int accumulate, void* stream) {
  NNR_REQUIRE(dout && cat && sub && dcat_table && dsub_table && N > 0 && col0 >= 0 && col0 + Ec + Es <= Dout, NNR_ERR_ARG,
              "nnr_news_fuse_tables_bwd: bad arguments");
  NNR_REQUIRE(Ec <= NF_MAXE && Es <= NF_MAXE, NNR_ERR_UNSUPPORTED, "nnr_news_fuse_tables_bwd: embedding dim > %d", NF_MAXE);
  return news_fuse_tables_launch(dout, cat, sub, N, Dout, col0, Ec, Es, n_cat, n_sub, p_drop, seed, dcat_table, dsub_table, accumulate,
                                 (cudaStream_t)stream);
}

// ------------------------------------------------------------------------------------------
// click predictor (model.py:127)
// ------------------------------------------------------------------------------------------
__global__ void rowdot_fwd_kernel(const float* __restrict__ a, const float* __restrict__ b, int R, int D, float* __restrict__ out) {
  int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (r >= R) return;
  float acc = 0.f;
  for (int d = lane; d < D; d += 32) acc += a[(size_t)r * D + d] * b[(size_t)r * D + d];
  acc = warp_sum(acc);
  if (lane == 0) out[r] = acc;
}
extern "C" int nnr_rowdot_fwd(const float* a, const float* b, int R, int D, float* out, void* stream) {
  NNR_REQUIRE(a && b && out && R > 0 && D > 0, NNR_ERR_ARG, "nnr_rowdot_fwd: bad arguments");
  rowdot_fwd_kernel<<<(R + 7) / 8, 256, 0, (cudaStream_t)stream>>>(a, b, R, D, out);
  NNR_LAUNCH_CHECK("rowdot_fwd_kernel");
  return 0;
}
__global__ void rowdot_bwd_kernel(const float* __restrict__ dout, const float* __restrict__ a, const float* __restrict__ b, int D,
                                  float* __restrict__ da, int acc_a, float* __restrict__ db, int acc_b) {
  int r = blockIdx.x;
  float g = dout[r];
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    size_t i = (size_t)r * D + d;
    float av = a[i], bv = b[i];
    if (da) da[i] = acc_a ? da[i] + g * bv : g * bv;
    if (db) db[i] = acc_b ? db[i] + g * av : g * av;
  }
}
extern "C" int nnr_rowdot_bwd(const float* dout, const float* a, const float* b, int R, int D, float* da, int accumulate_a,
                              float* db, int accumulate_b, void* stream) {
  NNR_REQUIRE(dout && a && b && R > 0 && D > 0, NNR_ERR_ARG, "nnr_rowdot_bwd: bad arguments");
  rowdot_bwd_kernel<<<R, 256, 0, (cudaStream_t)stream>>>(dout, a, b, D, da, accumulate_a, db, accumulate_b);
  NNR_LAUNCH_CHECK("rowdot_bwd_kernel");
  return 0;
}

// ------------------------------------------------------------------------------------------
// dropout
// ------------------------------------------------------------------------------------------
__global__ void dropout_kernel(const float* __restrict__ x, int64_t n, float p, float inv_keep, uint64_t seed, float* __restrict__ y) {
  if (p > 0.f) seed = nnr_resolve_seed(seed);
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  y[i] = x[i] * dropout_scale(seed, (uint64_t)i, p, inv_keep);
}
extern "C" int nnr_dropout(const float* x, int64_t n, float p_drop, uint64_t seed, float* y, void* stream) {
  NNR_REQUIRE(x && y && n > 0 && p_drop >= 0.f && p_drop < 1.f, NNR_ERR_ARG, "nnr_dropout: bad arguments");
  dropout_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(x, n, p_drop, 1.0f / (1.0f - p_drop), seed, y);
  NNR_LAUNCH_CHECK("dropout_kernel");
  return 0;
}

// ------------------------------------------------------------------------------------------
// clip_grad_norm_ + Adam over one flat buffer (trainer.py:118-120)
// ------------------------------------------------------------------------------------------
#define CA_BLOCKS 1184  // 148 SMs x 8
__global__ void sumsq_stage1(const float* __restrict__ g, int64_t n, double* __restrict__ part, int32_t* __restrict__ step_dev) {
  __shared__ double sh[8];
  if (step_dev && blockIdx.x == 0 && threadIdx.x == 0) *step_dev += 1;     // the update kernel (stream ordered after this one) reads it
  double acc = 0.0;
  int64_t stride = (int64_t)gridDim.x * blockDim.x * 4;
  for (int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < n; i += stride) {
    if (i + 3 < n) {
      float4 v = *reinterpret_cast<const float4*>(g + i);
      acc += (double)v.x * v.x + (double)v.y * v.y + (double)v.z * v.z + (double)v.w * v.w;
    } else {
      for (int64_t j = i; j < n; ++j) acc += (double)g[j] * g[j];
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += sh[w];
    part[blockIdx.x] = t;
  }
}
// total gradient norm from the per-CTA partials: every CTA adds them in the same fixed order (thread t takes partials
// t, t + 256, ...; then a shuffle tree and the eight warp sums in order), so all CTAs hold the identical value
__device__ __forceinline__ float grad_norm_from_partials(const double* __restrict__ part, int nparts, float grad_scale) {
  __shared__ double sh[8];
  __shared__ float s_norm;
  double acc = 0.0;
  for (int i = threadIdx.x; i < nparts; i += 256) acc += part[i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += sh[w];
    s_norm = (float)(sqrt(t) * (double)grad_scale);
  }
  __syncthreads();
  return s_norm;
}
template <bool VEC>
__global__ void __launch_bounds__(256) adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                   float* __restrict__ v, int64_t n, float lr, float b1, float b2, float eps,
                                                   float max_norm, float grad_scale, float bc1, float bc2_sqrt,
                                                   const double* __restrict__ part, int nparts, float* __restrict__ norm_out,
                                                   const int32_t* __restrict__ step_dev) {
  if (step_dev) {       // device-side step counter (CUDA-graph replay): the bias corrections of step *step_dev, in double like the host path
    const double st = (double)*step_dev;
    bc1 = (float)(1.0 - pow((double)b1, st));
    bc2_sqrt = (float)sqrt(1.0 - pow((double)b2, st));
  }
  const float total = grad_norm_from_partials(part, nparts, grad_scale);
  if (blockIdx.x == 0 && threadIdx.x == 0) norm_out[0] = total;
  float coef = 1.0f;
  if (max_norm > 0.f) coef = fminf(max_norm / (total + 1e-6f), 1.0f);   // torch clip_grad_norm_
  coef *= grad_scale;
  const float step_size = lr / bc1;
  auto upd = [&](float gi, float& pi, float& mi, float& vi) {
    gi *= coef;
    mi = mi * b1 + (1.f - b1) * gi;                  // exp_avg.lerp_(grad, 1-beta1)
    vi = vi * b2 + (1.f - b2) * gi * gi;
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    pi = pi - step_size * (mi / denom);
  };
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  if (VEC) {                                         // n % 4 == 0, 16-byte aligned buffers
    const int64_t n4 = n >> 2;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
      const float4 g4 = reinterpret_cast<const float4*>(g)[i];
      float4 p4 = reinterpret_cast<float4*>(p)[i], m4 = reinterpret_cast<float4*>(m)[i], v4 = reinterpret_cast<float4*>(v)[i];
      upd(g4.x, p4.x, m4.x, v4.x); upd(g4.y, p4.y, m4.y, v4.y); upd(g4.z, p4.z, m4.z, v4.z); upd(g4.w, p4.w, m4.w, v4.w);
      reinterpret_cast<float4*>(m)[i] = m4; reinterpret_cast<float4*>(v)[i] = v4; reinterpret_cast<float4*>(p)[i] = p4;
    }
  } else {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
      float pi = p[i], mi = m[i], vi = v[i];
      upd(g[i], pi, mi, vi);
      m[i] = mi; v[i] = vi; p[i] = pi;
    }
  }
}
extern "C" size_t nnr_flat_clip_adam_workspace_bytes(int64_t n) { (void)n; return CA_BLOCKS * sizeof(double); }
static int flat_clip_adam_impl(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, float lr,
                               float beta1, float beta2, float eps, float max_norm, float grad_scale, int32_t step, int32_t* step_dev,
                               float* norm_out, void* workspace, size_t workspace_bytes, void* stream) {
  NNR_REQUIRE(param && grad && exp_avg && exp_avg_sq && norm_out && workspace && n > 0 && (step >= 1 || step_dev), NNR_ERR_ARG,
              "nnr_flat_clip_adam: bad arguments");
  NNR_REQUIRE(workspace_bytes >= CA_BLOCKS * sizeof(double), NNR_ERR_WORKSPACE, "nnr_flat_clip_adam: workspace too small");
  NNR_REQUIRE(nnr_aligned16(grad), NNR_ERR_ALIGN, "nnr_flat_clip_adam: grad must be 16B aligned");
  cudaStream_t st = (cudaStream_t)stream;
  sumsq_stage1<<<CA_BLOCKS, 256, 0, st>>>(grad, n, (double*)workspace, step_dev);
  NNR_LAUNCH_CHECK("sumsq_stage1");
  double bc1d = 1.0, bc2d = 1.0;
  if (!step_dev) { bc1d = 1.0 - pow((double)beta1, (double)step); bc2d = 1.0 - pow((double)beta2, (double)step); }
  const bool vec = (n % 4 == 0) && nnr_aligned16(param) && nnr_aligned16(exp_avg) && nnr_aligned16(exp_avg_sq);
  if (vec)
    adam_kernel<true><<<CA_BLOCKS, 256, 0, st>>>(param, grad, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps, max_norm, grad_scale,
                                                 (float)bc1d, (float)sqrt(bc2d), (const double*)workspace, CA_BLOCKS, norm_out, step_dev);
  else
    adam_kernel<false><<<CA_BLOCKS, 256, 0, st>>>(param, grad, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps, max_norm, grad_scale,
                                                  (float)bc1d, (float)sqrt(bc2d), (const double*)workspace, CA_BLOCKS, norm_out, step_dev);
  NNR_LAUNCH_CHECK("adam_kernel");
  return 0;
}
extern "C" int nnr_flat_clip_adam(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, float lr,
                                  float beta1, float beta2, float eps, float max_norm, float grad_scale, int32_t step,
                                  float* norm_out, void* workspace, size_t workspace_bytes, void* stream) {
  return flat_clip_adam_impl(param, grad, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps, max_norm, grad_scale, step, nullptr, norm_out,
                             workspace, workspace_bytes, stream);
}
// the same with the step counter on the device: *step_dev is incremented by the call and the bias corrections are computed from
// it, so a captured CUDA graph of the call advances Adam's step on every replay
extern "C" int nnr_flat_clip_adam_dev(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, float lr,
                                      float beta1, float beta2, float eps, float max_norm, float grad_scale, int32_t* step_dev,
                                      float* norm_out, void* workspace, size_t workspace_bytes, void* stream) {
  NNR_REQUIRE(step_dev, NNR_ERR_ARG, "nnr_flat_clip_adam_dev: step_dev is NULL");
  return flat_clip_adam_impl(param, grad, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps, max_norm, grad_scale, 0, step_dev, norm_out,
                             workspace, workspace_bytes, stream);
}

// ------------------------------------------------------------------------------------------
// Stable descending sort of small integer keys (sequence lengths): sorted_idx[rank] = original index.
// Replaces the `torch.sort(length, descending=True)` of newsEncoders.py:112,114 on the device: ATen sorts these
// few thousand keys with a generic radix sort (~30 us per call, 9 calls per step); the keys are lengths <= 128, so a
// counting sort in one CTA is enough.  Ties keep ascending original index, which is what the (stable) CUDA radix
// sort behind torch.sort produces -- tests/test_ops_gpu.py compares the two on tie-heavy inputs.
// ------------------------------------------------------------------------------------------
#define LS_MAXKEY 1024
#define LS_MAXN 8192
__global__ void __launch_bounds__(1024) length_sort_desc_kernel(const int64_t* __restrict__ keys, int N, int max_key,
                                                                int64_t* __restrict__ sorted_idx) {
  __shared__ int s_cnt[LS_MAXKEY + 1];
  __shared__ short s_key[LS_MAXN];
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  for (int i = tid; i < N; i += 1024) s_key[i] = (short)min((long long)max_key, max(0ll, (long long)keys[i]));
  __syncthreads();
  for (int k = w; k <= max_key; k += 32) {                 // warp w owns keys w, w+32, ...
    int c = 0;
    for (int b = 0; b < N; b += 32) c += __popc(__ballot_sync(0xffffffffu, b + lane < N && s_key[b + lane] == k));
    if (lane == 0) s_cnt[k] = c;
  }
  __syncthreads();
  if (tid == 0) {                                          // start offsets, larger keys first
    int run = 0;
    for (int k = max_key; k >= 0; --k) { int c = s_cnt[k]; s_cnt[k] = run; run += c; }
  }
  __syncthreads();
  for (int k = w; k <= max_key; k += 32) {
    int pos = s_cnt[k];
    for (int b = 0; b < N; b += 32) {
      const bool hit = b + lane < N && s_key[b + lane] == k;
      const unsigned m = __ballot_sync(0xffffffffu, hit);
      if (hit) sorted_idx[pos + __popc(m & ((1u << lane) - 1))] = b + lane;
      pos += __popc(m);
    }
  }
}
// keys <= 255: one pass per 1024 elements.  Rank of an element = start of its key (larger keys first) + equal keys in
// earlier chunks + equal keys in earlier warps of the chunk + equal keys in earlier lanes of the warp (__match_any).
__global__ void __launch_bounds__(1024) length_sort_desc_small_kernel(const int64_t* __restrict__ keys, int N, int max_key,
                                                                      int64_t* __restrict__ sorted_idx) {
  __shared__ int s_base[256];            // running output position of each key
  __shared__ int s_wcnt[32][256];        // per warp: count, then exclusive start, of each key in the current chunk
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  if (tid < 256) s_base[tid] = 0;
  __syncthreads();
  for (int i = tid; i < N; i += 1024) atomicAdd(&s_base[(int)min((long long)max_key, max(0ll, (long long)keys[i]))], 1);
  __syncthreads();
  if (tid == 0) {
    int run = 0;
    for (int k = max_key; k >= 0; --k) { const int c = s_base[k]; s_base[k] = run; run += c; }
  }
  __syncthreads();
  for (int c0 = 0; c0 < N; c0 += 1024) {
    for (int j = tid; j < 32 * 256; j += 1024) (&s_wcnt[0][0])[j] = 0;
    __syncthreads();
    const int i = c0 + tid;
    const bool valid = i < N;
    const int key = valid ? (int)min((long long)max_key, max(0ll, (long long)keys[i])) : -1;
    const unsigned m = __match_any_sync(0xffffffffu, key);
    const int rank = __popc(m & ((1u << lane) - 1));
    if (valid && rank == 0) s_wcnt[w][key] = __popc(m);
    __syncthreads();
    if (tid <= max_key) {
      int run = s_base[tid];
#pragma unroll 8
      for (int ww = 0; ww < 32; ++ww) { const int c = s_wcnt[ww][tid]; s_wcnt[ww][tid] = run; run += c; }
      s_base[tid] = run;
    }
    __syncthreads();
    if (valid) sorted_idx[s_wcnt[w][key] + rank] = i;
    __syncthreads();
  }
}
extern "C" int nnr_length_sort_desc(const int64_t* keys, int N, int max_key, int64_t* sorted_idx, void* stream) {
  NNR_REQUIRE(keys && sorted_idx && N > 0, NNR_ERR_ARG, "nnr_length_sort_desc: bad arguments");
  NNR_REQUIRE(N <= LS_MAXN && max_key >= 0 && max_key <= LS_MAXKEY, NNR_ERR_UNSUPPORTED,
              "nnr_length_sort_desc: N <= %d and max_key <= %d", LS_MAXN, LS_MAXKEY);
  if (max_key <= 255) length_sort_desc_small_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(keys, N, max_key, sorted_idx);
  else length_sort_desc_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(keys, N, max_key, sorted_idx);
  NNR_LAUNCH_CHECK("length_sort_desc_kernel");
  return 0;
}
