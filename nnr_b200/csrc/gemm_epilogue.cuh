// Fused GEMM epilogues shared by the SIMT and tcgen05 backends (see NNR_EPI_* in nnr_b200.h).
#pragma once
#include "common.cuh"
#include "../../include/nnr_b200.h"

struct EpiP {
  float* C; int64_t ldc; int accumulate; int epilogue;
  const float* bias;
  const float* aux; int64_t ldaux;
  float* aux_out; int64_t ldaux_out;
  const float* rowbias; int64_t ldrowbias; const int32_t* rowmap;
  float p_drop, inv_keep; uint64_t seed; int N;
};

__device__ __forceinline__ void epi_store(const EpiP& e, int m, int n, float acc) {
  float v;
  switch (e.epilogue) {
    case NNR_EPI_BIAS: v = acc + e.bias[n]; break;
    case NNR_EPI_BIAS_TANH: v = tanh_fast(acc + e.bias[n]); break;
    case NNR_EPI_BIAS_RELU_RES: {
      float r = fmaxf(acc + (e.bias ? e.bias[n] : 0.f), 0.f);
      if (e.aux_out) e.aux_out[(size_t)m * e.ldaux_out + n] = r;
      v = r + (e.aux ? e.aux[(size_t)m * e.ldaux + n] : 0.f);
      v *= dropout_scale(nnr_resolve_seed(e.seed), (uint64_t)m * (uint64_t)e.N + n, e.p_drop, e.inv_keep);
      break;
    }
    case NNR_EPI_GATE: {
      float g = sigmoid_fast(acc + e.rowbias[(size_t)e.rowmap[m] * e.ldrowbias + n]);
      if (e.aux_out) e.aux_out[(size_t)m * e.ldaux_out + n] = g;
      v = e.aux[(size_t)m * e.ldaux + n] * g;
      break;
    }
    case NNR_EPI_ADD_AUX: v = acc + e.aux[(size_t)m * e.ldaux + n]; break;
    default: v = acc; break;
  }
  float* c = e.C + (size_t)m * e.ldc + n;
  *c = e.accumulate ? (*c + v) : v;
}


static inline EpiP make_epi(const nnr_gemm_args* a) {
  EpiP e;
  e.C = a->C; e.ldc = a->ldc; e.accumulate = a->accumulate; e.epilogue = a->epilogue; e.bias = a->bias;
  e.aux = a->aux; e.ldaux = a->ldaux; e.aux_out = a->aux_out; e.ldaux_out = a->ldaux_out;
  e.rowbias = a->rowbias; e.ldrowbias = a->ldrowbias; e.rowmap = a->rowmap;
  e.p_drop = a->p_drop; e.inv_keep = 1.0f / (1.0f - a->p_drop); e.seed = a->seed; e.N = a->N;
  return e;
}

// deterministic split-K: partial[s][M][N] summed in split order, then the epilogue
static __global__ void gemm_splitk_reduce_kernel(const float* __restrict__ partial, int splits, int M, int N,
                                          const int32_t* __restrict__ m_dev, EpiP epi) {
  if (m_dev) M = min(M, *m_dev);
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)M * N) return;
  int m = (int)(i / N), n = (int)(i - (size_t)m * N);
  float acc = 0.f;
  for (int s = 0; s < splits; ++s) acc += partial[((size_t)s * M + m) * N + n];
  epi_store(epi, m, n, acc);
}

