// tcgen05 / TMA GEMM backend (placeholder until the tensor-core kernel lands; nnr_gemm then
// routes every call to the exact-fp32 kernel in gemm_simt.cu).
#include "common.cuh"
#include "../../include/nnr_b200.h"

int nnr_gemm_tc_supported(const nnr_gemm_args* a) { (void)a; return 0; }
size_t nnr_gemm_tc_workspace_bytes(const nnr_gemm_args* a) { (void)a; return 0; }
int nnr_gemm_tc(const nnr_gemm_args* a, void* stream) {
  (void)a; (void)stream;
  nnr_set_error("nnr_gemm: tensor-core backend not built");
  return NNR_ERR_UNSUPPORTED;
}
