// Test-only entry points (declared in include/candlezip_b200.h under "test hooks"): run one dense contraction
// through either engine on host buffers so tests can compare tcgen05 vs SIMT vs a host reference.
#include <vector>

#include "cz_common.cuh"
#include "gemm.h"

namespace cz {
int require_device(cz_ctx *ctx);
}
using namespace cz;

extern "C" int cz_test_gemm(cz_ctx *ctx, int engine, int M, int N, int K, const uint16_t *a_bf16, const uint16_t *b_bf16, int epi,
                            int bn, void *c_inout, int ldc) {
  CZ_TRY(require_device(ctx));
  CZ_CUDA_TRY(cudaSetDevice(ctx->device));
  const bool out_bf16 = epi == EPI_STORE_BF16 || epi == EPI_SWIGLU_BF16 || epi >= EPI_TANH_BF16;
  const size_t c_bytes = (size_t)M * ldc * (out_bf16 ? 2 : 4);
  void *da = nullptr, *db = nullptr, *dc = nullptr;
  CZ_CUDA_TRY(cudaMalloc(&da, (size_t)M * K * 2));
  CZ_CUDA_TRY(cudaMalloc(&db, (size_t)N * K * 2));
  CZ_CUDA_TRY(cudaMalloc(&dc, c_bytes));
  CZ_CUDA_TRY(cudaMemcpy(da, a_bf16, (size_t)M * K * 2, cudaMemcpyHostToDevice));
  CZ_CUDA_TRY(cudaMemcpy(db, b_bf16, (size_t)N * K * 2, cudaMemcpyHostToDevice));
  CZ_CUDA_TRY(cudaMemcpy(dc, c_inout, c_bytes, cudaMemcpyHostToDevice));
  int *daux = nullptr;
  if (epi == EPI_STORE_F32_COLMAX) {  // the column max itself is checked through the CDF parity tests; here it only must not corrupt C
    CZ_CUDA_TRY(cudaMalloc((void **)&daux, (size_t)N * 4 + 16));
    CZ_CUDA_TRY(cudaMemset(daux, 0x80, (size_t)N * 4 + 16));
  }
  GemmArgs g{};
  g.aux = daux;
  g.a = da; g.b = db; g.c = dc; g.M = M; g.N = N; g.K = K; g.lda = K; g.ldb = K; g.ldc = ldc; g.epi = epi; g.bn = bn;
  int rc = gemm(ctx, engine, g, ctx->stream);
  if (rc == CZ_OK) {
    cudaError_t e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) {
      set_error(std::string("gemm execution failed: ") + cudaGetErrorString(e));
      rc = CZ_ERR_CUDA;
    } else {
      cudaMemcpy(c_inout, dc, c_bytes, cudaMemcpyDeviceToHost);
    }
  }
  if (daux) cudaFree(daux);
  cudaFree(da);
  cudaFree(db);
  cudaFree(dc);
  return rc;
}
