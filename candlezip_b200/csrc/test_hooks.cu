// Test-only entry points (declared in include/candlezip_b200.h under "test hooks"): run one dense contraction
// through either engine on host buffers so tests can compare tcgen05 vs SIMT vs a host reference.
#include <cstdlib>
#include <cstring>
#include <vector>

#include "cz_common.cuh"
#include "gemm.h"
#include "llama_kernels.h"

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

extern "C" int cz_test_gemm_norm(cz_ctx *ctx, int M, int N, int K, int N2, const uint16_t *a_bf16, const uint16_t *b_bf16,
                                 const uint16_t *b2_bf16, const float *w_next, float eps, float *x_inout, uint16_t *xb_out, float *ssq_out,
                                 uint16_t *out2) {
  CZ_TRY(require_device(ctx));
  CZ_CUDA_TRY(cudaSetDevice(ctx->device));
  if (N % 192 != 0 || K % 64 != 0 || N2 % 8 != 0) {
    set_error("cz_test_gemm_norm: N % 192, K % 64, N2 % 8");
    return CZ_ERR_INVALID;
  }
  const bool norm_tma = getenv("CZ_NORM_TRANSPOSE") == nullptr;  // same switch as the model (model_core.cu)
  const int n_part_host = (N / 192) * 3;                     // the caller's ssq_out row length
  const int n_part = norm_tma ? N / 192 : n_part_host;
  void *da = nullptr, *db = nullptr, *db2 = nullptr, *dx = nullptr, *dxb = nullptr, *dssq = nullptr, *dw = nullptr, *dout = nullptr;
  CZ_CUDA_TRY(cudaMalloc(&da, (size_t)M * K * 2));
  CZ_CUDA_TRY(cudaMalloc(&db, (size_t)N * K * 2));
  CZ_CUDA_TRY(cudaMalloc(&db2, (size_t)N2 * N * 2));
  CZ_CUDA_TRY(cudaMalloc(&dx, (size_t)M * N * 4));
  CZ_CUDA_TRY(cudaMalloc(&dxb, (size_t)M * N * 2));
  CZ_CUDA_TRY(cudaMalloc(&dssq, (size_t)M * n_part * 4));
  CZ_CUDA_TRY(cudaMalloc(&dw, (size_t)N * 4));
  CZ_CUDA_TRY(cudaMalloc(&dout, (size_t)M * N2 * 2));
  CZ_CUDA_TRY(cudaMemcpy(da, a_bf16, (size_t)M * K * 2, cudaMemcpyHostToDevice));
  CZ_CUDA_TRY(cudaMemcpy(db, b_bf16, (size_t)N * K * 2, cudaMemcpyHostToDevice));
  CZ_CUDA_TRY(cudaMemcpy(db2, b2_bf16, (size_t)N2 * N * 2, cudaMemcpyHostToDevice));
  CZ_CUDA_TRY(cudaMemcpy(dx, x_inout, (size_t)M * N * 4, cudaMemcpyHostToDevice));
  CZ_CUDA_TRY(cudaMemcpy(dw, w_next, (size_t)N * 4, cudaMemcpyHostToDevice));
  GemmArgs g{};
  g.a = da; g.b = db; g.c = dx; g.M = M; g.N = N; g.K = K; g.lda = K; g.ldb = K; g.ldc = N; g.epi = norm_tma ? EPI_ADD_NORM_TMA : EPI_ADD_NORM; g.bn = 192;
  g.norm.w_next = (const float *)dw; g.norm.xb = dxb; g.norm.ssq_out = (float *)dssq;
  int rc = gemm(ctx, CZ_ENGINE_TCGEN05, g, ctx->stream);
  if (rc == CZ_OK) {
    GemmArgs h{};
    h.a = dxb; h.b = db2; h.c = dout; h.M = M; h.N = N2; h.K = N; h.lda = N; h.ldb = N; h.ldc = N2; h.epi = EPI_STORE_BF16; h.bn = 192;
    h.norm.ssq_in = (const float *)dssq; h.norm.n_part_in = n_part; h.norm.inv_d = 1.0f / (float)N; h.norm.eps = eps;
    rc = gemm(ctx, CZ_ENGINE_TCGEN05, h, ctx->stream);
  }
  if (rc == CZ_OK) {
    cudaError_t e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) {
      set_error(std::string("gemm execution failed: ") + cudaGetErrorString(e));
      rc = CZ_ERR_CUDA;
    } else {
      cudaMemcpy(x_inout, dx, (size_t)M * N * 4, cudaMemcpyDeviceToHost);
      cudaMemcpy(xb_out, dxb, (size_t)M * N * 2, cudaMemcpyDeviceToHost);
      memset(ssq_out, 0, (size_t)M * n_part_host * 4);
      cudaMemcpy2D(ssq_out, (size_t)n_part_host * 4, dssq, (size_t)n_part * 4, (size_t)n_part * 4, (size_t)M, cudaMemcpyDeviceToHost);
      cudaMemcpy(out2, dout, (size_t)M * N2 * 2, cudaMemcpyDeviceToHost);
    }
  }
  for (void *p : {da, db, db2, dx, dxb, dssq, dw, dout}) cudaFree(p);
  return rc;
}

// One causal GQA attention pass of the tcgen05 kernel over ONE sequence of n_pos positions (the K5 kernel behind
// src/models.rs:94,110), on host bf16 buffers: q [n_pos][nh*64], k / v [n_pos][nkv*64] (already RoPE'd -- the hook feeds the kernel
// directly so that tests can place adversarial scores in chosen key blocks).  mode 0: teacher-forced 128-position tiles;
// mode 1: every position as its own single-row decode tile (the stacked-GQA path of the stepwise decoder).  out [n_pos][nh*64] bf16.
extern "C" int cz_test_attention(cz_ctx *ctx, int n_pos, int nh, int nkv, const uint16_t *q_bf16, const uint16_t *k_bf16,
                                 const uint16_t *v_bf16, int mode, uint16_t *out_bf16) {
  CZ_TRY(require_device(ctx));
  CZ_CUDA_TRY(cudaSetDevice(ctx->device));
  if (n_pos <= 0 || nh <= 0 || nkv <= 0 || nh % nkv || nh / nkv > 4 || !q_bf16 || !k_bf16 || !v_bf16 || !out_bf16) {
    set_error("cz_test_attention: bad arguments");
    return CZ_ERR_INVALID;
  }
  const size_t qn = (size_t)n_pos * nh * 64, kn = (size_t)n_pos * nkv * 64;
  std::vector<int> pos(n_pos), base(n_pos, 0), trow, tn;
  for (int i = 0; i < n_pos; i++) pos[i] = i;
  if (mode == 0) {
    for (int p = 0; p < n_pos; p += 128) {
      trow.push_back(p);
      tn.push_back(n_pos - p < 128 ? n_pos - p : 128);
    }
  } else {
    for (int p = 0; p < n_pos; p++) {
      trow.push_back(p);
      tn.push_back(1);
    }
  }
  const size_t nt = trow.size();
  void *dq = nullptr, *dk = nullptr, *dv = nullptr, *dout = nullptr, *dmeta = nullptr;
  CZ_CUDA_TRY(cudaMalloc(&dq, qn * 2));
  CZ_CUDA_TRY(cudaMalloc(&dk, kn * 2));
  CZ_CUDA_TRY(cudaMalloc(&dv, kn * 2));
  CZ_CUDA_TRY(cudaMalloc(&dout, qn * 2));
  CZ_CUDA_TRY(cudaMalloc(&dmeta, ((size_t)2 * n_pos + 2 * nt) * 4));
  int *dpos = (int *)dmeta, *dbase = dpos + n_pos, *dtrow = dbase + n_pos, *dtn = dtrow + nt;
  CZ_CUDA_TRY(cudaMemcpy(dq, q_bf16, qn * 2, cudaMemcpyHostToDevice));
  CZ_CUDA_TRY(cudaMemcpy(dk, k_bf16, kn * 2, cudaMemcpyHostToDevice));
  CZ_CUDA_TRY(cudaMemcpy(dv, v_bf16, kn * 2, cudaMemcpyHostToDevice));
  CZ_CUDA_TRY(cudaMemset(dout, 0, qn * 2));
  CZ_CUDA_TRY(cudaMemcpy(dpos, pos.data(), (size_t)n_pos * 4, cudaMemcpyHostToDevice));
  CZ_CUDA_TRY(cudaMemcpy(dbase, base.data(), (size_t)n_pos * 4, cudaMemcpyHostToDevice));
  CZ_CUDA_TRY(cudaMemcpy(dtrow, trow.data(), nt * 4, cudaMemcpyHostToDevice));
  CZ_CUDA_TRY(cudaMemcpy(dtn, tn.data(), nt * 4, cudaMemcpyHostToDevice));
  int rc = launch_attn_tc(ctx, (const __nv_bfloat16 *)dq, n_pos, (const __nv_bfloat16 *)dk, (const __nv_bfloat16 *)dv, n_pos, 0, dpos, dbase,
                          dtrow, dtn, (int)nt, (__nv_bfloat16 *)dout, nh, nkv, mode != 0, ctx->stream);
  if (rc == CZ_OK) {
    cudaError_t e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) {
      set_error(std::string("attention execution failed: ") + cudaGetErrorString(e));
      rc = CZ_ERR_CUDA;
    } else {
      cudaMemcpy(out_bf16, dout, qn * 2, cudaMemcpyDeviceToHost);
    }
  }
  for (void *p : {dq, dk, dv, dout, dmeta}) cudaFree(p);
  return rc;
}
