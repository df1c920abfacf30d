// Plain SIMT GEMM (CUDA cores, fp32 FMA): the second dense-contraction engine (CZ_ENGINE_SIMT).
// Same contract and epilogues as gemm_tcgen05.cu.  It exists (a) as an on-GPU cross-check for the tcgen05 kernel
// in tests and (b) as a bring-up engine; it is ~50x slower and never the benchmarked path.
// Row-invariant by construction: each output element is one sequential fp32 FMA chain over k = 0..K-1.
#include "cz_common.cuh"
#include "gemm.h"

namespace czk {

constexpr int ST = 64;   // tile
constexpr int SK = 16;

__device__ __forceinline__ float silu_mul_s(float g, float u) { return __fdividef(g, 1.0f + __expf(-g)) * u; }

template <int EPI>
__global__ void __launch_bounds__(256) gemm_simt_kernel(const __nv_bfloat16 *__restrict__ A, const __nv_bfloat16 *__restrict__ B,
                                                        void *__restrict__ C, int M, int N, int K, int lda, int ldb, int ldc,
                                                        int bn) {
  __shared__ float sa[SK][ST + 1], sb[SK][ST + 1];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int m0 = blockIdx.y * ST, n0 = blockIdx.x * ST;
  float acc[4][4] = {};
  for (int k0 = 0; k0 < K; k0 += SK) {
    for (int i = threadIdx.x; i < ST * SK; i += 256) {
      int r = i / SK, k = i % SK;
      sa[k][r] = (m0 + r < M) ? __bfloat162float(A[(size_t)(m0 + r) * lda + k0 + k]) : 0.f;
      sb[k][r] = (n0 + r < N) ? __bfloat162float(B[(size_t)(n0 + r) * ldb + k0 + k]) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < SK; k++) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; i++) a[i] = sa[k][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; j++) b[j] = sb[k][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const int m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const int n = n0 + tx * 4 + j;
      if (n >= N) continue;
      if (EPI == cz::EPI_STORE_F32) ((float *)C)[(size_t)m * ldc + n] = acc[i][j];
      else if (EPI == cz::EPI_ADD_F32) ((float *)C)[(size_t)m * ldc + n] += acc[i][j];
      else if (EPI == cz::EPI_STORE_BF16) ((__nv_bfloat16 *)C)[(size_t)m * ldc + n] = __float2bfloat16_rn(acc[i][j]);
      else if (EPI == cz::EPI_TANH_BF16) ((__nv_bfloat16 *)C)[(size_t)m * ldc + n] = __float2bfloat16_rn(tanhf(acc[i][j]));
      else if (EPI == cz::EPI_SIGMOID_BF16)
        ((__nv_bfloat16 *)C)[(size_t)m * ldc + n] = __float2bfloat16_rn(1.0f / (1.0f + expf(-acc[i][j])));
      else if (EPI == cz::EPI_RELUSQ_BF16) {
        const float q = fmaxf(acc[i][j], 0.f);
        ((__nv_bfloat16 *)C)[(size_t)m * ldc + n] = __float2bfloat16_rn(q * q);
      }
    }
  }
}

// swiglu over the packed gate/up layout: T[m][g*bn + j] gate, T[m][g*bn + bn/2 + j] up  ->  out[m][g*bn/2 + j]
__global__ void swiglu_packed_kernel(const float *__restrict__ T, __nv_bfloat16 *__restrict__ out, int M, int N, int ldt, int ldc, int bn) {
  const int half = bn / 2;
  size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t total = (size_t)M * (N / 2);
  if (idx >= total) return;
  int m = (int)(idx / (N / 2)), o = (int)(idx % (N / 2));
  int g = o / half, j = o % half;
  float gv = T[(size_t)m * ldt + g * bn + j], uv = T[(size_t)m * ldt + g * bn + half + j];
  out[(size_t)m * ldc + o] = __float2bfloat16_rn(silu_mul_s(gv, uv));
}

}  // namespace czk

namespace cz {

int gemm_simt(cz_ctx *ctx, const GemmArgs &g, cudaStream_t stream) {
  if (g.M <= 0 || g.N <= 0) return CZ_OK;
  if (g.K % czk::SK) {
    set_error("gemm_simt: K must be a multiple of 16");
    return CZ_ERR_INVALID;
  }
  dim3 grid((unsigned)ceil_div(g.N, czk::ST), (unsigned)ceil_div(g.M, czk::ST));
  const int g_fam = g.fam;
  const __nv_bfloat16 *A = (const __nv_bfloat16 *)g.a, *B = (const __nv_bfloat16 *)g.b;
  if (g.epi == EPI_SWIGLU_BF16) {
    // two-step: f32 temp then the packed swiglu (temp lives in the ctx scratch)
    size_t bytes = (size_t)g.M * g.N * sizeof(float);
    CZ_TRY(ensure_scratch(ctx, bytes));
    float *T = (float *)ctx->scratch;
    CZ_LAUNCH(ctx, g_fam,
              (czk::gemm_simt_kernel<EPI_STORE_F32><<<grid, 256, 0, stream>>>(A, B, T, g.M, g.N, g.K, g.lda, g.ldb, g.N, g.bn)));
    CZ_CHECK_LAUNCH();
    size_t total = (size_t)g.M * (g.N / 2);
    CZ_LAUNCH(ctx, CZ_K_ELEMWISE,
              (czk::swiglu_packed_kernel<<<(unsigned)ceil_div(total, 256), 256, 0, stream>>>(T, (__nv_bfloat16 *)g.c, g.M, g.N,
                                                                                            g.N, g.ldc, g.bn)));
    CZ_CHECK_LAUNCH();
    return CZ_OK;
  }
  if (g.epi == EPI_STORE_F32 || g.epi == EPI_STORE_F32_COLMAX)
    CZ_LAUNCH(ctx, g_fam,
              (czk::gemm_simt_kernel<EPI_STORE_F32><<<grid, 256, 0, stream>>>(A, B, g.c, g.M, g.N, g.K, g.lda, g.ldb, g.ldc, g.bn)));
  else if (g.epi == EPI_ADD_F32)
    CZ_LAUNCH(ctx, g_fam,
              (czk::gemm_simt_kernel<EPI_ADD_F32><<<grid, 256, 0, stream>>>(A, B, g.c, g.M, g.N, g.K, g.lda, g.ldb, g.ldc, g.bn)));
  else if (g.epi == EPI_STORE_BF16)
    CZ_LAUNCH(ctx, g_fam,
              (czk::gemm_simt_kernel<EPI_STORE_BF16><<<grid, 256, 0, stream>>>(A, B, g.c, g.M, g.N, g.K, g.lda, g.ldb, g.ldc, g.bn)));
#define CZ_SIMT_CASE(E_)                                                                                                          \
  else if (g.epi == E_) CZ_LAUNCH(ctx, g_fam,                                                                                     \
                                  (czk::gemm_simt_kernel<E_><<<grid, 256, 0, stream>>>(A, B, g.c, g.M, g.N, g.K, g.lda, g.ldb, g.ldc, g.bn)))
  CZ_SIMT_CASE(EPI_TANH_BF16);
  CZ_SIMT_CASE(EPI_SIGMOID_BF16);
  CZ_SIMT_CASE(EPI_RELUSQ_BF16);
#undef CZ_SIMT_CASE
  else {
    set_error("gemm_simt: unknown epilogue");
    return CZ_ERR_INVALID;
  }
  CZ_CHECK_LAUNCH();
  return CZ_OK;
}

int gemm(cz_ctx *ctx, int engine, const GemmArgs &g, cudaStream_t stream) {
  if (engine == CZ_ENGINE_SIMT) return gemm_simt(ctx, g, stream);
  return gemm_tcgen05(ctx, g, stream);
}

}  // namespace cz
