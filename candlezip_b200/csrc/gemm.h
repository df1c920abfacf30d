// Dense-contraction interface shared by the tcgen05 and SIMT engines.
#pragma once
#include <cuda_runtime.h>

struct cz_ctx;

namespace cz {

enum GemmEpilogue {
  EPI_STORE_F32 = 0,    // C[M][ldc] f32  = A*B^T
  EPI_ADD_F32 = 1,      // C[M][ldc] f32 += A*B^T          (residual add fused into o_proj / down_proj)
  EPI_SWIGLU_BF16 = 2,  // C[M][ldc] bf16 = silu(gate)*up   (B rows packed per tile: BN/2 gate rows then BN/2 up rows; N = 2*ffn)
  EPI_STORE_BF16 = 3,   // C[M][ldc] bf16 = A*B^T
  EPI_STORE_F32_COLMAX = 4,  // EPI_STORE_F32 + aux[n] = max over rows m of C[m][n], as an order-preserving int (atomicMax);
                             // used by the vocab-major LM head so the CDF kernel does not need its own max pass
  EPI_TANH_BF16 = 5,     // C bf16 = tanh(A*B^T)       (RWKV-7 decay LoRA, rwkv7.rs:212)
  EPI_SIGMOID_BF16 = 6,  // C bf16 = sigmoid(A*B^T)    (RWKV-7 gate LoRA, rwkv7.rs:234)
  EPI_RELUSQ_BF16 = 7,   // C bf16 = relu(A*B^T)^2     (RWKV-7 FFN, rwkv7.rs:426)
  EPI_QKV_ROPE = 8,      // SmolLM q/k/v projection: rotate-half RoPE on the q and k heads, bf16 q rows + K/V arena scatter
                         // (N = (nh + 2 nkv) * 64, BN = 192 = three whole heads per tile); needs GemmArgs::rope
  EPI_ADD_NORM = 9,      // residual add WITH the next RMSNorm's inputs (tcgen05 engine only; needs GemmArgs::norm):
                         //   x = C[M][ldc] f32 += A*B^T;  xb[M][N] bf16 = bf16(x * w_next[n]);  ssq_out[m][part] = partial sums of x^2.
                         // The consumer GEMM takes xb as its A operand and multiplies its accumulator rows by
                         // 1 / sqrt(sum(parts) / N + eps): RMSNorm without a separate pass over the fp32 residual.
  EPI_ADD_NORM_TMA = 10,  // the same contract as EPI_ADD_NORM with n_part_out = N / BN: thread-per-row epilogue, residual in and out by TMA
};

// RMSNorm fusion operands (all device pointers).  Producer side: EPI_ADD_NORM.  Consumer side (any bf16-output epilogue and
// EPI_QKV_ROPE): ssq_in != nullptr scales accumulator row m by 1 / sqrt((sum_p ssq_in[m * n_part_in + p]) * inv_d + eps).
struct NormExt {
  const float *w_next = nullptr;  // [N] weight of the norm that follows the residual add
  void *xb = nullptr;             // bf16 [M][N] (ld = N)
  float *ssq_out = nullptr;       // [M][n_part_out], n_part_out = (N / BN) * 3 (three epilogue warps per TMEM lane quadrant)
  const float *ssq_in = nullptr;  // [M][n_part_in]
  int n_part_in = 0;
  float inv_d = 0.f, eps = 0.f;
};

// extra operands of EPI_QKV_ROPE (all device pointers)
struct RopeExt {
  const int *pos = nullptr, *kv_base = nullptr;        // per row: position, first KV slot of its sequence
  const float *cos_tab = nullptr, *sin_tab = nullptr;  // [max_pos][32]
  void *q = nullptr, *k_arena = nullptr, *v_arena = nullptr;  // bf16: q [M][nh*64]; arenas [slot][nkv*64]
  int nh = 0, nkv = 0;
};

struct GemmArgs {
  const void *a;  // bf16 [M][lda], K-major
  const void *b;  // bf16 [N][ldb], K-major (nn.Linear weight layout [out][in])
  void *c;
  int M, N, K;
  int lda, ldb, ldc;
  int epi;  // GemmEpilogue
  int bn;   // tile width: 192 or 256 (also the gate/up packing granularity for EPI_SWIGLU_BF16)
  int fam = 0;         // profiling family (CZ_K_GEMM, CZ_K_GEMM_O, ...)
  int *aux = nullptr;  // EPI_STORE_F32_COLMAX: per-column running max, must be pre-filled with INT_MIN
  RopeExt rope;        // EPI_QKV_ROPE
  NormExt norm;        // EPI_ADD_NORM (producer) / row scale of the consumer epilogues
};

int gemm_tcgen05(cz_ctx *ctx, const GemmArgs &g, cudaStream_t stream);
int gemm_simt(cz_ctx *ctx, const GemmArgs &g, cudaStream_t stream);
// engine: CZ_ENGINE_TCGEN05 / CZ_ENGINE_SIMT
int gemm(cz_ctx *ctx, int engine, const GemmArgs &g, cudaStream_t stream);

}  // namespace cz
