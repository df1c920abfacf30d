#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
struct cz_ctx;
namespace cz {
// token ids >= vocab read row 0 and raise bit 1 of ctx->err_flag_dev (CZ_ERR_SYMBOL_RANGE)
int launch_embed(cz_ctx *ctx, const __nv_bfloat16 *table, const uint32_t *tok, float *x, int n_rows, int d, int vocab, cudaStream_t st);
int launch_embed_norm(cz_ctx *ctx, const __nv_bfloat16 *table, const uint32_t *tok, const float *w, float *x, __nv_bfloat16 *xb,
                      float *ssq, int n_rows, int d, int n_part, int vocab, cudaStream_t st);
int launch_rmsnorm(cz_ctx *ctx, const float *x, const float *w, const int *rows, __nv_bfloat16 *y, int n_out, int d, float eps,
                   cudaStream_t st);
int launch_rope_split(cz_ctx *ctx, const float *qkv, const int *pos, const int *kv_base, const float *cos_tab, const float *sin_tab,
                      __nv_bfloat16 *q, __nv_bfloat16 *k_arena, __nv_bfloat16 *v_arena, int n_rows, int nh, int nkv,
                      cudaStream_t st);
int launch_attn_rows(cz_ctx *ctx, const __nv_bfloat16 *q, const __nv_bfloat16 *k_arena, const __nv_bfloat16 *v_arena, const int *pos,
                     const int *kv_base, __nv_bfloat16 *out, int n_rows, int nh, int nkv, cudaStream_t st);
int launch_attn_mma(cz_ctx *ctx, const __nv_bfloat16 *q, const __nv_bfloat16 *k_arena, const __nv_bfloat16 *v_arena, const int *pos,
                    const int *kv_base, const int *tile_row0, const int *tile_n, int n_tiles, __nv_bfloat16 *out, int nh, int nkv,
                    cudaStream_t st);
int launch_attn_tc(cz_ctx *ctx, const __nv_bfloat16 *q, int n_rows, const __nv_bfloat16 *k_arena, const __nv_bfloat16 *vt_arena, int n_slots,
                   int ldv, const int *pos, const int *kv_base, const int *tile_row0, const int *tile_n, int n_tiles, __nv_bfloat16 *out,
                   int nh, int nkv, bool single_rows, cudaStream_t st);
}  // namespace cz
