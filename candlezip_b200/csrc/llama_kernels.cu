// K5 / K6: SmolLM (LLaMA-architecture) non-GEMM kernels: embedding gather, RMSNorm, RoPE + KV scatter,
// KV-cache attention (GQA), written so that every output row is a pure function of that row's inputs
// (fixed reduction orders, no atomics) -- the same kernels serve teacher-forced encode (many rows per sequence)
// and stepwise decode (one row per stream) and give bit-identical results for the same (sequence, position).
//
// Replaces candle-transformers 0.9.1 models::llama pieces reached from src/models.rs:94,110
// (embedding, RmsNorm, rotary embedding non-interleaved, Cache/KV concat, repeat_kv + softmax attention).
#include "cz_common.cuh"
#include "llama_kernels.h"

namespace czk {

// A token id >= vocab never indexes the table: it sets bit 1 of the ctx status word (-> CZ_ERR_SYMBOL_RANGE when the call
// collects its status) and row 0 is read instead.  Callers validate host-side ids before launching; this is the second line
// of defence for ids that only ever exist on the device (cz_encode_dev) and for anything a caller missed.
__device__ __forceinline__ uint32_t checked_token(const uint32_t *__restrict__ tok, int row, uint32_t vocab, int *__restrict__ err,
                                                  bool reporter) {
  const uint32_t t = tok[row];
  if (t < vocab) return t;
  if (reporter && err) atomicOr(err, 2);
  return 0u;
}

__global__ void embed_kernel(const __nv_bfloat16 *__restrict__ table, const uint32_t *__restrict__ tok, float *__restrict__ x,
                             int n_rows, int d, uint32_t vocab, int *__restrict__ err) {
  const int row = blockIdx.x;
  if (row >= n_rows) return;
  const __nv_bfloat16 *src = table + (size_t)checked_token(tok, row, vocab, err, threadIdx.x == 0) * d;
  float *dst = x + (size_t)row * d;
  for (int i = threadIdx.x; i < d; i += blockDim.x) dst[i] = __bfloat162float(src[i]);
}

// Embedding gather that also leaves the first layer's fused-RMSNorm inputs (see EPI_ADD_NORM in gemm.h): xb = bf16(x * w),
// ssq[row][0] = sum of x^2 (fixed order: per-lane strided partials, then an xor tree), ssq[row][1..n_part) = 0.  One warp per row.
__global__ void __launch_bounds__(128) embed_norm_kernel(const __nv_bfloat16 *__restrict__ table, const uint32_t *__restrict__ tok,
                                                         const float *__restrict__ w, float *__restrict__ x, __nv_bfloat16 *__restrict__ xb,
                                                         float *__restrict__ ssq, int n_rows, int d, int n_part, uint32_t vocab,
                                                         int *__restrict__ err) {
  const int row = blockIdx.x * 4 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= n_rows) return;
  const __nv_bfloat16 *src = table + (size_t)checked_token(tok, row, vocab, err, lane == 0) * d;
  float *dst = x + (size_t)row * d;
  __nv_bfloat16 *db = xb + (size_t)row * d;
  float ss = 0.f;
  for (int i = lane; i < d; i += 32) {
    const float v = __bfloat162float(src[i]);
    dst[i] = v;
    db[i] = __float2bfloat16_rn(v * w[i]);
    ss = fmaf(v, v, ss);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  if (lane < n_part) ssq[(size_t)row * n_part + lane] = lane == 0 ? ss : 0.f;
}

// One warp per output row.  src row = rows ? rows[j] : j.  y = bf16( (x * inv_rms) * w )
__global__ void __launch_bounds__(128) rmsnorm_kernel(const float *__restrict__ x, const float *__restrict__ w,
                                                      const int *__restrict__ rows, __nv_bfloat16 *__restrict__ y, int n_out,
                                                      int d, float eps) {
  const int j = blockIdx.x * 4 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (j >= n_out) return;
  const float *xr = x + (size_t)(rows ? rows[j] : j) * d;
  float ss = 0.f;
  for (int i = lane; i < d; i += 32) {
    float v = xr[i];
    ss = fmaf(v, v, ss);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  const float inv = 1.0f / sqrtf(ss / (float)d + eps);
  __nv_bfloat16 *yr = y + (size_t)j * d;
  for (int i = lane; i < d; i += 32) yr[i] = __float2bfloat16_rn(xr[i] * inv * w[i]);
}

// qkv f32 [row][(nh + 2*nkv) * 64] -> q bf16 [row][nh*64] (rotated), K/V arena rows at slot kv_base[row] + pos[row].
// rotate-half RoPE: (a, b) = (v[j], v[j+32]);  v[j] = a*cos - b*sin;  v[j+32] = b*cos + a*sin
__global__ void __launch_bounds__(128) rope_split_kernel(const float *__restrict__ qkv, const int *__restrict__ pos,
                                                         const int *__restrict__ kv_base, const float *__restrict__ cos_tab,
                                                         const float *__restrict__ sin_tab, __nv_bfloat16 *__restrict__ q,
                                                         __nv_bfloat16 *__restrict__ k_arena, __nv_bfloat16 *__restrict__ v_arena,
                                                         int n_rows, int nh, int nkv) {
  const int row = blockIdx.x;
  if (row >= n_rows) return;
  const int p = pos[row];
  const size_t slot = (size_t)kv_base[row] + (size_t)p;
  const int dq = nh * 64, dkv = nkv * 64;
  const float *src = qkv + (size_t)row * (dq + 2 * dkv);
  const float *cs = cos_tab + (size_t)p * 32, *sn = sin_tab + (size_t)p * 32;
  const int n_pairs = (nh + nkv) * 32;
  for (int i = threadIdx.x; i < n_pairs; i += blockDim.x) {
    const int head = i >> 5, j = i & 31;
    const float a = src[head * 64 + j], b = src[head * 64 + j + 32];
    const float c = cs[j], s = sn[j];
    const float ra = a * c - b * s, rb = b * c + a * s;
    if (head < nh) {
      __nv_bfloat16 *dst = q + (size_t)row * dq + head * 64;
      dst[j] = __float2bfloat16_rn(ra);
      dst[j + 32] = __float2bfloat16_rn(rb);
    } else {
      __nv_bfloat16 *dst = k_arena + slot * dkv + (head - nh) * 64;
      dst[j] = __float2bfloat16_rn(ra);
      dst[j + 32] = __float2bfloat16_rn(rb);
    }
  }
  const float *vsrc = src + dq + dkv;
  __nv_bfloat16 *vdst = v_arena + slot * dkv;
  for (int i = threadIdx.x; i < dkv; i += blockDim.x) vdst[i] = __float2bfloat16_rn(vsrc[i]);
}

// ---- attention, engine v1 ("rows"): one warp per (row, kv head); the G = nh/nkv query heads that share the kv head
// are processed together.  Canonical per-row order:
//   s_j   = (sum_{d=0..63} q_d k_jd, sequential fp32 FMA) * scale          lane owns keys j = lane, lane+32, ...
//   m     = max_j s_j
//   p_j   = expf(s_j - m);  l = per-lane sequential partial sums, then a fixed xor-shuffle tree
//   o_d   = sum_{j ascending} p_j v_jd (sequential fp32 FMA), lane owns d = 2*lane, 2*lane+1;  out = o * (1 / l)
constexpr int ATT_MAX_KEYS = 2048;

template <int G>
__global__ void __launch_bounds__(128) attn_rows_kernel(const __nv_bfloat16 *__restrict__ q, const __nv_bfloat16 *__restrict__ k_arena,
                                                        const __nv_bfloat16 *__restrict__ v_arena, const int *__restrict__ pos,
                                                        const int *__restrict__ kv_base, __nv_bfloat16 *__restrict__ out,
                                                        int n_rows, int nkv, float scale) {
  extern __shared__ float s_scores[];  // [4 warps][G][max_keys]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int item = blockIdx.x * 4 + warp;
  if (item >= n_rows * nkv) return;
  const int row = item / nkv, kvh = item % nkv;
  const int n_keys = pos[row] + 1;
  const int dkv = nkv * 64, dq = nkv * G * 64;
  const __nv_bfloat16 *kb = k_arena + (size_t)kv_base[row] * dkv + kvh * 64;
  const __nv_bfloat16 *vb = v_arena + (size_t)kv_base[row] * dkv + kvh * 64;
  float *sc = s_scores + (size_t)warp * G * ATT_MAX_KEYS;

  // scores
  float mx[G];
#pragma unroll
  for (int g = 0; g < G; g++) mx[g] = -INFINITY;
  for (int j = lane; j < n_keys; j += 32) {
    const uint4 *kr = reinterpret_cast<const uint4 *>(kb + (size_t)j * dkv);
    float acc[G];
#pragma unroll
    for (int g = 0; g < G; g++) acc[g] = 0.f;
#pragma unroll
    for (int c = 0; c < 8; c++) {
      uint4 kk = __ldg(kr + c);
      const __nv_bfloat162 *k2 = reinterpret_cast<const __nv_bfloat162 *>(&kk);
#pragma unroll
      for (int g = 0; g < G; g++) {
        const uint4 qq = *reinterpret_cast<const uint4 *>(q + (size_t)row * dq + (kvh * G + g) * 64 + c * 8);
        const __nv_bfloat162 *q2 = reinterpret_cast<const __nv_bfloat162 *>(&qq);
#pragma unroll
        for (int e = 0; e < 4; e++) {
          float2 kf = __bfloat1622float2(k2[e]), qf = __bfloat1622float2(q2[e]);
          acc[g] = fmaf(qf.x, kf.x, acc[g]);
          acc[g] = fmaf(qf.y, kf.y, acc[g]);
        }
      }
    }
#pragma unroll
    for (int g = 0; g < G; g++) {
      float s = acc[g] * scale;
      sc[g * ATT_MAX_KEYS + j] = s;
      mx[g] = fmaxf(mx[g], s);
    }
  }
  float lsum[G];
#pragma unroll
  for (int g = 0; g < G; g++) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx[g] = fmaxf(mx[g], __shfl_xor_sync(0xffffffffu, mx[g], o));
    float part = 0.f;
    for (int j = lane; j < n_keys; j += 32) {
      float p = expf(sc[g * ATT_MAX_KEYS + j] - mx[g]);
      sc[g * ATT_MAX_KEYS + j] = p;
      part += p;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    lsum[g] = part;
  }
  __syncwarp();
  // P.V : lane owns output dims 2*lane, 2*lane+1
  float o0[G], o1[G];
#pragma unroll
  for (int g = 0; g < G; g++) o0[g] = o1[g] = 0.f;
  for (int j = 0; j < n_keys; j++) {
    const __nv_bfloat162 v2 = *reinterpret_cast<const __nv_bfloat162 *>(vb + (size_t)j * dkv + 2 * lane);
    const float2 vf = __bfloat1622float2(v2);
#pragma unroll
    for (int g = 0; g < G; g++) {
      const float p = sc[g * ATT_MAX_KEYS + j];
      o0[g] = fmaf(p, vf.x, o0[g]);
      o1[g] = fmaf(p, vf.y, o1[g]);
    }
  }
#pragma unroll
  for (int g = 0; g < G; g++) {
    const float inv = 1.0f / lsum[g];
    __nv_bfloat162 h = __floats2bfloat162_rn(o0[g] * inv, o1[g] * inv);
    *reinterpret_cast<__nv_bfloat162 *>(out + (size_t)row * dq + (kvh * G + g) * 64 + 2 * lane) = h;
  }
}

}  // namespace czk

namespace cz {

// Every kernel of the step prefers the maximum shared-memory carve-out, like the 227 KB tcgen05 GEMM it alternates with:
// a carve-out change between two kernels makes the SM drain and reconfigure, several microseconds per launch in the
// launch-bound stepwise decoder.
template <class K>
static void prefer_max_smem(K kernel) {
  cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
}
void llama_kernels_set_carveout() {
  prefer_max_smem(czk::embed_kernel);
  prefer_max_smem(czk::rmsnorm_kernel);
  prefer_max_smem(czk::rope_split_kernel);
}

int launch_embed(cz_ctx *ctx, const __nv_bfloat16 *table, const uint32_t *tok, float *x, int n_rows, int d, int vocab, cudaStream_t st) {
  if (n_rows == 0) return CZ_OK;
  CZ_LAUNCH(ctx, CZ_K_ELEMWISE, (czk::embed_kernel<<<n_rows, 128, 0, st>>>(table, tok, x, n_rows, d, (uint32_t)vocab, ctx->err_flag_dev)));
  CZ_CHECK_LAUNCH();
  return CZ_OK;
}

int launch_embed_norm(cz_ctx *ctx, const __nv_bfloat16 *table, const uint32_t *tok, const float *w, float *x, __nv_bfloat16 *xb,
                      float *ssq, int n_rows, int d, int n_part, int vocab, cudaStream_t st) {
  if (n_rows == 0) return CZ_OK;
  if (n_part < 1 || n_part > 32) {
    set_error("embed_norm: 1 <= n_part <= 32");
    return CZ_ERR_INVALID;
  }
  CZ_LAUNCH(ctx, CZ_K_ELEMWISE,
            (czk::embed_norm_kernel<<<(unsigned)ceil_div(n_rows, 4), 128, 0, st>>>(table, tok, w, x, xb, ssq, n_rows, d, n_part, (uint32_t)vocab,
                                                                                   ctx->err_flag_dev)));
  CZ_CHECK_LAUNCH();
  return CZ_OK;
}

int launch_rmsnorm(cz_ctx *ctx, const float *x, const float *w, const int *rows, __nv_bfloat16 *y, int n_out, int d, float eps,
                   cudaStream_t st) {
  if (n_out == 0) return CZ_OK;
  CZ_LAUNCH(ctx, CZ_K_ELEMWISE, (czk::rmsnorm_kernel<<<(unsigned)ceil_div(n_out, 4), 128, 0, st>>>(x, w, rows, y, n_out, d, eps)));
  CZ_CHECK_LAUNCH();
  return CZ_OK;
}

int launch_rope_split(cz_ctx *ctx, const float *qkv, const int *pos, const int *kv_base, const float *cos_tab, const float *sin_tab,
                      __nv_bfloat16 *q, __nv_bfloat16 *k_arena, __nv_bfloat16 *v_arena, int n_rows, int nh, int nkv,
                      cudaStream_t st) {
  if (n_rows == 0) return CZ_OK;
  CZ_LAUNCH(ctx, CZ_K_ELEMWISE,
            (czk::rope_split_kernel<<<n_rows, 128, 0, st>>>(qkv, pos, kv_base, cos_tab, sin_tab, q, k_arena, v_arena, n_rows, nh, nkv)));
  CZ_CHECK_LAUNCH();
  return CZ_OK;
}

int launch_attn_rows(cz_ctx *ctx, const __nv_bfloat16 *q, const __nv_bfloat16 *k_arena, const __nv_bfloat16 *v_arena, const int *pos,
                     const int *kv_base, __nv_bfloat16 *out, int n_rows, int nh, int nkv, cudaStream_t st) {
  if (n_rows == 0) return CZ_OK;
  const int G = nh / nkv;
  const float scale = 0.125f;  // 1/sqrt(64)
  const size_t smem = (size_t)4 * G * czk::ATT_MAX_KEYS * sizeof(float);
  const unsigned grid = (unsigned)ceil_div((size_t)n_rows * nkv, 4);
#define CZ_ATT_CASE(G_)                                                                                                   \
  case G_: {                                                                                                              \
    static bool attr = false;                                                                                             \
    if (!attr) {                                                                                                          \
      CZ_CUDA_TRY(cudaFuncSetAttribute(czk::attn_rows_kernel<G_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
      attr = true;                                                                                                        \
    }                                                                                                                     \
    CZ_LAUNCH(ctx, CZ_K_ATTN,                                                                                             \
              (czk::attn_rows_kernel<G_><<<grid, 128, smem, st>>>(q, k_arena, v_arena, pos, kv_base, out, n_rows, nkv, scale))); \
    break;                                                                                                                \
  }
  switch (G) {
    CZ_ATT_CASE(1)
    CZ_ATT_CASE(2)
    CZ_ATT_CASE(3)
    CZ_ATT_CASE(4)
    default:
      set_error("attention: unsupported GQA group size");
      return CZ_ERR_UNSUPPORTED;
  }
#undef CZ_ATT_CASE
  CZ_CHECK_LAUNCH();
  return CZ_OK;
}

}  // namespace cz
