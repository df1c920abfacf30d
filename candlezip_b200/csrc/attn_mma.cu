// K5, engine v2: causal GQA attention on the tensor cores (mma.sync m16n8k16 bf16, fp32 accumulate), flash-style
// online softmax, K/V staged through shared memory with cp.async double buffering and XOR-swizzled rows.
// Tiles hold up to 64 query positions.
//
// Replaces candle-transformers llama attention (repeat_kv + QK^T/sqrt(d) + f32 softmax + PV) reached from
// src/models.rs:94,110.  One CTA = one tile of up to 64 consecutive query positions of ONE sequence x one KV head;
// its 4*G warps are (16-position sub-tile) x (the G query heads sharing that KV head), so a K/V block is fetched
// once and used by all G heads and all 64 positions.  The same kernel serves stepwise decode: a tile with a single
// valid position (only the first G warps compute; everyone still helps stream K/V).
//
// ROW INVARIANCE (decode safety).  For a given (sequence, position, head) the arithmetic is a fixed sequence:
// keys are consumed in blocks of 64 anchored at key 0; per block  S = Q K^T (4 MMAs of k=16 per 8-key tile, in
// order), mask, m' = max(m, rowmax), P = exp2(S*c - m'), l = l*a + sum(P), O = O*a + bf16(P) V.  Rows of an MMA
// are independent, fully-masked blocks are exact no-ops (a = exp2(0) = 1, P = 0), so the result does not depend on
// which other rows share the tile, on the tile's size, or on batch / GPU count.  Teacher-forced encode and stepwise
// decode therefore produce bit-identical outputs (tests: stepwise == teacher-forced bitwise).
#include "cz_common.cuh"
#include "llama_kernels.h"

namespace czk {

constexpr int BKV = 64;  // keys per block

__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t &r0, uint32_t &r1, uint32_t &r2, uint32_t &r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t &r0, uint32_t &r1, uint32_t &r2, uint32_t &r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void mma_bf16(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void *src, int src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t *>(&h);
}

template <int G>
__global__ void __launch_bounds__(128 * G) attn_mma_kernel(const __nv_bfloat16 *__restrict__ q, const __nv_bfloat16 *__restrict__ k_arena,
                                                           const __nv_bfloat16 *__restrict__ v_arena, const int *__restrict__ pos,
                                                           const int *__restrict__ kv_base, const int *__restrict__ tile_row0,
                                                           const int *__restrict__ tile_n, __nv_bfloat16 *__restrict__ out, int nkv) {
  __shared__ __align__(128) uint8_t smem[2 * 2 * BKV * 128];  // [buf][K|V][64 rows][128 B], rows XOR-swizzled by 16-byte chunk
  const int tile = blockIdx.x, kvh = blockIdx.y;
  const int row0 = tile_row0[tile], nq = tile_n[tile];
  const int p0 = pos[row0];
  const size_t base = (size_t)kv_base[row0];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int sub = warp / G, head = kvh * G + (warp % G);
  const int dq = nkv * G * 64, dkv = nkv * 64;
  const bool active = sub * 16 < nq;
  const int n_keys = p0 + nq;  // keys 0 .. p0+nq-1 exist for this tile
  const int n_blocks = (n_keys + BKV - 1) / BKV;
  const uint32_t smem_base = (uint32_t)__cvta_generic_to_shared(smem);
  const __nv_bfloat16 *kg = k_arena + base * dkv + kvh * 64;
  const __nv_bfloat16 *vg = v_arena + base * dkv + kvh * 64;

  auto issue_block = [&](int kb, int buf) {
    // 64 rows x 8 chunks for K and for V = 1024 16-byte copies, spread over the CTA; rows past n_keys are zero-filled
    for (int i = threadIdx.x; i < 2 * BKV * 8; i += blockDim.x) {
      const int which = i >> 9, r = (i >> 3) & 63, c = i & 7;
      const int key = kb * BKV + r;
      const __nv_bfloat16 *src = (which ? vg : kg) + (size_t)(key < n_keys ? key : 0) * dkv + c * 8;
      const uint32_t dst = smem_base + (uint32_t)(((buf * 2 + which) * BKV + r) * 128 + ((c ^ (r & 7)) << 4));
      cp_async16(dst, src, key < n_keys ? 16 : 0);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  // Q fragments (A operand), 4 k-steps of 16 dims; rows g and g+8 of this warp's 16-position sub-tile
  const int r_lo = sub * 16 + g, r_hi = r_lo + 8;
  uint32_t qa[4][4];
#pragma unroll
  for (int ks = 0; ks < 4; ks++) {
    const __nv_bfloat16 *ql = q + (size_t)(row0 + r_lo) * dq + head * 64 + ks * 16 + 2 * t;
    const __nv_bfloat16 *qh = q + (size_t)(row0 + r_hi) * dq + head * 64 + ks * 16 + 2 * t;
    qa[ks][0] = (active && r_lo < nq) ? *reinterpret_cast<const uint32_t *>(ql) : 0u;
    qa[ks][1] = (active && r_hi < nq) ? *reinterpret_cast<const uint32_t *>(qh) : 0u;
    qa[ks][2] = (active && r_lo < nq) ? *reinterpret_cast<const uint32_t *>(ql + 8) : 0u;
    qa[ks][3] = (active && r_hi < nq) ? *reinterpret_cast<const uint32_t *>(qh + 8) : 0u;
  }
  float o[8][4];
#pragma unroll
  for (int j = 0; j < 8; j++) o[j][0] = o[j][1] = o[j][2] = o[j][3] = 0.f;
  float m_lo = -INFINITY, m_hi = -INFINITY, l_lo = 0.f, l_hi = 0.f;
  const float c_log2 = 0.125f * 1.4426950408889634f;  // 1/sqrt(64) * log2(e)
  const int pos_lo = p0 + r_lo, pos_hi = p0 + r_hi;

  issue_block(0, 0);
  for (int kb = 0; kb < n_blocks; kb++) {
    const int buf = kb & 1;
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();  // block kb visible to everyone; everyone is done with the other buffer
    if (kb + 1 < n_blocks) issue_block(kb + 1, buf ^ 1);
    if (!active) continue;
    const uint32_t ks_base = smem_base + (uint32_t)((buf * 2 + 0) * BKV * 128);
    const uint32_t vs_base = smem_base + (uint32_t)((buf * 2 + 1) * BKV * 128);
    // ---- S = Q K^T ----
    float s[8][4];
#pragma unroll
    for (int j = 0; j < 8; j++) s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f;
#pragma unroll
    for (int ks = 0; ks < 4; ks++) {
#pragma unroll
      for (int j2 = 0; j2 < 4; j2++) {
        const int mtx = lane >> 3;
        const int row = 16 * j2 + (mtx >> 1) * 8 + (lane & 7);
        const int chunk = ks * 2 + (mtx & 1);
        uint32_t b0, b1, b2, b3;
        ldsm_x4(ks_base + (uint32_t)(row * 128 + ((chunk ^ (row & 7)) << 4)), b0, b1, b2, b3);
        mma_bf16(s[2 * j2], qa[ks], b0, b1);
        mma_bf16(s[2 * j2 + 1], qa[ks], b2, b3);
      }
    }
    // ---- scale, causal mask, online softmax ----
    const int key0 = kb * BKV + 2 * t;
    float mx_lo = m_lo, mx_hi = m_hi;
#pragma unroll
    for (int j = 0; j < 8; j++) {
      const int kx = key0 + 8 * j;
      s[j][0] = (kx <= pos_lo) ? s[j][0] * c_log2 : -INFINITY;
      s[j][1] = (kx + 1 <= pos_lo) ? s[j][1] * c_log2 : -INFINITY;
      s[j][2] = (kx <= pos_hi) ? s[j][2] * c_log2 : -INFINITY;
      s[j][3] = (kx + 1 <= pos_hi) ? s[j][3] * c_log2 : -INFINITY;
      mx_lo = fmaxf(mx_lo, fmaxf(s[j][0], s[j][1]));
      mx_hi = fmaxf(mx_hi, fmaxf(s[j][2], s[j][3]));
    }
    mx_lo = fmaxf(mx_lo, __shfl_xor_sync(0xffffffffu, mx_lo, 1));
    mx_lo = fmaxf(mx_lo, __shfl_xor_sync(0xffffffffu, mx_lo, 2));
    mx_hi = fmaxf(mx_hi, __shfl_xor_sync(0xffffffffu, mx_hi, 1));
    mx_hi = fmaxf(mx_hi, __shfl_xor_sync(0xffffffffu, mx_hi, 2));
    const float a_lo = ex2(m_lo - mx_lo), a_hi = ex2(m_hi - mx_hi);  // first block: ex2(-inf) = 0
    m_lo = mx_lo;
    m_hi = mx_hi;
    float sum_lo = 0.f, sum_hi = 0.f;
    uint32_t pa[4][4];
#pragma unroll
    for (int j = 0; j < 8; j++) {
      const float p0f = ex2(s[j][0] - mx_lo), p1f = ex2(s[j][1] - mx_lo);
      const float p2f = ex2(s[j][2] - mx_hi), p3f = ex2(s[j][3] - mx_hi);
      sum_lo += p0f;
      sum_lo += p1f;
      sum_hi += p2f;
      sum_hi += p3f;
      pa[j >> 1][(j & 1) * 2 + 0] = pack_bf16(p0f, p1f);
      pa[j >> 1][(j & 1) * 2 + 1] = pack_bf16(p2f, p3f);
    }
    l_lo = l_lo * a_lo + sum_lo;
    l_hi = l_hi * a_hi + sum_hi;
#pragma unroll
    for (int j = 0; j < 8; j++) {
      o[j][0] *= a_lo;
      o[j][1] *= a_lo;
      o[j][2] *= a_hi;
      o[j][3] *= a_hi;
    }
    // ---- O += P V ----
#pragma unroll
    for (int ks2 = 0; ks2 < 4; ks2++) {
#pragma unroll
      for (int dj2 = 0; dj2 < 4; dj2++) {
        const int mtx = lane >> 3;
        const int row = 16 * ks2 + (mtx & 1) * 8 + (lane & 7);
        const int chunk = 2 * dj2 + (mtx >> 1);
        uint32_t b0, b1, b2, b3;
        ldsm_x4_t(vs_base + (uint32_t)(row * 128 + ((chunk ^ (row & 7)) << 4)), b0, b1, b2, b3);
        mma_bf16(o[2 * dj2], pa[ks2], b0, b1);
        mma_bf16(o[2 * dj2 + 1], pa[ks2], b2, b3);
      }
    }
  }
  if (!active) return;
  l_lo += __shfl_xor_sync(0xffffffffu, l_lo, 1);
  l_lo += __shfl_xor_sync(0xffffffffu, l_lo, 2);
  l_hi += __shfl_xor_sync(0xffffffffu, l_hi, 1);
  l_hi += __shfl_xor_sync(0xffffffffu, l_hi, 2);
  const float inv_lo = 1.0f / l_lo, inv_hi = 1.0f / l_hi;
#pragma unroll
  for (int j = 0; j < 8; j++) {
    if (r_lo < nq)
      *reinterpret_cast<uint32_t *>(out + (size_t)(row0 + r_lo) * dq + head * 64 + 8 * j + 2 * t) = pack_bf16(o[j][0] * inv_lo, o[j][1] * inv_lo);
    if (r_hi < nq)
      *reinterpret_cast<uint32_t *>(out + (size_t)(row0 + r_hi) * dq + head * 64 + 8 * j + 2 * t) = pack_bf16(o[j][2] * inv_hi, o[j][3] * inv_hi);
  }
}

}  // namespace czk

namespace cz {

void attn_set_carveout() {
  cudaFuncSetAttribute(czk::attn_mma_kernel<1>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  cudaFuncSetAttribute(czk::attn_mma_kernel<2>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  cudaFuncSetAttribute(czk::attn_mma_kernel<3>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  cudaFuncSetAttribute(czk::attn_mma_kernel<4>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
}

int launch_attn_mma(cz_ctx *ctx, const __nv_bfloat16 *q, const __nv_bfloat16 *k_arena, const __nv_bfloat16 *v_arena, const int *pos,
                    const int *kv_base, const int *tile_row0, const int *tile_n, int n_tiles, __nv_bfloat16 *out, int nh, int nkv,
                    cudaStream_t st) {
  if (n_tiles == 0) return CZ_OK;
  const int G = nh / nkv;
  dim3 grid((unsigned)n_tiles, (unsigned)nkv);
#define CZ_ATT2_CASE(G_)                                                                                                             \
  case G_:                                                                                                                           \
    CZ_LAUNCH(ctx, CZ_K_ATTN,                                                                                                        \
              (czk::attn_mma_kernel<G_><<<grid, 128 * G_, 0, st>>>(q, k_arena, v_arena, pos, kv_base, tile_row0, tile_n, out, nkv))); \
    break;
  switch (G) {
    CZ_ATT2_CASE(1)
    CZ_ATT2_CASE(2)
    CZ_ATT2_CASE(3)
    CZ_ATT2_CASE(4)
    default:
      set_error("attention: unsupported GQA group size");
      return CZ_ERR_UNSUPPORTED;
  }
#undef CZ_ATT2_CASE
  CZ_CHECK_LAUNCH();
  return CZ_OK;
}

}  // namespace cz
