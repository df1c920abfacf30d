// K5, engine v3: causal GQA attention on the 5th-generation tensor cores (tcgen05.mma, accumulators in TMEM, operands fed by
// TMA), flash-style online softmax.  One CTA = one tile of up to 128 consecutive query positions of ONE sequence x one query
// head; two CTAs are resident per SM (112 KB shared memory, 256 TMEM columns each) so one CTA's softmax overlaps the other's MMAs.
//
// Replaces candle-transformers llama attention (repeat_kv + QK^T/sqrt(d) + f32 softmax + PV) reached from src/models.rs:94,110.
//
//   warp 4 (one elected thread): TMA loads (Q once; K block and V block [128 keys x 64 dims]; K single-, V double-buffered) and the MMAs:  S = Q K^T  (128 x 128 x 64: 4 UMMA k-steps)  and  O_blk = P V  (128 x 64 x 128: 8 k-steps)
//   warps 0-3: thread r owns query row r of the tile (tcgen05.ld 32x32b hands a TMEM lane to a thread): row max, exp2, row sum
//           need no shuffles; P goes back as bf16 through shared memory (K-major SWIZZLE_128B, the A operand of P V);
//           the running output row lives in 64 registers: O = O * alpha + O_blk.
// V stays in its natural row layout ([slot][kv dim]): the P V product takes it as an MN-major B operand (instruction
// descriptor bit 16), whose canonical SWIZZLE_128B shared-memory layout is exactly what TMA writes for a [128 keys][64 dims] box.
//
// ROW INVARIANCE (decode safety).  For a given (sequence, position, head) the arithmetic is a fixed sequence: keys in blocks of
// 128 anchored at key 0; per block one UMMA chain over the 64 dims, mask (exact -inf -> exp2 = 0), m' = max(m, rowmax),
// p = exp2(s*c - m'), l = l*a + sum(p) in column order, O = O*a + (bf16(p) V by one UMMA chain over the block's 128 keys).
// Nothing depends on which other rows share the tile or on the tile's size, so the same kernel serves 128-row teacher-forced
// tiles and single-row decode tiles and gives bit-identical outputs (tests: stepwise == teacher-forced bitwise).
#include <cuda.h>

#include "cz_common.cuh"
#include "llama_kernels.h"
#include "tc_ptx.cuh"

namespace czk {

constexpr int AT_THREADS = 160;
// shared memory: Q 16 KB | K 16 KB (single buffer: it is free again as soon as S = Q K^T has been computed, long before the
// next block needs it) | V^T 2 x 16 KB | P 32 KB | barriers; 1 KB of slack to align the SWIZZLE_128B tiles.  ~98 KB: two CTAs per SM.
constexpr int AT_Q = 0, AT_K = 16384, AT_V = 32768, AT_P = 65536, AT_BAR = 98304, AT_SMEM = AT_BAR + 128 + 1024;

__device__ __forceinline__ float at_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap *map, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst),
      "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

// STACKED = false: tile rows are up to 128 consecutive positions of one query head (grid.y = query heads).
// STACKED = true (stepwise decode, one position per sequence): tile rows are the G query heads that share one KV head, all at
//   the same position, and ONE CTA walks all KV heads of the sequence back to back (grid.y = 1): the per-CTA set-up (TMEM
//   allocation, barriers, descriptor prefetch) is paid once per sequence and the next head's Q / K / V loads are already in
//   flight while the current head is processed.
// A (position, head) row goes through exactly the same arithmetic in both modes.
template <bool STACKED>
__global__ void __launch_bounds__(AT_THREADS) attn_tc_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                                                             const __grid_constant__ CUtensorMap tm_v, const int *__restrict__ pos,
                                                             const int *__restrict__ kv_base, const int *__restrict__ tile_row0,
                                                             const int *__restrict__ tile_n, __nv_bfloat16 *__restrict__ out, int nh, int nkv) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int G = nh / nkv;
  // One CTA walks n_loop (head, key block) sequences back to back with ONE flattened load / MMA pipeline, so the per-CTA set-up
  // is amortised and the next head's loads are in flight while the current one is processed:
  //   STACKED:  all nkv KV heads of the sequence (grid.y = 1);  else: the G query heads of KV head blockIdx.y (grid.y = nkv).
  const int tile = blockIdx.x, kvh0 = STACKED ? 0 : (int)blockIdx.y, head0 = STACKED ? 0 : (int)blockIdx.y * G;
  const int n_loop = STACKED ? nkv : G;
  const int row0 = tile_row0[tile], n_pos = STACKED ? 1 : tile_n[tile];
  const int nq = STACKED ? G : n_pos;  // valid tile rows
  const int p0 = pos[row0], base = kv_base[row0];
  const int nb = (p0 + n_pos + 127) >> 7;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint64_t *bars = reinterpret_cast<uint64_t *>(smem + AT_BAR);
  const uint32_t q_full = smem_u32(&bars[0]), k_full = smem_u32(&bars[1]), k_empty = smem_u32(&bars[2]), v_full = smem_u32(&bars[3]) /*[2]*/,
                 v_empty = smem_u32(&bars[5]) /*[2]*/, s_full = smem_u32(&bars[7]), p_ready = smem_u32(&bars[8]), o_full = smem_u32(&bars[9]);
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(&bars[10]);
  const uint32_t sQ = smem_u32(smem + AT_Q), sK = smem_u32(smem + AT_K), sV = smem_u32(smem + AT_V), sP = smem_u32(smem + AT_P);

  if (warp == 4 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_q) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_k) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_v) : "memory");
    mbar_init(q_full, 1);
    mbar_init(k_full, 1);
    mbar_init(k_empty, 1);
    mbar_init(v_full, 1);
    mbar_init(v_full + 8, 1);
    mbar_init(v_empty, 1);
    mbar_init(v_empty + 8, 1);
    mbar_init(s_full, 1);
    mbar_init(p_ready, 128);
    mbar_init(o_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 4) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(256) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 4) {
    if (lane == 0) {
      const int total = n_loop * nb;  // flattened (head, key block) iterations; every barrier's phase follows this counter
      auto load_k = [&](int g) {
        mbar_expect_tx(k_full, 16384);
        tma_load_2d(sK, &tm_k, k_full, (STACKED ? g / nb : kvh0) * 64, base + (g % nb) * 128);
      };
      auto load_v = [&](int g, int st) {
        const uint32_t fb = v_full + 8 * st;
        mbar_expect_tx(fb, 16384);
        tma_load_2d(sV + st * 16384, &tm_v, fb, (STACKED ? g / nb : kvh0) * 64, base + (g % nb) * 128);  // [128 keys][64 dims], like the K block
      };
      // q viewed as [row][head][64]: a box of 128 rows x 1 head, or 1 row x G heads
      auto load_q = [&](int hh) {
        mbar_expect_tx(q_full, STACKED ? G * 128 : 16384);
        tma_load_3d(sQ, &tm_q, q_full, 0, STACKED ? hh * G : head0 + hh, row0);
      };
      load_q(0);
      load_k(0);
      load_v(0, 0);
      if (total > 1) load_v(1, 1);
      // P V: the B operand V [128 keys][64 dims] has the dims (N) contiguous -> MN-major B (instruction descriptor bit 16)
      constexpr uint32_t idesc_qk = make_idesc_mn(128, 128), idesc_pv = make_idesc_mn(128, 64) | (1u << 16);
      const uint64_t q_desc = make_kmajor_sw128_desc(sQ), k_desc = make_kmajor_sw128_desc(sK);
      auto issue_qk = [&]() {
#pragma unroll
        for (int ks = 0; ks < 4; ks++) tc_mma_bf16(tmem_base, q_desc + (uint64_t)(ks * 2), k_desc + (uint64_t)(ks * 2), idesc_qk, ks ? 1u : 0u);
        tc_commit(s_full);
        tc_commit(k_empty);
      };
      mbar_wait(q_full, 0);
      mbar_wait(k_full, 0);
      tc_fence_after();
      issue_qk();
      for (int g = 0; g < total; g++) {
        const int st = g & 1;
        const bool next_head = (g + 1) % nb == 0;  // iteration g + 1 starts a new KV head
        if (g + 1 < total) {
          mbar_wait(k_empty, g & 1);  // S(g) = Q K^T is done: the K buffer (and, at a head boundary, Q) can be refilled while the softmax runs
          if (next_head) load_q((g + 1) / nb);
          load_k(g + 1);
        }
        mbar_wait(p_ready, g & 1);  // P(g) is in shared memory and S has been consumed
        tc_fence_after();
        if (g + 1 < total) {
          if (next_head) mbar_wait(q_full, ((g + 1) / nb) & 1);
          mbar_wait(k_full, (g + 1) & 1);
          tc_fence_after();
          issue_qk();  // S(g+1): overlaps P V (g) and the softmax warps' output update
        }
        mbar_wait(v_full + 8 * st, (g >> 1) & 1);
        tc_fence_after();
#pragma unroll
        for (int ks = 0; ks < 8; ks++) {
          const uint64_t p_desc = make_kmajor_sw128_desc(sP + (ks >> 2) * 16384) + (uint64_t)((ks & 3) * 2);
          // canonical MN-major SWIZZLE_128B layout ((8,n),(8,k)):((1,LBO),(8,SBO)) in 16-byte units: one 128-byte row per key,
          // 8-key groups 1024 B apart (SBO), n = 1; a K-step of 16 keys advances the start address by 16 rows = 2048 B
          const uint64_t v_desc = make_kmajor_sw128_desc(sV + st * 16384) + (uint64_t)(ks * 128);
          tc_mma_bf16(tmem_base + 128, p_desc, v_desc, idesc_pv, ks ? 1u : 0u);
        }
        tc_commit(o_full);
        tc_commit(v_empty + 8 * st);
        if (g + 2 < total) {
          mbar_wait(v_empty + 8 * st, (g >> 1) & 1);
          load_v(g + 2, st);
        }
      }
    }
  } else {
    const int r = warp * 32 + lane;
    const bool warp_valid = warp * 32 < nq;
    const int pos_r = STACKED ? p0 : p0 + r;
    const uint32_t t_s = tmem_base + ((uint32_t)(warp * 32) << 16), t_o = t_s + 128;
    const float c_log2 = 0.125f * 1.4426950408889634f;  // 1/sqrt(64) * log2(e)
    float m = -INFINITY, l = 0.f, alpha_prev = 1.f;
    bool pending = false;  // the previous block's O update (O = O * alpha + P V) is still owed
    float o[64];
    const int total = n_loop * nb;
    // O = O * a + (P V of iteration gi), read from TMEM once that product has completed
    auto o_update = [&](int gi, float a) {
      mbar_wait(o_full, gi & 1);
      tc_fence_after();
      if (warp_valid) {
        uint32_t v[64];
        tc_ld_32x32(t_o, *reinterpret_cast<uint32_t(*)[32]>(&v[0]));
        tc_ld_32x32(t_o + 32, *reinterpret_cast<uint32_t(*)[32]>(&v[32]));
        tc_ld_wait();
#pragma unroll
        for (int j = 0; j < 64; j++) o[j] = fmaf(o[j], a, __uint_as_float(v[j]));
      }
      tc_fence_before();
    };
    for (int g = 0; g < total; g++) {
      const int kb = g % nb;
      if (kb == 0) {  // new head: fresh online-softmax state
        m = -INFINITY;
        l = 0.f;
#pragma unroll
        for (int j = 0; j < 64; j++) o[j] = 0.f;
      }
      mbar_wait(s_full, g & 1);
      tc_fence_after();
      float alpha = 1.f, mx = 0.f;
      const int lim = pos_r - kb * 128;
      const bool need_mask = kb * 128 + 127 > p0;
      if (warp_valid) {
        // causal mask: columns > lim of this block are keys after the row's position.  Blocks entirely at or before the
        // tile's first position take the mask-free path (warp-uniform); both paths compute identical values for unmasked keys.
        // four independent running maxima / sums (combined in a fixed order): a single 128-long dependent chain of
        // FMNMX / FADD would cost more cycles than the exp2 work itself
        float r4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll 1
        for (int c = 0; c < 2; c++) {
          uint32_t v[64];
          tc_ld_32x32(t_s + (uint32_t)(c * 64), *reinterpret_cast<uint32_t(*)[32]>(&v[0]));
          tc_ld_32x32(t_s + (uint32_t)(c * 64 + 32), *reinterpret_cast<uint32_t(*)[32]>(&v[32]));
          tc_ld_wait();
          if (need_mask) {
#pragma unroll
            for (int j = 0; j < 64; j++) r4[j & 3] = fmaxf(r4[j & 3], (c * 64 + j <= lim) ? __uint_as_float(v[j]) : -INFINITY);
          } else {
#pragma unroll
            for (int j = 0; j < 64; j++) r4[j & 3] = fmaxf(r4[j & 3], __uint_as_float(v[j]));
          }
        }
        const float raw = fmaxf(fmaxf(r4[0], r4[1]), fmaxf(r4[2], r4[3]));
        mx = fmaxf(m, raw * c_log2);
        alpha = at_ex2(m - mx);  // first block: ex2(-inf) = 0
        m = mx;
      }
      // The previous block's output update is applied HERE, after this block's max pass: its P V product ran while the max was
      // being computed, so its latency is hidden.  It must precede pass 2, which overwrites the P buffer that product read.
      if (pending) {
        o_update(g - 1, alpha_prev);
        pending = false;
      }
      if (warp_valid) {
        float s4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 1
        for (int c = 0; c < 4; c++) {
          uint32_t v[32];
          tc_ld_32x32(t_s + (uint32_t)(c * 32), v);
          tc_ld_wait();
          uint32_t pk[16];
#pragma unroll
          for (int j = 0; j < 32; j += 2) {
            float pa = at_ex2(fmaf(__uint_as_float(v[j]), c_log2, -mx)), pb = at_ex2(fmaf(__uint_as_float(v[j + 1]), c_log2, -mx));
            if (need_mask) {
              pa = (c * 32 + j <= lim) ? pa : 0.f;
              pb = (c * 32 + j + 1 <= lim) ? pb : 0.f;
            }
            s4[j & 2] += pa;
            s4[(j & 2) + 1] += pb;
            __nv_bfloat162 h = __floats2bfloat162_rn(pa, pb);
            pk[j >> 1] = *reinterpret_cast<uint32_t *>(&h);
          }
          // row r, keys [c*32, c*32+32): 16-byte chunks (c&1)*4 .. +3 of the row's 128-byte line in K-atom (c>>1)
          const uint32_t rowp = sP + (uint32_t)((c >> 1) * 16384 + r * 128);
#pragma unroll
          for (int q = 0; q < 4; q++) {
            const uint32_t dst = rowp + (uint32_t)(((((c & 1) * 4 + q) ^ (r & 7))) << 4);
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(pk[4 * q]), "r"(pk[4 * q + 1]), "r"(pk[4 * q + 2]),
                         "r"(pk[4 * q + 3])
                         : "memory");
          }
        }
        l = fmaf(l, alpha, (s4[0] + s4[1]) + (s4[2] + s4[3]));
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      tc_fence_before();
      mbar_arrive(p_ready);
      if (kb == nb - 1) {
        o_update(g, alpha);  // last block of the head: nothing left to hide behind
      } else {
        pending = true;
        alpha_prev = alpha;
      }
      if (kb == nb - 1 && r < nq) {  // head finished: normalise and store its output row
        const float inv = 1.0f / l;
        const int head = STACKED ? (g / nb) * G + r : head0 + g / nb;
        __nv_bfloat16 *dst = out + (size_t)(STACKED ? row0 : row0 + r) * (nh * 64) + head * 64;
#pragma unroll
        for (int j = 0; j < 64; j += 8) {
          uint32_t w[4];
#pragma unroll
          for (int e = 0; e < 4; e++) {
            __nv_bfloat162 h = __floats2bfloat162_rn(o[j + 2 * e] * inv, o[j + 2 * e + 1] * inv);
            w[e] = *reinterpret_cast<uint32_t *>(&h);
          }
          *reinterpret_cast<uint4 *>(dst + j) = make_uint4(w[0], w[1], w[2], w[3]);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    __syncwarp();
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(256) : "memory");
  }
}

}  // namespace czk

namespace cz {

int make_map_bf16(CUtensorMap *map, const void *ptr, int rows, int K, int ld_elems, int box_rows);  // gemm_tcgen05.cu

typedef CUresult (*EncodeTiledFn3)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                   const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// q [n_rows][nh][64] bf16 as a 3D tensor, box = 64 x box_heads x box_rows, 128-byte swizzle
static int make_q_map(CUtensorMap *map, const void *q, int n_rows, int nh, int box_heads, int box_rows) {
  static EncodeTiledFn3 fn = nullptr;
  if (!fn) {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult qr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qr) != cudaSuccess || qr != cudaDriverEntryPointSuccess) {
      set_error("cuTensorMapEncodeTiled entry point not available");
      return CZ_ERR_CUDA;
    }
    fn = (EncodeTiledFn3)p;
  }
  cuuint64_t gdim[3] = {64, (cuuint64_t)nh, (cuuint64_t)n_rows};
  cuuint64_t gstride[2] = {128, (cuuint64_t)nh * 128};
  cuuint32_t box[3] = {64, (cuuint32_t)box_heads, (cuuint32_t)box_rows};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void *>(q), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled (q) failed with CUresult " + std::to_string((int)r));
    return CZ_ERR_CUDA;
  }
  return CZ_OK;
}

// q [n_rows][nh*64] bf16; k_arena, v_arena [n_slots][nkv*64] bf16.
// single_rows: every tile is one position (stepwise decode) -> the GQA group is stacked into one CTA.
int launch_attn_tc(cz_ctx *ctx, const __nv_bfloat16 *q, int n_rows, const __nv_bfloat16 *k_arena, const __nv_bfloat16 *v_arena, int n_slots,
                   int /*unused*/, const int *pos, const int *kv_base, const int *tile_row0, const int *tile_n, int n_tiles, __nv_bfloat16 *out,
                   int nh, int nkv, bool single_rows, cudaStream_t st) {
  if (n_tiles == 0) return CZ_OK;
  static bool attr = false;
  if (!attr) {
    CZ_CUDA_TRY(cudaFuncSetAttribute(czk::attn_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, czk::AT_SMEM));
    CZ_CUDA_TRY(cudaFuncSetAttribute(czk::attn_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, czk::AT_SMEM));
    attr = true;
  }
  CUtensorMap tq, tk, tv;
  CZ_TRY(make_q_map(&tq, q, n_rows, nh, single_rows ? nh / nkv : 1, single_rows ? 1 : 128));
  CZ_TRY(make_map_bf16(&tk, k_arena, n_slots, nkv * 64, nkv * 64, 128));
  CZ_TRY(make_map_bf16(&tv, v_arena, n_slots, nkv * 64, nkv * 64, 128));  // V rows [slot][nkv*64], same box as K
  if (single_rows) {
    dim3 grid((unsigned)n_tiles, 1);
    CZ_LAUNCH(ctx, CZ_K_ATTN,
              (czk::attn_tc_kernel<true><<<grid, czk::AT_THREADS, czk::AT_SMEM, st>>>(tq, tk, tv, pos, kv_base, tile_row0, tile_n, out, nh, nkv)));
  } else {
    dim3 grid((unsigned)n_tiles, (unsigned)nkv);
    CZ_LAUNCH(ctx, CZ_K_ATTN,
              (czk::attn_tc_kernel<false><<<grid, czk::AT_THREADS, czk::AT_SMEM, st>>>(tq, tk, tv, pos, kv_base, tile_row0, tile_n, out, nh, nkv)));
  }
  CZ_CHECK_LAUNCH();
  return CZ_OK;
}

}  // namespace cz
