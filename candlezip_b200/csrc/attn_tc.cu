// K5, engine v4: causal GQA attention on the 5th-generation tensor cores (tcgen05.mma, accumulators in TMEM, operands fed by
// TMA), flash-style online softmax.  One CTA = one tile of up to 128 consecutive query positions of ONE sequence x one KV head
// (it walks the G query heads of that group); two CTAs are resident per SM (64 KB shared memory, 256 TMEM columns each) so one
// CTA's softmax overlaps the other's MMAs.
//
// Replaces candle-transformers llama attention (repeat_kv + QK^T/sqrt(d) + f32 softmax + PV) reached from src/models.rs:94,110.
//
//   warp 4 (one elected thread): TMA loads (Q per head; K block and V block [128 keys x 64 dims]; K single-, V double-buffered)
//           and the MMAs:  S = Q K^T  (128 x 128 x 64: 4 UMMA k-steps, both operands in shared memory)  and
//           O += P V  (128 x 64 x 128: 8 k-steps, A = P read from TENSOR MEMORY, B = V in shared memory, accumulating in TMEM
//           across the key blocks of a head).
//   warps 0-3: thread r owns query row r of the tile (tcgen05.ld 32x32b hands a TMEM lane to a thread).  S is read from TMEM
//           ONCE into 128 registers (the S columns are handed back to the MMA warp right away, so S(g+1) = Q K^T runs under
//           the softmax of block g); row max (3-input FMNMX), p = exp2(s*c - m) (packed FFMA2), row sum (packed FADD2); P goes
//           back as packed bf16 with tcgen05.st into 64 TMEM columns -- no shared-memory round trip, no proxy fence.
//           The output row is NOT kept in registers: it stays in TMEM and is rescaled there only when a row's running max
//           grows by more than 2^8 (lazy rescale, below); the only regular TMEM read per block is S itself
//           (tcgen05.ld moves 64 B/clk: re-reading S for a second pass and reading O_blk every block was what bound v3).
// V stays in its natural row layout ([slot][kv dim]): the P V product takes it as an MN-major B operand (instruction
// descriptor bit 16), whose canonical SWIZZLE_128B shared-memory layout is exactly what TMA writes for a [128 keys][64 dims] box.
// TMEM columns (per CTA, 256): S [0,128) fp32 | O [128,192) fp32 | P [192,256) bf16 pairs (A operand, K-major: lane = row,
// one 32-bit column = two consecutive keys; a 16-key UMMA k-step is 8 columns).
//
// ROW INVARIANCE (decode safety).  For a given (sequence, position, head) the arithmetic is a fixed sequence: keys in blocks of
// 128 anchored at key 0; per block one UMMA chain over the 64 dims, mask (exact -inf -> exp2 = 0), row max, then
//   if (rowmax*c > m + 8) { a = exp2(m - rowmax*c); m = rowmax*c; l *= a; O *= a; }     (decision from the row's own data only)
//   p = exp2(s*c - m);  l += sum(p) in a fixed order;  O += bf16(p) V  (one UMMA chain over the block's 128 keys, accumulated in TMEM).
// The O rescale is a warp-collective TMEM read-modify-write performed when ANY row of the warp asks for it; rows that did not ask
// are multiplied by exactly 1.0f, so what the other rows of a tile do never changes a row's bits.  Nothing depends on which other rows share the tile or
// on the tile's size, so the same kernel serves 128-row teacher-forced tiles and single-row decode tiles and gives bit-identical
// outputs (tests: stepwise == teacher-forced bitwise).
#include <cuda.h>

#include "cz_common.cuh"
#include "llama_kernels.h"
#include "tc_ptx.cuh"

namespace czk {

constexpr int AT_THREADS = 256;  // warpgroup 0: four softmax warps; warpgroup 1: warp 4 = loader + MMA issuer, warps 5-7 only donate registers
// shared memory: Q 2 x 16 KB | K 2 x 16 KB | V 2 x 16 KB | barriers | item ring; 1 KB of slack to align the SWIZZLE_128B tiles.
// ~98 KB: two CTAs per SM (the TMEM columns, 2 x 256, are what limits residency).
constexpr int AT_Q = 0, AT_K = 32768, AT_V = 65536, AT_BAR = 98304, AT_RING = AT_BAR + 256, AT_SMEM = AT_RING + 4 * 32 + 1024;
constexpr int AT_TM_S = 0, AT_TM_O = 128, AT_TM_P = 192;
constexpr float AT_RESCALE_LOG2 = 8.0f;  // lazy-rescale threshold on the log2-domain running max

__device__ __forceinline__ float at_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float at_max3(float a, float b, float c) {
  float d;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}
// (x0, x1) * c + nm on the packed-f32 FMA path (FFMA2): two IEEE fmas, bit-identical to two scalar fmaf
__device__ __forceinline__ void at_fma2(float &x0, float &x1, uint64_t c2, uint64_t nm2) {
  uint64_t x, y;
  asm("mov.b64 %0, {%1, %2};" : "=l"(x) : "f"(x0), "f"(x1));
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(y) : "l"(x), "l"(c2), "l"(nm2));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(x0), "=f"(x1) : "l"(y));
}
__device__ __forceinline__ void at_unpack2(uint64_t x, float &a, float &b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(x)); }
__device__ __forceinline__ uint64_t at_pack2(float a, float b) {
  uint64_t x;
  asm("mov.b64 %0, {%1, %2};" : "=l"(x) : "f"(a), "f"(b));
  return x;
}
__device__ __forceinline__ uint64_t at_add2(uint64_t a, uint64_t b) {
  uint64_t y;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(y) : "l"(a), "l"(b));
  return y;
}

// exp2 of a PAIR on the FMA pipe (no MUFU): round-to-nearest split x = n + f, f in [-0.5, 0.5] (magic-number add), degree-3
// minimax polynomial for 2^f (max relative error 7.5e-5, far below the bf16 rounding of P), 2^n applied by adding n to the
// exponent field.  The XU pipe (16 ex2 / clk / SM) is what bounds the softmax, so a fixed subset of the columns -- pairs
// AT_POLY_MASK of every 8 pairs, the same for every row, tile and mode -- takes this path instead (as FlashAttention-4 does).
constexpr uint32_t AT_POLY_MASK = 0x52;  // pairs 1, 4, 6 of every 8: 3/8 of the exponentials
__device__ __forceinline__ void at_ex2_poly2(float &x0, float &x1) {
  x0 = fmaxf(x0, -125.f);  // keeps 2^n a normal number; a masked key (-inf) becomes 2^-125, far below one ulp of l and of any O sum
  x1 = fmaxf(x1, -125.f);
  const uint64_t x = at_pack2(x0, x1);
  const uint64_t magic = at_pack2(12582912.f, 12582912.f), nmagic = at_pack2(-12582912.f, -12582912.f), mone = at_pack2(-1.f, -1.f);
  const uint64_t c3 = at_pack2(0x1.c3f76p-5f, 0x1.c3f76p-5f), c2 = at_pack2(0x1.f0de1ap-3f, 0x1.f0de1ap-3f),
                 c1 = at_pack2(0x1.62f31ap-1f, 0x1.62f31ap-1f), c0 = at_pack2(0x1.fff692p-1f, 0x1.fff692p-1f);
  const uint64_t t = at_add2(x, magic);  // low mantissa bits = round(x)
  const uint64_t n = at_add2(t, nmagic);
  uint64_t f, q;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(f) : "l"(n), "l"(mone), "l"(x));
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(q) : "l"(f), "l"(c3), "l"(c2));
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(q) : "l"(q), "l"(f), "l"(c1));
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(q) : "l"(q), "l"(f), "l"(c0));
  float t0, t1, q0, q1;
  at_unpack2(t, t0, t1);
  at_unpack2(q, q0, q1);
  x0 = __int_as_float(__float_as_int(q0) + (__float_as_int(t0) << 23));
  x1 = __int_as_float(__float_as_int(q1) + (__float_as_int(t1) << 23));
}

__device__ __forceinline__ void tc_st_32x32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]),
      "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]),
      "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tc_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// D[tmem] (+)= A[tmem] * B[smem]: the A operand (P) is read from tensor memory
__device__ __forceinline__ void tc_mma_bf16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n"
      "}\n" ::"r"(d_tmem),
      "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(acc)
      : "memory");
}

__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap *map, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst),
      "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

// Work items.  STACKED = false: item = (tile of up to 128 consecutive positions of one sequence, KV head); the CTA walks the G
//   query heads of that KV head, so a K / V block is fetched once per group.
// STACKED = true (stepwise decode, one position per sequence): item = one sequence; tile rows are the G query heads that share
//   one KV head, all at the same position, and the CTA walks all nkv KV heads back to back.
// A (position, head) row goes through exactly the same arithmetic in both modes.
//
// PERSISTENT: the grid is 2 CTAs per SM; each CTA pulls items from a global counter and runs them through ONE flattened
// (item, head, key block) pipeline: TMEM allocation / barrier set-up is paid once per CTA, the next item's Q / K / V loads and its
// first S = Q K^T are in flight while the softmax warps finish the current item, and a head's output rows are read out of TMEM
// one iteration late (when the next iteration has to wait for that P V product anyway), so no warp ever idles on a P V.
struct AtItem {  // published by the loader thread, one per item (ring of 4)
  int row0, n_pos, p0, base, nb, head0, valid, pad;
};

template <bool STACKED>
__global__ void __launch_bounds__(AT_THREADS, 2) attn_tc_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                                                                const __grid_constant__ CUtensorMap tm_v, const int *__restrict__ pos,
                                                                const int *__restrict__ kv_base, const int *__restrict__ tile_row0,
                                                                const int *__restrict__ tile_n, __nv_bfloat16 *__restrict__ out, int nh, int nkv,
                                                                int n_tiles, int *__restrict__ work_counter) {
  // The kernel has no static shared memory, so the dynamic window starts at shared address 0 and the declared 1024-byte
  // alignment (SWIZZLE_128B tiles) holds; every barrier / tile address is then a compile-time constant instead of a value the
  // compiler re-derives from S2R each iteration.  Checked once below.
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  const int G = nh / nkv;
  const int n_loop = STACKED ? nkv : G;            // heads walked per item
  const int n_items = STACKED ? n_tiles : n_tiles * nkv;
  const int nq_stacked = G;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint64_t *bars = reinterpret_cast<uint64_t *>(smem + AT_BAR);
  const uint32_t q_full = smem_u32(&bars[0]) /*[2]*/, k_full = smem_u32(&bars[2]) /*[2]*/, k_empty = smem_u32(&bars[4]) /*[2]*/,
                 v_full = smem_u32(&bars[6]) /*[2]*/, v_empty = smem_u32(&bars[8]) /*[2]*/, s_full = smem_u32(&bars[10]),
                 p_ready = smem_u32(&bars[11]), o_full = smem_u32(&bars[12]), s_free = smem_u32(&bars[13]), item_ready = smem_u32(&bars[14]),
                 item_taken = smem_u32(&bars[15]);
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(&bars[16]);
  volatile AtItem *ring = reinterpret_cast<volatile AtItem *>(smem + AT_RING);
  const uint32_t sQ = smem_u32(smem + AT_Q), sK = smem_u32(smem + AT_K), sV = smem_u32(smem + AT_V);

  if (warp == 4 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_q) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_k) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_v) : "memory");
    mbar_init(q_full, 1);
    mbar_init(q_full + 8, 1);
    mbar_init(k_full, 1);
    mbar_init(k_full + 8, 1);
    mbar_init(k_empty, 1);
    mbar_init(k_empty + 8, 1);
    mbar_init(v_full, 1);
    mbar_init(v_full + 8, 1);
    mbar_init(v_empty, 1);
    mbar_init(v_empty + 8, 1);
    mbar_init(s_full, 1);
    mbar_init(p_ready, 128);
    mbar_init(o_full, 1);
    mbar_init(s_free, 128);
    mbar_init(item_ready, 1);
    mbar_init(item_taken, 128);
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

  // Register split (setmaxnreg is a warpgroup-wide instruction, hence the padded second warpgroup): the launch gives every
  // thread 128 registers (2 CTAs x 256 threads); warpgroup 1 keeps 40 and the softmax warps -- which hold a whole 128-column S
  // row per thread -- grow to 216.
  if (warp >= 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 216;");
  }
  if (warp >= 4) {
    if (warp == 4 && lane == 0) {
      // cursor over the flattened (item, head, key block) iteration space
      struct Cur {
        int row0, base, nb, head0, kvh, hh, kb;
        bool valid;
      };
      int n_pub = 0;  // items published so far
      auto fetch_item = [&](Cur &c) {
        const int item = atomicAdd(work_counter, 1);
        // never more than one unconsumed descriptor: item_ready's phase parity stays unambiguous for the waiting warps
        if (n_pub >= 1) mbar_wait(item_taken, (n_pub - 1) & 1);
        volatile AtItem *slot = &ring[n_pub & 3];
        if (item >= n_items) {
          c.valid = false;
          slot->valid = 0;
        } else {
          const int tile = STACKED ? item : item / nkv;
          c.kvh = STACKED ? 0 : item - tile * nkv;
          c.row0 = tile_row0[tile];
          const int n_pos = STACKED ? 1 : tile_n[tile];
          const int p0 = pos[c.row0];
          c.base = kv_base[c.row0];
          c.nb = (p0 + n_pos + 127) >> 7;
          c.head0 = STACKED ? 0 : c.kvh * G;
          c.hh = 0;
          c.kb = 0;
          c.valid = true;
          slot->row0 = c.row0;
          slot->n_pos = n_pos;
          slot->p0 = p0;
          slot->base = c.base;
          slot->nb = c.nb;
          slot->head0 = c.head0;
          slot->valid = 1;
        }
        n_pub++;
        mbar_arrive(item_ready);  // release: the slot's contents are visible to whoever observes this phase
      };
      auto advance = [&](Cur &c) {
        if (++c.kb == c.nb) {
          c.kb = 0;
          if (++c.hh == n_loop) fetch_item(c);
        }
      };
      // Q and K are double-buffered (stage = head count & 1 / iteration & 1), V double-buffered: K(it+2) and, at a head
      // boundary, Q are requested two iterations ahead so that S(it+1) = Q K^T never waits for a TMA round trip.
      auto load_k = [&](const Cur &c, int st) {
        const uint32_t fb = k_full + 8 * st;
        mbar_expect_tx(fb, 16384);
        tma_load_2d(sK + st * 16384, &tm_k, fb, (STACKED ? c.hh : c.kvh) * 64, c.base + c.kb * 128);
      };
      auto load_v = [&](const Cur &c, int st) {
        const uint32_t fb = v_full + 8 * st;
        mbar_expect_tx(fb, 16384);
        tma_load_2d(sV + st * 16384, &tm_v, fb, (STACKED ? c.hh : c.kvh) * 64, c.base + c.kb * 128);  // [128 keys][64 dims], like the K block
      };
      // q viewed as [row][head][64]: a box of 128 rows x 1 head, or 1 row x G heads
      int n_q_loaded = 0, n_q_used = 0;  // heads whose Q tile has been requested / consumed by a first Q K^T
      auto load_q = [&](const Cur &c) {
        const int st = n_q_loaded & 1;
        const uint32_t fb = q_full + 8 * st;
        mbar_expect_tx(fb, STACKED ? G * 128 : 16384);
        tma_load_3d(sQ + st * 16384, &tm_q, fb, 0, STACKED ? c.hh * G : c.head0 + c.hh, c.row0);
        n_q_loaded++;
      };
      // P V: the B operand V [128 keys][64 dims] has the dims (N) contiguous -> MN-major B (instruction descriptor bit 16)
      constexpr uint32_t idesc_qk = make_idesc_mn(128, 128), idesc_pv = make_idesc_mn(128, 64) | (1u << 16);
      // S(j) = Q K^T for iteration j (cursor c): waits for its operands, issues, and signals s_full / k_empty[j & 1]
      auto issue_qk = [&](const Cur &c, int j) {
        if (c.kb == 0) {  // first block of a head: its Q tile
          mbar_wait(q_full + 8 * (n_q_used & 1), (n_q_used >> 1) & 1);
          n_q_used++;
        }
        const int qs = (n_q_used - 1) & 1, ks_ = j & 1;
        mbar_wait(k_full + 8 * ks_, (j >> 1) & 1);
        tc_fence_after();
        const uint64_t q_desc = make_kmajor_sw128_desc(sQ + qs * 16384), k_desc = make_kmajor_sw128_desc(sK + ks_ * 16384);
#pragma unroll
        for (int ks = 0; ks < 4; ks++)
          tc_mma_bf16(tmem_base + AT_TM_S, q_desc + (uint64_t)(ks * 2), k_desc + (uint64_t)(ks * 2), idesc_qk, ks ? 1u : 0u);
        tc_commit(s_full);
        tc_commit(k_empty + 8 * ks_);
      };
      Cur n0;
      fetch_item(n0);
      if (n0.valid) {
        Cur n1 = n0;
        advance(n1);
        load_q(n0);
        load_k(n0, 0);
        load_v(n0, 0);
        if (n1.valid) {
          if (n1.kb == 0) load_q(n1);
          load_k(n1, 1);
        }
        issue_qk(n0, 0);
        for (int it = 0;; it++) {  // `it` counts iterations across items: every barrier's phase follows it
          const Cur cur = n0;
          const int st = it & 1;
          if (n1.valid) {
            mbar_wait(s_free, it & 1);  // the softmax warps hold S(it) in registers
            issue_qk(n1, it + 1);       // S(it+1): runs under the softmax of iteration it
          }
          Cur n2 = n1;
          if (n1.valid) advance(n2);
          if (n1.valid && n2.valid) {
            // S(it) is done (and with it every earlier Q K^T): K stage it & 1 and the Q stage of the head before last are free
            mbar_wait(k_empty + 8 * st, (it >> 1) & 1);
            if (n2.kb == 0) load_q(n2);
            load_k(n2, st);
          }
          if (n1.valid) {
            if (it >= 1) mbar_wait(v_empty + 8 * (st ^ 1), ((it - 1) >> 1) & 1);  // P V (it-1) is done: its V stage takes block it + 1
            load_v(n1, st ^ 1);
          }
          mbar_wait(p_ready, it & 1);  // P(it) is in tensor memory; O has been rescaled / the previous head's O has been read out
          mbar_wait(v_full + 8 * st, (it >> 1) & 1);
          tc_fence_after();
          const uint32_t first = cur.kb == 0 ? 0u : 1u;  // a head's first block overwrites O
#pragma unroll
          for (int ks = 0; ks < 8; ks++) {
            // canonical MN-major SWIZZLE_128B layout ((8,n),(8,k)):((1,LBO),(8,SBO)) in 16-byte units: one 128-byte row per key,
            // 8-key groups 1024 B apart (SBO), n = 1; a K-step of 16 keys advances the start address by 16 rows = 2048 B
            const uint64_t v_desc = make_kmajor_sw128_desc(sV + st * 16384) + (uint64_t)(ks * 128);
            tc_mma_bf16_ts(tmem_base + AT_TM_O, tmem_base + AT_TM_P + (uint32_t)(ks * 8), v_desc, idesc_pv, ks ? 1u : first);
          }
          tc_commit(o_full);
          tc_commit(v_empty + 8 * st);
          if (!n1.valid) break;
          n0 = n1;
          n1 = n2;
        }
      }
    }
  } else {
    const int r = warp * 32 + lane;
    const uint32_t t_lane = tmem_base + ((uint32_t)(warp * 32) << 16);
    const uint32_t t_s = t_lane + AT_TM_S, t_o = t_lane + AT_TM_O, t_p = t_lane + AT_TM_P;
    const float c_log2 = 0.125f * 1.4426950408889634f;  // 1/sqrt(64) * log2(e)
    const uint64_t c2 = at_pack2(c_log2, c_log2);
    float m = -INFINITY, l = 0.f;
    // a finished head whose output rows are still in TMEM: read out one iteration late (or after the last item)
    bool owed = false;
    float owed_inv = 0.f;
    __nv_bfloat16 *owed_dst = nullptr;
    auto read_out = [&]() {
      uint32_t v[64];
      tc_ld_32x32(t_o, *reinterpret_cast<uint32_t(*)[32]>(&v[0]));
      tc_ld_32x32(t_o + 32, *reinterpret_cast<uint32_t(*)[32]>(&v[32]));
      tc_ld_wait();
      if (owed_dst != nullptr) {
#pragma unroll
        for (int j = 0; j < 64; j += 8) {
          uint32_t o4[4];
#pragma unroll
          for (int e = 0; e < 4; e++) {
            __nv_bfloat162 h = __floats2bfloat162_rn(__uint_as_float(v[j + 2 * e]) * owed_inv, __uint_as_float(v[j + 2 * e + 1]) * owed_inv);
            o4[e] = *reinterpret_cast<uint32_t *>(&h);
          }
          *reinterpret_cast<uint4 *>(owed_dst + j) = make_uint4(o4[0], o4[1], o4[2], o4[3]);
        }
      }
    };
    int it = 0;
    for (int n_item = 0;; n_item++) {
      mbar_wait(item_ready, n_item & 1);
      volatile AtItem *slot = &ring[n_item & 3];
      if (!slot->valid) break;  // (no arrival on item_taken: nothing is published after the end marker)
      const int row0 = slot->row0, n_pos = slot->n_pos, p0 = slot->p0, nb = slot->nb, head0 = slot->head0;
      mbar_arrive(item_taken);
      const int nq = STACKED ? nq_stacked : n_pos;  // valid tile rows
      const int pos_r = STACKED ? p0 : p0 + r;
      const int pos_w_lo = STACKED ? p0 : p0 + warp * 32;  // the warp's first position
      for (int hh = 0; hh < n_loop; hh++) {
        m = -INFINITY;  // new head: fresh online-softmax state
        l = 0.f;
        for (int kb = 0; kb < nb; kb++, it++) {
          const int lim = pos_r - kb * 128;
          // STACKED (decode): every valid row sits at the same position, so the 32-key chunks beyond it are masked for the whole
          // tile: they are neither loaded nor exponentiated (their P is an exact zero either way), and warps without a valid row
          // only keep the barrier protocol.  Teacher-forced tiles always walk all four chunks.
          const int n_ch = STACKED ? min(4, ((p0 - kb * 128) >> 5) + 1) : 4;
          const bool active = !STACKED || warp * 32 < nq;
          uint32_t w[128];  // S as raw f32 bits; the packed bf16 P overwrites w[0, 64) in place (pair (j, j+1) -> w[j/2], j/2 <= j)
          mbar_wait(s_full, it & 1);
          tc_fence_after();
          if (active) {
#pragma unroll
            for (int c = 0; c < 4; c++)
              if (!STACKED || c < n_ch) tc_ld_32x32(t_s + (uint32_t)(c * 32), *reinterpret_cast<uint32_t(*)[32]>(&w[c * 32]));
            tc_ld_wait();
          }
          tc_fence_before();
          mbar_arrive(s_free);  // S is in registers: the MMA warp may overwrite the S columns with the next S

          float alpha = 1.f;
          if (active) {
            // causal mask: keys after the row's position become -inf before the row max (exp2 -> exact 0 on the MUFU columns; the
            // polynomial columns of P are cleared after the exponentials).  Blocks entirely at or before the warp's first position take
            // the mask-free path (warp-uniform); both paths compute identical values for unmasked keys.
            const bool need_mask = kb * 128 + 127 > pos_w_lo;
            float r4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
            for (int c = 0; c < 4; c++) {
              if (!STACKED || c < n_ch) {
                if (need_mask) {
#pragma unroll
                  for (int j = c * 32; j < c * 32 + 32; j++) w[j] = (j <= lim) ? w[j] : 0xff800000u;
                }
#pragma unroll
                for (int j = c * 32; j < c * 32 + 32; j += 8) {
                  r4[0] = at_max3(r4[0], __uint_as_float(w[j]), __uint_as_float(w[j + 1]));
                  r4[1] = at_max3(r4[1], __uint_as_float(w[j + 2]), __uint_as_float(w[j + 3]));
                  r4[2] = at_max3(r4[2], __uint_as_float(w[j + 4]), __uint_as_float(w[j + 5]));
                  r4[3] = at_max3(r4[3], __uint_as_float(w[j + 6]), __uint_as_float(w[j + 7]));
                }
              }
            }
            const float raw = fmaxf(fmaxf(r4[0], r4[1]), fmaxf(r4[2], r4[3])) * c_log2;
            // lazy rescale: the running max only moves when the block's max exceeds it by more than 2^8 (p stays <= 256)
            if (raw > m + AT_RESCALE_LOG2) {
              alpha = at_ex2(m - raw);  // first block: ex2(-inf) = 0 (l = 0 and O is overwritten anyway)
              m = raw;
            }
            const uint64_t nm2 = at_pack2(-m, -m);
            uint64_t sa = at_pack2(0.f, 0.f), sb = sa;  // (sum of p[j], p[j+1]) over j = 0 mod 4 / j = 2 mod 4
#pragma unroll
            for (int c = 0; c < 4; c++) {
              if (!STACKED || c < n_ch) {
#pragma unroll
                for (int j = c * 32; j < c * 32 + 32; j += 4) {
                  float a0 = __uint_as_float(w[j]), a1 = __uint_as_float(w[j + 1]), b0 = __uint_as_float(w[j + 2]), b1 = __uint_as_float(w[j + 3]);
                  at_fma2(a0, a1, c2, nm2);
                  at_fma2(b0, b1, c2, nm2);
                  if ((AT_POLY_MASK >> ((j >> 1) & 7)) & 1u) {
                    at_ex2_poly2(a0, a1);
                  } else {
                    a0 = at_ex2(a0);
                    a1 = at_ex2(a1);
                  }
                  if ((AT_POLY_MASK >> (((j >> 1) + 1) & 7)) & 1u) {
                    at_ex2_poly2(b0, b1);
                  } else {
                    b0 = at_ex2(b0);
                    b1 = at_ex2(b1);
                  }
                  sa = at_add2(sa, at_pack2(a0, a1));
                  sb = at_add2(sb, at_pack2(b0, b1));
                  __nv_bfloat162 ha = __floats2bfloat162_rn(a0, a1), hb = __floats2bfloat162_rn(b0, b1);
                  w[j >> 1] = *reinterpret_cast<uint32_t *>(&ha);
                  w[(j >> 1) + 1] = *reinterpret_cast<uint32_t *>(&hb);
                }
              } else {
#pragma unroll
                for (int j = c * 16; j < c * 16 + 16; j++) w[j] = 0u;
              }
            }
            // The polynomial turns a masked key's -inf into 2^-125, not 0.  In the row sum that is far below one ulp (the sum holds
            // a term >= 2^-8), but P must carry exact zeros for masked keys (decode skips those chunks altogether): clear the
            // polynomial pairs of P beyond the row's position.  Only blocks that straddle the causal boundary get here.
            if (need_mask) {
#pragma unroll
              for (int i = 0; i < 64; i++) {
                if ((AT_POLY_MASK >> (i & 7)) & 1u) {
                  const uint32_t keep = (2 * i <= lim ? 0x0000ffffu : 0u) | (2 * i + 1 <= lim ? 0xffff0000u : 0u);
                  w[i] &= keep;
                }
              }
            }
            float s0, s1, s2, s3;
            at_unpack2(sa, s0, s1);
            at_unpack2(sb, s2, s3);
            l = fmaf(l, alpha, (s0 + s1) + (s2 + s3));
          }
          // The previous iteration's P V must have completed before its A operand (the P columns) is overwritten, before O is
          // rescaled, and before a finished head's O is read out (kb == 0: the head that ended in the previous iteration).
          if (it > 0) {
            mbar_wait(o_full, (it - 1) & 1);
            tc_fence_after();
          }
          if (owed) {  // warp-uniform
            if (active) read_out();
            owed = false;
          } else if (active && kb > 0 && __any_sync(0xffffffffu, alpha != 1.f)) {  // O *= alpha in tensor memory (rows that did not ask: * 1.0f, exact)
            uint32_t v[64];
            tc_ld_32x32(t_o, *reinterpret_cast<uint32_t(*)[32]>(&v[0]));
            tc_ld_32x32(t_o + 32, *reinterpret_cast<uint32_t(*)[32]>(&v[32]));
            tc_ld_wait();
#pragma unroll
            for (int j = 0; j < 64; j++) v[j] = __float_as_uint(__uint_as_float(v[j]) * alpha);
            tc_st_32x32(t_o, *reinterpret_cast<uint32_t(*)[32]>(&v[0]));
            tc_st_32x32(t_o + 32, *reinterpret_cast<uint32_t(*)[32]>(&v[32]));
          }
          if (active) {
            tc_st_32x32(t_p, *reinterpret_cast<uint32_t(*)[32]>(&w[0]));
            tc_st_32x32(t_p + 32, *reinterpret_cast<uint32_t(*)[32]>(&w[32]));
            tc_st_wait();
          }
          tc_fence_before();
          mbar_arrive(p_ready);
        }
        // head finished: its P V is still running; remember where its rows go
        owed = true;
        owed_inv = 1.0f / l;
        const int head = STACKED ? hh * G + r : head0 + hh;
        owed_dst = r < nq ? out + (size_t)(STACKED ? row0 : row0 + r) * (nh * 64) + head * 64 : nullptr;
      }
    }
    if (owed) {  // the last head of the last item
      mbar_wait(o_full, (it - 1) & 1);
      tc_fence_after();
      if (!STACKED || owed_dst != nullptr || warp == 0) read_out();
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
  // persistent grid: two CTAs per SM pull (tile, KV head) items from a counter that is zeroed in stream order before the launch
  // (the counter lives in the ctx's 64-byte device status block, after the error flag)
  int *work_counter = ctx->err_flag_dev + 8;
  const int n_sm = ctx->sm_count;
  CZ_CUDA_TRY(cudaMemsetAsync(work_counter, 0, sizeof(int), st));
  const int n_items = single_rows ? n_tiles : n_tiles * nkv;
  const unsigned grid = (unsigned)(n_items < 2 * n_sm ? n_items : 2 * n_sm);
  if (single_rows) {
    CZ_LAUNCH(ctx, CZ_K_ATTN,
              (czk::attn_tc_kernel<true><<<grid, czk::AT_THREADS, czk::AT_SMEM, st>>>(tq, tk, tv, pos, kv_base, tile_row0, tile_n, out, nh, nkv,
                                                                                     n_tiles, work_counter)));
  } else {
    CZ_LAUNCH(ctx, CZ_K_ATTN,
              (czk::attn_tc_kernel<false><<<grid, czk::AT_THREADS, czk::AT_SMEM, st>>>(tq, tk, tv, pos, kv_base, tile_row0, tile_n, out, nh, nkv,
                                                                                      n_tiles, work_counter)));
  }
  CZ_CHECK_LAUNCH();
  return CZ_OK;
}

}  // namespace cz
