// Round-2 CDF arithmetic: the same values as cdf_device.cuh, bit for bit, in far fewer issue slots.
//
// What bounded the round-1 kernel (profiles/ncu_summary_r01g.md: 33 instructions per logit, 58% issue, 0.32 of HBM) was not the
// ten FP64 operations of the glibc expf but what surrounds them:
//   * three F2F conversions per logit (f32 -> f64 of the argument, f64 -> f32 -> f64 of the result).  Conversions to / from 64-bit
//     types run at a quarter of the FP64 rate, so they cost more pipe time than the arithmetic.  Here the argument is widened with
//     one integer multiply-add ((u64)bits * 2^29 + bias: exact for a normal f32) and the result is rounded to the f32 grid
//     (round-to-nearest-even at mantissa bit 29) by integer adds on the f64 bit pattern -- valid while the result is a normal f32,
//     i.e. for max - logit <= 87; anything else (tiny results, -inf logits, NaN) takes the original conversion path, decided once
//     per group of rows.  cz_test_expf_exhaustive checks fast == original for every one of the 2^31 + NaN argument patterns.
//   * an IEEE division per element in every pass after the sum (p = e / S; the RWKV alphabet divides three times).  The divisor is
//     a per-column constant, so its correctly rounded reciprocal y = RN(1 / b) is computed once and each quotient costs five
//     FMA-pipe operations:  q0 = RN(a y);  r0 = RN(a - b q0);  q1 = RN(q0 + r0 y);  r1 = a - b q1 (exact);  q = RN(q1 + r1 y).
//     q1 is a faithful quotient (error < 1 ulp), and Markstein's theorem (Handbook of Floating-Point Arithmetic, thm 4.10: y within
//     1/2 ulp of 1/b, q1 faithful  =>  RN(q1 + r1 y) = RN(a / b)) makes q the correctly rounded quotient, i.e. __ddiv_rn(a, b).
//     cz_test_div_random compares the two on 2^36 random operand pairs of the shapes that occur here.
//   * the encode-side prefix walk stops at the coded symbol, but a warp walks until its LAST lane stops: with real token ids (mean
//     id / V = 0.1, heavy tail) some lane of 32 nearly always needs most of the vocabulary.  The prefix kernel therefore sorts the
//     256 columns of a CTA by symbol and hands each warp 32 columns of similar walk length (any column -> any lane: the columns
//     are independent, results unchanged).
#pragma once
#include "cdf_device.cuh"

namespace czk {

// 64-bit [entry][lane] copy of the exp2 table: one conflict-free LDS.64 per lookup (a half-warp's 16 lanes hit 16 distinct
// 8-byte bank pairs)
__device__ __forceinline__ void exp_tab64_init(uint64_t *s_tab) {
  for (int i = threadIdx.x; i < 32 * 32; i += blockDim.x) s_tab[i] = c_exp2f_tab[i >> 5];
  __syncthreads();
}

constexpr float CZ_EXP_FAST_MAX = 87.0f;  // e^-87 = 1.6e-38 > 2^-126: the result is a normal f32

// (double)expf(-a) for 0 <= a <= CZ_EXP_FAST_MAX (a = max - logit), identical to (double)cz_expf(-a): the operation sequence of
// cz_expf on the negated constant (RN is sign-symmetric), with the conversions done in integer arithmetic.
// tab_lane = table base + lane.
__device__ __forceinline__ double cz_exp_neg_fast(float a, const uint64_t *__restrict__ tab_lane) {
  const double ninv = -(0x1.71547652b82fep+0 * 32);
  const double shift = 0x1.8p+52;
  const double c0 = 0x1.c6af84b912394p-5 / 32 / 32 / 32;
  const double c1 = 0x1.ebfce50fac4f3p-3 / 32 / 32;
  const double c2 = 0x1.62e42ff0c52d6p-1 / 32;
  // f32 -> f64 of a normal non-negative float: exponent rebias 0x380 << 52, mantissa << 29.  For a = 0 or a subnormal this yields
  // a value below 2^-126 instead: harmless, both give expf = 1.0f exactly (|x| < 2^-25 rounds to 1).
  const uint64_t w = (uint64_t)__float_as_uint(a) * 0x20000000ull + 0x3800000000000000ull;
  const double ad = __longlong_as_double((long long)w);
  const double z = __dmul_rn(ninv, ad);
  double kd = __dadd_rn(z, shift);
  const uint32_t ki = (uint32_t)__double2loint(kd);
  kd = __dsub_rn(kd, shift);
  const double r = __fma_rn(ninv, ad, -kd);
  // (explicit shared-window address: lane slot + entry * 256 bytes is one LOP3 + one IMAD; through the generic pointer the compiler
  // rebuilds the lane part of the index for every lookup)
  uint64_t t;
  asm("ld.shared.u64 %0, [%1];" : "=l"(t) : "r"((uint32_t)__cvta_generic_to_shared(tab_lane) + ((ki & 31u) << 8)));
  const double s = __hiloint2double((int)((uint32_t)(t >> 32) + (ki << 15)), (int)(uint32_t)t);  // t += ki << 47
  const double p = __fma_rn(c0, r, c1);
  const double r2 = __dmul_rn(r, r);
  double y = __fma_rn(c2, r, 1.0);
  y = __fma_rn(p, r2, y);
  y = __dmul_rn(y, s);
  // RN-even to 24 significant bits: add half an f32 ulp minus one, plus the parity of the bit that stays; clear the 29 dropped bits
  uint64_t yb = (uint64_t)__double_as_longlong(y);
  yb += 0x0FFFFFFFull + ((yb >> 29) & 1ull);
  yb &= ~0x1FFFFFFFull;
  return __longlong_as_double((long long)yb);
}

// the f32 bit pattern of a double that lies on the normal-f32 grid (what cz_exp_neg_fast returns)
__device__ __forceinline__ float cz_f32_of_gridded(double d) {
  const uint32_t hi = (uint32_t)__double2hiint(d), lo = (uint32_t)__double2loint(d);
  return __uint_as_float(__funnelshift_l(lo, hi - 0x38000000u, 3));
}

// exact f32 -> f64 for values that only matter when they are >= 2^-126 (see above): one IMAD.WIDE instead of an F2F
__device__ __forceinline__ double cz_widen_pos(float e) {
  return __longlong_as_double((long long)((uint64_t)__float_as_uint(e) * 0x20000000ull + 0x3800000000000000ull));
}

// (cz_div_rcp / cz_div_rcp_ok: cdf_device.cuh -- the decode-side search uses them too)

// ---- vectorised full-column walk: NCOL adjacent columns per thread ------------------------------------------------------------
template <int NCOL>
struct CdfVec;
template <>
struct CdfVec<1> {
  using T = float;
  static __device__ __forceinline__ void unpack(const T &t, float (&x)[1]) { x[0] = t; }
};
template <>
struct CdfVec<2> {
  using T = float2;
  static __device__ __forceinline__ void unpack(const T &t, float (&x)[2]) {
    x[0] = t.x;
    x[1] = t.y;
  }
};
template <>
struct CdfVec<4> {
  using T = float4;
  static __device__ __forceinline__ void unpack(const T &t, float (&x)[4]) {
    x[0] = t.x;
    x[1] = t.y;
    x[2] = t.z;
    x[3] = t.w;
  }
};

// Calls f(v0, rows, cnt) for consecutive groups of up to GRP rows (cnt valid) of the NCOL columns starting at p, in ascending
// row order, with the next group's loads in flight and an L2 prefetch CDF_PF rows ahead (cdf_walk's pipelining).  NC: read-only path.
template <int NCOL, int GRP, bool NC, class F>
__device__ __forceinline__ void cdf_walk_groups(const float *p, size_t ld, int n, F f) {
  using T = typename CdfVec<NCOL>::T;
  auto ldv = [&](size_t row) -> T {
    const T *q = reinterpret_cast<const T *>(p + row * ld);
    return NC ? __ldg(q) : *q;
  };
  T cur[GRP], nxt[GRP];
  const T zero = T();
#pragma unroll
  for (int k = 0; k < GRP; k++) cur[k] = k < n ? ldv((size_t)k) : zero;
  int v0 = 0;
  for (; v0 + CDF_PF + GRP <= n; v0 += GRP) {
#pragma unroll
    for (int k = 0; k < GRP; k++) {
      nxt[k] = ldv((size_t)(v0 + GRP + k));
      asm volatile("prefetch.global.L2 [%0];" ::"l"(p + (size_t)(v0 + CDF_PF + k) * ld));
    }
    f(v0, cur, GRP);
#pragma unroll
    for (int k = 0; k < GRP; k++) cur[k] = nxt[k];
  }
  for (; v0 < n; v0 += GRP) {
#pragma unroll
    for (int k = 0; k < GRP; k++) {
      const int v = v0 + GRP + k;
      nxt[k] = v < n ? ldv((size_t)v) : zero;
    }
    f(v0, cur, n - v0 < GRP ? n - v0 : GRP);
#pragma unroll
    for (int k = 0; k < GRP; k++) cur[k] = nxt[k];
  }
}

// per-column results of the full passes, consumed by the prefix / XE kernels
struct __align__(16) CdfStats {
  double S;      // sum_i (f64)expf(l_i - max)                                        (src/main.rs:789-793)
  double norm;   // sum_i max(e_i / S, 2^-29)          (softmax_pdf_floor, :763-764)  -- RWKV alphabet and XE only, else 1
  double sum2;   // combined_pdf_with_literals' second normaliser (:773-779)           -- RWKV alphabet only, else 1
  float mx;      // max_i l_i
  int pad;
};

}  // namespace czk
