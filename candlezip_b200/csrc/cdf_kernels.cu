// K1 / K9: logits -> quantised integer CDF bounds, CDF search, cross-entropy bits.
//
// Replaces src/main.rs:784-824 (softmax_pdf + quantize_pdf_to_cdf), 758-782 (softmax_pdf_floor,
// combined_pdf_with_literals), 2294-2326 (encode-side use), 2622-2625 (decode-side search), 1743-1749 /
// 1776-1784 (XE accumulation).  The reference's arithmetic is ORDER-DEPENDENT f64 (sequential sum over the
// vocab, sequential prefix), so bit-exact parity forbids tree reductions: one thread owns one column (one
// stream / one teacher-forced position) and walks the vocab in ascending order.  Parallelism comes from the
// columns (all positions of all chunks are independent on the encode side).  Logits are VOCAB-MAJOR
// [V][ld] so that the 32 columns of a warp read one coalesced 128-byte line per vocab entry.
//
// expf: identical operation sequence to glibc 2.39 x86_64 expf (the FMA ifunc variant), which is what Rust's
// f32::exp lowers to on Linux; every mul/add/fma is an explicit round-to-nearest intrinsic so nvcc cannot
// re-contract.  HBM-bound by design: algorithmic bytes = 4*V per column (one read); this implementation
// reads the column up to 3 times (max / sum / prefix), the 2nd and 3rd mostly from L2 for decode-sized M.
#include "cdf_device.cuh"

namespace czk {

// One thread per column of the vocab-major logits.
template <int MODE, int OP>
__global__ void __launch_bounds__(128) cdf_cols_kernel(const float *__restrict__ logits, int V, size_t M, size_t ld,
                                                       const uint32_t *__restrict__ syms_or_values,
                                                       uint32_t *__restrict__ sym_out, uint32_t *__restrict__ c_lo_out,
                                                       uint32_t *__restrict__ c_hi_out, double *__restrict__ xe_out,
                                                       int *__restrict__ err, const int *__restrict__ colmax) {
  __shared__ uint32_t s_lo[32 * 32], s_hi[32 * 32];
  exp_tab_init(s_lo, s_hi);
  ExpTab tab{s_lo, s_hi, (int)(threadIdx.x & 31)};
  size_t col = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool active = col < M;
  if (!active) col = M - 1;  // keep the warp converged; results of inactive lanes are discarded
  uint32_t sym, lo, hi;
  double xe;
  int errbits = 0;
  cdf_col<MODE, OP>(logits + col, ld, V, syms_or_values[col], active, tab, sym, lo, hi, xe, errbits, colmax != nullptr,
                    colmax ? colmax_decode(colmax[col]) : 0.f);
  if (!active) return;
  if (errbits) atomicOr(err, errbits);
  if (OP == OP_XE) {
    xe_out[col] = xe;
  } else {
    if (OP == OP_SEARCH) sym_out[col] = sym;
    c_lo_out[col] = lo;
    c_hi_out[col] = hi;
  }
}

// OP_SEARCH for a small number of columns (stepwise decode): one warp per column, see cdf_search_warp
template <int MODE>
__global__ void __launch_bounds__(256) cdf_search_warp_kernel(const float *__restrict__ logits, int V, size_t M, size_t ld,
                                                              const uint32_t *__restrict__ values, uint32_t *__restrict__ sym_out,
                                                              uint32_t *__restrict__ c_lo_out, uint32_t *__restrict__ c_hi_out,
                                                              int *__restrict__ err, const int *__restrict__ colmax) {
  __shared__ uint32_t s_lo[32 * 32], s_hi[32 * 32];
  __shared__ __align__(16) double s_xch[8 * 64];  // per warp: two 32-value exchange lines (cdf_search_warp)
  exp_tab_init(s_lo, s_hi);
  ExpTab tab{s_lo, s_hi, (int)(threadIdx.x & 31)};
  const size_t col = (size_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (col >= M) return;
  float mx;
  if (colmax) {
    mx = colmax_decode(colmax[col]);
  } else {
    mx = __int_as_float(0xff800000);
    for (int v = threadIdx.x & 31; v < V; v += 32) {
      const float x = logits[(size_t)v * ld + col];
      if (x > mx) mx = x;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  }
  uint32_t sym, lo, hi;
  int errbits = 0;
  cdf_search_warp<MODE>(logits + col, ld, V, values[col], mx, tab, sym, lo, hi, errbits, s_xch + (threadIdx.x >> 5) * 64);
  if ((threadIdx.x & 31) == 0) {
    if (errbits) atomicOr(err, errbits);
    sym_out[col] = sym;
    c_lo_out[col] = lo;
    c_hi_out[col] = hi;
  }
}

// full CDF of one column (debug / watchdog parity): single thread, sequential
template <int MODE>
__global__ void cdf_full_kernel(const float *__restrict__ logits, int V, uint32_t *__restrict__ cdf) {
  __shared__ uint32_t s_lo[32 * 32], s_hi[32 * 32];
  exp_tab_init(s_lo, s_hi);
  if (threadIdx.x != 0) return;
  ExpTab tab{s_lo, s_hi, 0};
  const int n_sym = MODE == CZ_CDF_RWKV_LITERALS ? V + 256 : V;
  float mx = __int_as_float(0xff800000);
  for (int v = 0; v < V; v++)
    if (logits[v] > mx) mx = logits[v];
  double S = 0.0;
  for (int v = 0; v < V; v++) S = __dadd_rn(S, (double)cz_expf(__fsub_rn(logits[v], mx), tab));
  double norm = 1.0, sum2 = 1.0;
  const double scale = 1.0 - 256.0 * CZ_P_FLOOR;
  if (MODE == CZ_CDF_RWKV_LITERALS) {
    double acc = 0.0;
    for (int v = 0; v < V; v++)
      acc = __dadd_rn(acc, fmax(__ddiv_rn((double)cz_expf(__fsub_rn(logits[v], mx), tab), S), CZ_P_FLOOR));
    norm = acc;
    acc = 0.0;
    for (int v = 0; v < V; v++) {
      double q = fmax(__ddiv_rn((double)cz_expf(__fsub_rn(logits[v], mx), tab), S), CZ_P_FLOOR);
      acc = __dadd_rn(acc, __dmul_rn(__ddiv_rn(q, norm), scale));
    }
    for (int j = 0; j < 256; j++) acc = __dadd_rn(acc, CZ_P_FLOOR);
    sum2 = acc;
  }
  const bool uniform = MODE == CZ_CDF_SMOLLM && S <= 0.0;
  double acc = 0.0;
  uint32_t prev = 0;
  cdf[0] = 0;
  for (int v = 0; v < n_sym; v++) {
    double q;
    if (MODE == CZ_CDF_RWKV_LITERALS) {
      if (v < V) {
        q = fmax(__ddiv_rn((double)cz_expf(__fsub_rn(logits[v], mx), tab), S), CZ_P_FLOOR);
        q = __dmul_rn(__ddiv_rn(q, norm), scale);
      } else {
        q = CZ_P_FLOOR;
      }
      if (sum2 > 0.0) q = __ddiv_rn(q, sum2);
    } else {
      q = uniform ? 1.0 / (double)V : __ddiv_rn((double)cz_expf(__fsub_rn(logits[v], mx), tab), S);
    }
    acc = __dadd_rn(acc, q);
    uint32_t cur = quant(acc);
    if (cur < prev) cur = prev;
    cdf[v + 1] = cur;
    prev = cur;
  }
  cdf[n_sym] = CZ_AC_CDF_TOTAL;
}

__global__ void fill_i32_kernel(int *__restrict__ p, int v, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

}  // namespace czk

namespace cz {

int launch_fill_i32(cz_ctx *ctx, int *p, int v, size_t n, cudaStream_t stream) {
  if (n == 0) return CZ_OK;
  CZ_LAUNCH(ctx, CZ_K_OTHER, (czk::fill_i32_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, stream>>>(p, v, n)));
  CZ_CHECK_LAUNCH();
  return CZ_OK;
}

// Device-pointer launchers (used by the executor and by the host-buffer C-ABI wrappers in api.cu)
int launch_cdf_cols(cz_ctx *ctx, int op, int mode, const float *logits_dev, size_t V, size_t M, size_t ld,
                    const uint32_t *arg_dev, uint32_t *sym_out_dev, uint32_t *c_lo_dev, uint32_t *c_hi_dev,
                    double *xe_dev, cudaStream_t stream, const int *colmax_dev) {
  if (M == 0) return CZ_OK;
  if (V == 0 || V > (1u << 24) || ld < M) {
    set_error("cdf: bad shape");
    return CZ_ERR_INVALID;
  }
  if (op == czk::OP_SEARCH && M <= 8192) {  // decode-sized: warp per column
    const unsigned g = (unsigned)ceil_div(M, 8);
    if (mode == CZ_CDF_SMOLLM)
      CZ_LAUNCH(ctx, CZ_K_CDF,
                (czk::cdf_search_warp_kernel<CZ_CDF_SMOLLM><<<g, 256, 0, stream>>>(logits_dev, (int)V, M, ld, arg_dev, sym_out_dev, c_lo_dev,
                                                                                   c_hi_dev, ctx->err_flag_dev, colmax_dev)));
    else if (mode == CZ_CDF_RWKV_LITERALS)
      CZ_LAUNCH(ctx, CZ_K_CDF,
                (czk::cdf_search_warp_kernel<CZ_CDF_RWKV_LITERALS><<<g, 256, 0, stream>>>(logits_dev, (int)V, M, ld, arg_dev, sym_out_dev,
                                                                                          c_lo_dev, c_hi_dev, ctx->err_flag_dev, colmax_dev)));
    else {
      set_error("cdf: unknown mode");
      return CZ_ERR_INVALID;
    }
    CZ_CHECK_LAUNCH();
    return CZ_OK;
  }
  const int threads = 128;
  dim3 grid((unsigned)ceil_div(M, threads));
#define CZ_CDF_CASE(MODE, OP)                                                                                 \
  CZ_LAUNCH(ctx, CZ_K_CDF,                                                                                    \
            (czk::cdf_cols_kernel<MODE, OP><<<grid, threads, 0, stream>>>(logits_dev, (int)V, M, ld, arg_dev, \
                                                                          sym_out_dev, c_lo_dev, c_hi_dev,    \
                                                                          xe_dev, ctx->err_flag_dev, colmax_dev)))
  if (mode == CZ_CDF_SMOLLM) {
    if (op == czk::OP_BOUNDS) CZ_CDF_CASE(CZ_CDF_SMOLLM, czk::OP_BOUNDS);
    else if (op == czk::OP_SEARCH) CZ_CDF_CASE(CZ_CDF_SMOLLM, czk::OP_SEARCH);
    else CZ_CDF_CASE(CZ_CDF_SMOLLM, czk::OP_XE);
  } else if (mode == CZ_CDF_RWKV_LITERALS) {
    if (op == czk::OP_BOUNDS) CZ_CDF_CASE(CZ_CDF_RWKV_LITERALS, czk::OP_BOUNDS);
    else if (op == czk::OP_SEARCH) CZ_CDF_CASE(CZ_CDF_RWKV_LITERALS, czk::OP_SEARCH);
    else CZ_CDF_CASE(CZ_CDF_RWKV_LITERALS, czk::OP_XE);
  } else {
    set_error("cdf: unknown mode");
    return CZ_ERR_INVALID;
  }
#undef CZ_CDF_CASE
  CZ_CHECK_LAUNCH();
  return CZ_OK;
}

int launch_cdf_full(cz_ctx *ctx, int mode, const float *logits_dev, size_t V, uint32_t *cdf_dev, cudaStream_t stream) {
  if (mode == CZ_CDF_SMOLLM)
    CZ_LAUNCH(ctx, CZ_K_CDF, (czk::cdf_full_kernel<CZ_CDF_SMOLLM><<<1, 32, 0, stream>>>(logits_dev, (int)V, cdf_dev)));
  else
    CZ_LAUNCH(ctx, CZ_K_CDF,
              (czk::cdf_full_kernel<CZ_CDF_RWKV_LITERALS><<<1, 32, 0, stream>>>(logits_dev, (int)V, cdf_dev)));
  CZ_CHECK_LAUNCH();
  return CZ_OK;
}

}  // namespace cz
