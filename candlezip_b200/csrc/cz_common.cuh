// Shared helpers for the candlezip_b200 CUDA sources (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <string>
#include <vector>

#include "../../include/candlezip_b200.h"

namespace cz {

void set_error(const std::string &msg);
const char *get_error();

#define CZ_CUDA_TRY(expr)                                                                              \
  do {                                                                                                 \
    cudaError_t _e = (expr);                                                                           \
    if (_e != cudaSuccess) {                                                                           \
      cz::set_error(std::string(#expr) + ": " + cudaGetErrorString(_e) + " @" + __FILE__ + ":" +       \
                    std::to_string(__LINE__));                                                         \
      return CZ_ERR_CUDA;                                                                              \
    }                                                                                                  \
  } while (0)

#define CZ_TRY(expr)              \
  do {                            \
    int _rc = (expr);             \
    if (_rc != CZ_OK) return _rc; \
  } while (0)

static inline size_t ceil_div(size_t a, size_t b) { return (a + b - 1) / b; }

}  // namespace cz

// One ctx == one GPU == one stream of work (plus a side stream for overlap).
struct cz_ctx {
  int device = -1;          // -1: host-only ctx (weight generation / container work without a GPU)
  int sm_count = 0;
  cudaStream_t stream = nullptr;
  cudaStream_t stream2 = nullptr;
  cudaStream_t prof_stream = nullptr;  // stream the next launches go to when it is not `stream` (profiling events are recorded there)
  uint64_t launches = 0;
  // per-family profiling: mode 1 = synchronous (event pair + sync per launch), mode 2 = deferred (event pairs are
  // pooled and only read in cz_profile_read, so the timed region is not perturbed by host syncs)
  int prof_mode = 0;
  bool capturing = false;  // a CUDA graph capture is in progress on `stream`: no profiling events, no host syncs
  std::vector<cudaEvent_t> ev_pool;       // [2*i], [2*i+1] = start/stop of pooled launch i
  std::vector<int> ev_fam;
  size_t ev_used = 0;
  bool prof_on = false;
  double prof_ms[CZ_K_FAMILIES] = {0};
  uint64_t prof_launches[CZ_K_FAMILIES] = {0};
  cudaEvent_t prof_ev0 = nullptr, prof_ev1 = nullptr;
  // scratch for the small kernels' host entry points
  void *scratch = nullptr;
  size_t scratch_bytes = 0;
  void *cdf_stats = nullptr;  // per-column results of the CDF kernels' full passes (cdf_kernels.cu), grow-only
  size_t cdf_stats_bytes = 0;
  // compact per-column copy of e_v for v <= coded symbol, written by the stats pass and read by the prefix walk (cdf_kernels.cu)
  void *cdf_ecache = nullptr;
  size_t cdf_ecache_bytes = 0;
  void *cdf_eoff = nullptr;  // u32 offsets [M + 1] (units of 8 doubles) + the "cache in use" flag
  size_t cdf_eoff_bytes = 0;
  const void *cdf_eflag_last = nullptr;  // device address of the last batch's flag; the total follows the offsets right before it
  bool cdf_ecache_failed = false;  // the allocation did not fit once: the prefix walk reads the logits from then on
  int *err_flag_dev = nullptr;  // 64-byte device status block: [0] flag set by kernels on zero-width intervals etc., [8] the
                                // attention kernel's work-item counter (attn_tc.cu)
};

namespace cz {

int ensure_scratch(cz_ctx *ctx, size_t bytes);

// RAII-free launch accounting: CZ_LAUNCH(ctx, family, kernel<<<...>>>(...))
struct LaunchScope {
  cz_ctx *ctx;
  int fam;
  size_t slot = 0;
  cudaStream_t st;
  LaunchScope(cz_ctx *c, int f) : ctx(c), fam(f), st(c->prof_stream ? c->prof_stream : c->stream) {
    if (ctx->capturing) return;
    if (ctx->prof_mode == 2) {
      slot = ctx->ev_used++;
      if (ctx->ev_pool.size() < 2 * ctx->ev_used) {
        cudaEvent_t a, b;
        cudaEventCreate(&a);
        cudaEventCreate(&b);
        ctx->ev_pool.push_back(a);
        ctx->ev_pool.push_back(b);
        ctx->ev_fam.push_back(f);
      }
      ctx->ev_fam[slot] = f;
      cudaEventRecord(ctx->ev_pool[2 * slot], st);
    } else if (ctx->prof_on) {
      cudaEventRecord(ctx->prof_ev0, st);
    }
  }
  ~LaunchScope() {
    ctx->launches++;
    ctx->prof_launches[fam]++;
    if (ctx->capturing) return;
    if (ctx->prof_mode == 2) {
      cudaEventRecord(ctx->ev_pool[2 * slot + 1], st);
    } else if (ctx->prof_on) {
      cudaEventRecord(ctx->prof_ev1, st);
      cudaEventSynchronize(ctx->prof_ev1);
      float ms = 0.f;
      cudaEventElapsedTime(&ms, ctx->prof_ev0, ctx->prof_ev1);
      ctx->prof_ms[fam] += ms;
    }
  }
};

#define CZ_LAUNCH(ctx, fam, ...)       \
  do {                                 \
    cz::LaunchScope _ls((ctx), (fam)); \
    __VA_ARGS__;                       \
  } while (0)

#define CZ_CHECK_LAUNCH()                                                                          \
  do {                                                                                             \
    cudaError_t _e = cudaGetLastError();                                                           \
    if (_e != cudaSuccess) {                                                                       \
      cz::set_error(std::string("kernel launch: ") + cudaGetErrorString(_e) + " @" + __FILE__ +   \
                    ":" + std::to_string(__LINE__));                                               \
      return CZ_ERR_CUDA;                                                                          \
    }                                                                                              \
  } while (0)

}  // namespace cz
