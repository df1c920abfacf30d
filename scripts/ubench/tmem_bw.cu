// Micro-benchmark: tcgen05.ld / tcgen05.st throughput per warp, per SM (how many bytes per clock move between TMEM and registers),
// with 1..4 active warps per CTA and 1 or 2 CTAs per SM.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 tmem_bw.cu -o tmem_bw
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../candlezip_b200/csrc/tc_ptx.cuh"
using namespace czk;

__device__ __forceinline__ void st32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]),
      "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]),
      "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}

// mode 0: loads (4 x32 loads then one wait, like the attention softmax); mode 1: stores; mode 2: loads with a wait after each
__global__ void __launch_bounds__(128) k(int iters, int active_warps, int mode, long long *cyc, uint32_t *sink) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "n"(256) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t base = slot + ((uint32_t)(warp * 32) << 16);
  uint32_t acc = 0;
  uint32_t v[4][32];
#pragma unroll
  for (int c = 0; c < 4; c++)
#pragma unroll
    for (int j = 0; j < 32; j++) v[c][j] = threadIdx.x + j;
  __syncthreads();
  long long t0 = clock64();
  if (warp < active_warps) {
    for (int i = 0; i < iters; i++) {
      if (mode == 0) {
#pragma unroll
        for (int c = 0; c < 4; c++) tc_ld_32x32(base + c * 32, v[c]);
        tc_ld_wait();
#pragma unroll
        for (int c = 0; c < 4; c++) acc += v[c][0] ^ v[c][31];
      } else if (mode == 1) {
#pragma unroll
        for (int c = 0; c < 4; c++) st32(base + c * 32, v[c]);
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      } else {
#pragma unroll
        for (int c = 0; c < 4; c++) {
          tc_ld_32x32(base + c * 32, v[c]);
          tc_ld_wait();
          acc += v[c][0] ^ v[c][31];
        }
      }
    }
  }
  long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
  if (acc == 0x12345678) sink[0] = acc;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(slot), "n"(256) : "memory");
  }
}

int main() {
  long long *cyc;
  uint32_t *sink;
  cudaMalloc(&cyc, 1024 * 8);
  cudaMalloc(&sink, 4);
  const int iters = 2000;
  long long h[1024];
  for (int mode = 0; mode < 3; mode++)
    for (int ctas_per_sm = 1; ctas_per_sm <= 2; ctas_per_sm++)
      for (int aw = 1; aw <= 4; aw *= 2) {
        int grid = 148 * ctas_per_sm;
        k<<<grid, 128>>>(iters, aw, mode, cyc, sink);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
        cudaMemcpy(h, cyc, grid * 8, cudaMemcpyDeviceToHost);
        double avg = 0;
        for (int i = 0; i < grid; i++) avg += h[i];
        avg /= grid;
        const double bytes_warp = (double)iters * 4 * 4096;  // per warp
        printf("mode %d (%s) ctas/SM %d active warps/CTA %d: %.0f cycles, %.1f B/clk per warp, %.1f B/clk per SM\n", mode,
               mode == 0 ? "ld x4 + wait" : mode == 1 ? "st x4 + wait" : "ld + wait each", ctas_per_sm, aw, avg, bytes_warp / avg,
               bytes_warp * aw * ctas_per_sm / avg);
      }
  return 0;
}
