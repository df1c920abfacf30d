// Micro-benchmark: MUFU ex2 throughput per SM, f32 vs packed f16x2 (two exponentials per operation), for the attention softmax.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/ubench/mufu_ex2 scripts/ubench/mufu_ex2.cu && scripts/ubench/mufu_ex2
#include <cstdint>
#include <cstdio>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

template <int MODE>  // 0: ex2.approx.ftz.f32   1: ex2.approx.f16x2 (packed halves: no .ftz form)   2: ex2.approx.ftz.bf16x2
__global__ void k(float *out, int iters, long long *cycles) {
  uint32_t a[8];
#pragma unroll
  for (int c = 0; c < 8; c++) a[c] = MODE == 0 ? __float_as_uint(-0.001f * (threadIdx.x + c)) : 0xB800B900u + c;  // small negative halves
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int c = 0; c < 8; c++) {
      if (MODE == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+r"(a[c]));
      else if (MODE == 1) asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(a[c]));
      else asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(a[c]));
    }
  }
  const long long t1 = clock64();
  uint32_t s = 0;
#pragma unroll
  for (int c = 0; c < 8; c++) s ^= a[c];
  out[blockIdx.x * blockDim.x + threadIdx.x] = __uint_as_float(s);
  if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

template <int MODE>
static void run(const char *name) {
  float *out;
  long long *cyc, h;
  cudaMalloc(&out, 4 * 148 * 1024);
  cudaMalloc(&cyc, 8);
  const int iters = 4000, warps = 16;
  k<MODE><<<148, 32 * warps>>>(out, iters, cyc);
  cudaDeviceSynchronize();
  cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  const double ops = (double)iters * 8 * warps;  // warp-instructions per SM
  printf("%-24s %.2f SM-cycles per warp instruction (%0.1f exponentials per clock per SM)\n", name, h / ops, 32.0 * (MODE ? 2 : 1) * ops / h);
  cudaFree(out);
  cudaFree(cyc);
}

int main() {
  run<0>("ex2.approx.ftz.f32");
  run<1>("ex2.approx.f16x2");
  run<2>("ex2.approx.ftz.bf16x2");
  return 0;
}
