// Micro-benchmark: FP64 dependent-issue latency and throughput on B200 (the CDF kernels' sequential f64 sums are chains of DADDs).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/ubench/fp64_lat scripts/ubench/fp64_lat.cu && scripts/ubench/fp64_lat
#include <cstdio>
#include <cuda_runtime.h>

template <int CHAINS, bool FMA>
__global__ void chain_kernel(double *out, const double *in, int iters, long long *cycles) {
  double a[CHAINS];
  const double x = in[0], y = in[1];
#pragma unroll
  for (int c = 0; c < CHAINS; c++) a[c] = in[2 + c];
  const long long t0 = clock64();
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int c = 0; c < CHAINS; c++) a[c] = FMA ? __fma_rn(a[c], x, y) : __dadd_rn(a[c], y);
  }
  const long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int c = 0; c < CHAINS; c++) s += a[c];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

template <int CHAINS, bool FMA>
static void run(const char *name, int warps_per_sm) {
  double *out, *in;
  long long *cyc, h;
  cudaMalloc(&out, 8 * 1024 * 1024);
  cudaMalloc(&in, 8 * 64);
  cudaMalloc(&cyc, 8);
  double hin[64];
  for (int i = 0; i < 64; i++) hin[i] = 1.0 + 1e-9 * i;
  cudaMemcpy(in, hin, sizeof(hin), cudaMemcpyHostToDevice);
  const int iters = 20000;
  chain_kernel<CHAINS, FMA><<<148, 32 * warps_per_sm>>>(out, in, iters, cyc);
  cudaDeviceSynchronize();
  cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("%-6s chains/thread=%d warps/SM=%2d: %.2f cycles per dependent step, %.2f cycles per warp-instruction per SM sub-partition\n", name, CHAINS,
         warps_per_sm, (double)h / iters, (double)h / iters / CHAINS / (warps_per_sm / 4.0 < 1 ? 1 : warps_per_sm / 4.0));
  cudaFree(out);
  cudaFree(in);
  cudaFree(cyc);
}

int main() {
  run<1, false>("DADD", 1);
  run<1, true>("DFMA", 1);
  run<2, false>("DADD", 1);
  run<4, false>("DADD", 1);
  run<8, false>("DADD", 1);
  run<16, false>("DADD", 1);
  run<8, true>("DFMA", 1);
  run<8, false>("DADD", 4);
  run<8, false>("DADD", 16);
  run<8, true>("DFMA", 16);
  run<1, false>("DADD", 32);
  return 0;
}
