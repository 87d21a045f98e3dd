// FP64 latency / throughput micro-benchmark for B200 (sm_100a): dependent DFMA chains with ILP 1..8 per warp,
// 1..16 warps per SM sub-partition.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_lat fp64_lat.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int ILP> __global__ void k(double *out, int iters, double a, double b)
{
  double x[ILP];
#pragma unroll
  for (int i = 0; i < ILP; i++) x[i] = threadIdx.x * 1e-3 + i;
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < ILP; i++) x[i] = fma(x[i], a, b);
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; i++) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s + (double) (t1 - t0) * 1e-300;
  if (threadIdx.x == 0 && blockIdx.x == 0) ((long long *) out)[1 << 20] = t1 - t0;
}
template <int ILP> void run(double *d, int threads)
{
  const int iters = 4096;
  k<ILP><<<148, threads>>>(d, iters, 1.0000001, 1e-9);
  cudaDeviceSynchronize();
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k<ILP><<<148, threads>>>(d, iters, 1.0000001, 1e-9);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  long long cyc; cudaMemcpy(&cyc, ((long long *) d) + (1 << 20), 8, cudaMemcpyDeviceToHost);
  const double per_dep = (double) cyc / iters;                 // cycles per round of ILP independent DFMAs in one warp
  const double warp_instr = (double) iters * ILP * (threads / 32);
  printf("threads/SM %4d (warps/SMSP %2d) ILP %d: %.1f cycles per dependent step, %.2f warp-DFMA/clk/SM, %.1f TFLOP/s\n", threads, threads / 128,
         ILP, per_dep, warp_instr / cyc, 2.0 * iters * ILP * threads * 148 / (ms * 1e-3) / 1e12);
}
int main()
{
  double *d; cudaMalloc(&d, (1 << 20) * 8 + 64);
  for (int threads : {128, 256, 640, 1024}) { run<1>(d, threads); run<2>(d, threads); run<4>(d, threads); run<8>(d, threads); }
  return 0;
}
