// Dev micro-benchmark (not part of libslate_b200.so): dependent-issue latency of DFMA / DADD / DMMA.8x8x4 / SHFL on
// one warp, in SM clocks.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_latency fp64_latency.cu
#include <cstdio>
#include <cuda_runtime.h>

__global__ void dfma_chain(double* out, long long* cyc, int iters, double a, double b) {
  double x = threadIdx.x;
  long long t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < iters; ++i) x = fma(x, a, b);
  long long t1 = clock64();
  out[threadIdx.x] = x;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void dadd_chain(double* out, long long* cyc, int iters, double b) {
  double x = threadIdx.x;
  long long t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < iters; ++i) x = x + b;
  long long t1 = clock64();
  out[threadIdx.x] = x;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void shfl_chain(double* out, long long* cyc, int iters) {
  double x = threadIdx.x;
  long long t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < iters; ++i) x = __shfl_xor_sync(0xffffffffu, x, 1);
  long long t1 = clock64();
  out[threadIdx.x] = x;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void dmma_chain(double* out, long long* cyc, int iters, double a, double b) {
  double c0 = threadIdx.x, c1 = 1.0;
  long long t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < iters; ++i)
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
  long long t1 = clock64();
  out[threadIdx.x] = c0 + c1;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
// throughput: `chains` independent DFMA chains per thread, 4 warps per SMSP
template <int CH>
__global__ void dfma_tput(double* out, long long* cyc, int iters, double a, double b) {
  double x[CH];
  for (int c = 0; c < CH; ++c) x[c] = threadIdx.x + c;
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i)
#pragma unroll
    for (int c = 0; c < CH; ++c) x[c] = fma(x[c], a, b);
  long long t1 = clock64();
  double s = 0;
  for (int c = 0; c < CH; ++c) s += x[c];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

int main() {
  double* out; long long* cyc;
  cudaMalloc(&out, 1 << 20); cudaMallocManaged(&cyc, 8);
  const int iters = 1 << 16;
  dfma_chain<<<1, 32>>>(out, cyc, iters, 1.0000001, 1e-9); cudaDeviceSynchronize();
  printf("DFMA dependent latency  : %.2f clk\n", (double)cyc[0] / iters);
  dadd_chain<<<1, 32>>>(out, cyc, iters, 1e-9); cudaDeviceSynchronize();
  printf("DADD dependent latency  : %.2f clk\n", (double)cyc[0] / iters);
  shfl_chain<<<1, 32>>>(out, cyc, iters); cudaDeviceSynchronize();
  printf("SHFL.64 dependent latency: %.2f clk (two 32-bit shuffles)\n", (double)cyc[0] / iters);
  dmma_chain<<<1, 32>>>(out, cyc, iters, 1.0000001, 1e-9); cudaDeviceSynchronize();
  printf("DMMA.8x8x4 dependent latency: %.2f clk\n", (double)cyc[0] / iters);
  dfma_tput<1><<<1, 512>>>(out, cyc, iters, 1.0000001, 1e-9); cudaDeviceSynchronize();
  printf("DFMA 16 warps x 1 chain : %.2f clk per DFMA per warp (SM issue interval %.2f)\n", (double)cyc[0] / iters, (double)cyc[0] / iters / 16);
  dfma_tput<8><<<1, 512>>>(out, cyc, iters / 8, 1.0000001, 1e-9); cudaDeviceSynchronize();
  printf("DFMA 16 warps x 8 chains: %.2f clk per DFMA per warp -> %.1f DFMA lanes / clk / SM\n", (double)cyc[0] / iters,
         16.0 * 32 * iters / (double)cyc[0]);
  dfma_tput<8><<<1, 128>>>(out, cyc, iters / 8, 1.0000001, 1e-9); cudaDeviceSynchronize();
  printf("DFMA 4 warps x 8 chains : %.1f DFMA lanes / clk / SM\n", 4.0 * 32 * iters / (double)cyc[0]);
  return 0;
}
