// Dev micro-benchmark: DMMA.8x8x4 throughput versus warps per SM and independent chains per warp.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma_occupancy dmma_occupancy.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int CHAINS>
__global__ void k(long iters, double* sink) {
  double c0[CHAINS], c1[CHAINS];
#pragma unroll
  for (int i = 0; i < CHAINS; ++i) { c0[i] = 0; c1[i] = 0; }
  double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
  for (long it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < CHAINS; ++i)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(c0[i]), "+d"(c1[i]) : "d"(a), "d"(b));
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < CHAINS; ++i) s += c0[i] + c1[i];
  if (s == 1234.5) sink[0] = s;
}

template <int CHAINS>
void run(int warps_per_sm, int ctas_per_sm) {
  double* sink; cudaMalloc(&sink, 64);
  int threads = warps_per_sm * 32 / ctas_per_sm;
  long iters = 200000 / CHAINS;
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  k<CHAINS><<<148 * ctas_per_sm, threads>>>(iters, sink);
  cudaDeviceSynchronize();
  cudaEventRecord(a);
  k<CHAINS><<<148 * ctas_per_sm, threads>>>(iters, sink);
  cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  double flops = (double)iters * CHAINS * 512.0 * warps_per_sm * 148;
  printf("warps/SM %2d ctas/SM %d chains %2d : %.2f TFLOP/s\n", warps_per_sm, ctas_per_sm, CHAINS, flops / ms / 1e9);
  cudaFree(sink);
}

int main() {
  for (int w : {4, 8, 16, 32}) { run<4>(w, 1); run<16>(w, 1); run<64>(w, 1); }
  run<16>(32, 4);
  return 0;
}
