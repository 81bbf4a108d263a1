// Measured-peak micro-benchmarks for the roofline denominators (not on the hot path):
// DMMA.8x8x4 issue rate and plain DFMA issue rate with enough independent chains per warp to hide
// latency, 148 SMs x 8 warps x 4 CTAs.  bench.py times them with CUDA events.
#include "sqb_common.cuh"

namespace {

constexpr int kPeakThreads = 256;
constexpr int kPeakCtas = 148 * 4;
constexpr int kChains = 16;

__global__ void __launch_bounds__(kPeakThreads)
peak_dmma_kernel(int64_t iters, double* sink) {
  double c0[kChains], c1[kChains];
#pragma unroll
  for (int k = 0; k < kChains; ++k) {
    c0[k] = 0.0;
    c1[k] = 0.0;
  }
  const double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
  for (int64_t it = 0; it < iters; ++it) {
#pragma unroll
    for (int k = 0; k < kChains; ++k)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(c0[k]), "+d"(c1[k])
                   : "d"(a), "d"(b));
  }
  double s = 0.0;
#pragma unroll
  for (int k = 0; k < kChains; ++k) s += c0[k] + c1[k];
  if (s == 12345.678) sink[0] = s;  // keep the chains alive
}

__global__ void __launch_bounds__(kPeakThreads)
peak_dfma_kernel(int64_t iters, double* sink) {
  double c[kChains];
#pragma unroll
  for (int k = 0; k < kChains; ++k) c[k] = threadIdx.x * 1e-3 + k;
  const double a = 1.0000001, b = 1e-9;
  for (int64_t it = 0; it < iters; ++it) {
#pragma unroll
    for (int k = 0; k < kChains; ++k) c[k] = fma(c[k], a, b);
  }
  double s = 0.0;
#pragma unroll
  for (int k = 0; k < kChains; ++k) s += c[k];
  if (s == 12345.678) sink[0] = s;
}

}  // namespace

extern "C" int sqb_peak_dmma(int64_t iters, double* sink, double* flops_host, void* stream) {
  if (iters < 1 || !sink) return SQB_E_BADARG;
  peak_dmma_kernel<<<kPeakCtas, kPeakThreads, 0, sqb_stream(stream)>>>(iters, sink);
  SQB_LAUNCH_CHECK();
  // one DMMA.8x8x4 = 8*8*4 FMA = 512 flop per warp instruction
  if (flops_host) *flops_host = (double)iters * kChains * 512.0 * (kPeakThreads / 32) * kPeakCtas;
  return SQB_OK;
}

extern "C" int sqb_peak_dfma(int64_t iters, double* sink, double* flops_host, void* stream) {
  if (iters < 1 || !sink) return SQB_E_BADARG;
  peak_dfma_kernel<<<kPeakCtas, kPeakThreads, 0, sqb_stream(stream)>>>(iters, sink);
  SQB_LAUNCH_CHECK();
  if (flops_host) *flops_host = (double)iters * kChains * 2.0 * kPeakThreads * kPeakCtas;
  return SQB_OK;
}
