// Measured-peak micro-benchmarks for the roofline denominators (not on the hot path):
// DMMA.8x8x4 issue rate and plain DFMA issue rate.  The DMMA loop is sensitive to how many independent
// accumulator chains a warp carries and to code placement, so several variants are provided and bench.py
// takes the best one as "the peak" (variant = chains per warp: 0 -> 4, 1 -> 8, 2 -> 16; all with
// 148 x 4 CTAs of 256 threads).
#include "sqb_common.cuh"

namespace {

constexpr int kPeakThreads = 256;
constexpr int kPeakCtas = 148 * 4;

template <int CHAINS>
__global__ void __launch_bounds__(kPeakThreads)
peak_dmma_kernel(int64_t iters, double* sink) {
  double c0[CHAINS], c1[CHAINS];
#pragma unroll
  for (int k = 0; k < CHAINS; ++k) {
    c0[k] = 0.0;
    c1[k] = 0.0;
  }
  const double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
  for (int64_t it = 0; it < iters; ++it) {
#pragma unroll
    for (int k = 0; k < CHAINS; ++k)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(c0[k]), "+d"(c1[k])
                   : "d"(a), "d"(b));
  }
  double s = 0.0;
#pragma unroll
  for (int k = 0; k < CHAINS; ++k) s += c0[k] + c1[k];
  if (s == 12345.678) sink[0] = s;  // keep the chains alive
}

constexpr int kFmaChains = 16;

__global__ void __launch_bounds__(kPeakThreads)
peak_dfma_kernel(int64_t iters, double* sink) {
  double c[kFmaChains];
#pragma unroll
  for (int k = 0; k < kFmaChains; ++k) c[k] = threadIdx.x * 1e-3 + k;
  const double a = 1.0000001, b = 1e-9;
  for (int64_t it = 0; it < iters; ++it) {
#pragma unroll
    for (int k = 0; k < kFmaChains; ++k) c[k] = fma(c[k], a, b);
  }
  double s = 0.0;
#pragma unroll
  for (int k = 0; k < kFmaChains; ++k) s += c[k];
  if (s == 12345.678) sink[0] = s;
}

}  // namespace

extern "C" int sqb_peak_dmma(int variant, int64_t iters, double* sink, double* flops_host, void* stream) {
  if (iters < 1 || !sink || variant < 0 || variant > 2) return SQB_E_BADARG;
  int chains;
  if (variant == 0) {
    chains = 4;
    peak_dmma_kernel<4><<<kPeakCtas, kPeakThreads, 0, sqb_stream(stream)>>>(iters, sink);
  } else if (variant == 1) {
    chains = 8;
    peak_dmma_kernel<8><<<kPeakCtas, kPeakThreads, 0, sqb_stream(stream)>>>(iters, sink);
  } else {
    chains = 16;
    peak_dmma_kernel<16><<<kPeakCtas, kPeakThreads, 0, sqb_stream(stream)>>>(iters, sink);
  }
  SQB_LAUNCH_CHECK();
  // one DMMA.8x8x4 = 8*8*4 FMA = 512 flop per warp instruction
  if (flops_host) *flops_host = (double)iters * chains * 512.0 * (kPeakThreads / 32) * kPeakCtas;
  return SQB_OK;
}

extern "C" int sqb_peak_dfma(int64_t iters, double* sink, double* flops_host, void* stream) {
  if (iters < 1 || !sink) return SQB_E_BADARG;
  peak_dfma_kernel<<<kPeakCtas, kPeakThreads, 0, sqb_stream(stream)>>>(iters, sink);
  SQB_LAUNCH_CHECK();
  if (flops_host) *flops_host = (double)iters * kFmaChains * 2.0 * kPeakThreads * kPeakCtas;
  return SQB_OK;
}
