// Shared helpers for the slate_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/slate_b200.h"

#define SQB_LAUNCH_CHECK()                         \
  do {                                             \
    cudaError_t e__ = cudaGetLastError();          \
    if (e__ != cudaSuccess) return (int)e__;       \
  } while (0)

typedef double2 c128;  // .x = re, .y = im

__device__ __forceinline__ c128 cmul(c128 a, c128 b) {
  return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
// a * conj(b)
__device__ __forceinline__ c128 cmulc(c128 a, c128 b) {
  return make_double2(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y);
}
// acc += a*b
__device__ __forceinline__ void cfma(c128& acc, c128 a, c128 b) {
  acc.x = fma(a.x, b.x, acc.x);
  acc.x = fma(-a.y, b.y, acc.x);
  acc.y = fma(a.x, b.y, acc.y);
  acc.y = fma(a.y, b.x, acc.y);
}
// acc += conj(a)*b
__device__ __forceinline__ void cfmac(c128& acc, c128 a, c128 b) {
  acc.x = fma(a.x, b.x, acc.x);
  acc.x = fma(a.y, b.y, acc.x);
  acc.y = fma(a.x, b.y, acc.y);
  acc.y = fma(-a.y, b.x, acc.y);
}
__device__ __forceinline__ c128 cadd(c128 a, c128 b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ c128 csub(c128 a, c128 b) { return make_double2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ c128 cscale(c128 a, double s) { return make_double2(a.x * s, a.y * s); }
__device__ __forceinline__ c128 cconj(c128 a) { return make_double2(a.x, -a.y); }

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

static inline cudaStream_t sqb_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

struct SqbDims {
  int nd;
  int shape[SQB_MAX_DIMS];
  int repeat[SQB_MAX_DIMS];
};

static inline int sqb_fill_dims(SqbDims& d, int nd, const int* shape, const int* repeat, int64_t* n_out,
                                int64_t* nk_out) {
  if (nd < 1 || nd > SQB_MAX_DIMS || !shape || !repeat) return SQB_E_BADARG;
  d.nd = nd;
  int64_t n = 1, nk = 1;
  for (int a = 0; a < SQB_MAX_DIMS; ++a) {
    d.shape[a] = a < nd ? shape[a] : 1;
    d.repeat[a] = a < nd ? repeat[a] : 1;
    if (d.shape[a] < 1 || d.repeat[a] < 1) return SQB_E_BADARG;
    n *= d.shape[a];
    nk *= d.repeat[a];
  }
  *n_out = n;
  *nk_out = nk;
  return SQB_OK;
}

__host__ __device__ __forceinline__ int64_t sqb_min64(int64_t a, int64_t b) { return a < b ? a : b; }
