// K7: batched complex128 GEMM on the FP64 tensor pipe + the row-wise reductions of the operator / state algebra
// (SURVEY section 8f, ranks 1 and 4).
//
// Replaces the einsum contractions behind slate_quantum.operator.{matmul, commute, matmul_list_operator,
// matmul_operator_list, get_commutator_operator_list} (operator/_linalg.py:108-215), operator.{apply, expectation,
// expectation_of_each} (operator/_operator.py:165-215) and state.{inner_product_each, normalize_each,
// get_all_occupations} (state/_state.py:208-308).  The reference routes all of them through numpy einsum on
// dense [M, M] arrays; here one kernel covers every operand layout through (row stride, column stride, conj)
// triples, so transposes, daggers and the "each operator of a list" batching never materialise a copy.
//
//   C[b] = alpha * op(A[b]) * op(B[b]) + beta * C[b],   A(i, k) = a[b*stride_a + i*a_rs + k*a_cs] (conjugated if asked)
//
// Tensor roofline: 8 m n k real flop per batch entry.  64x64 output tile per CTA, 4 warps of 32x32, K chunks of 16
// through a 3-stage cp.async ring of interleaved (re, im) tiles (16-byte fragment loads, conflict free with a row
// stride == 2 mod 8 elements), 4 real DMMA.8x8x4 products per complex MAC.
#include "sqb_common.cuh"

namespace {

constexpr int kZgTile = 64;
constexpr int kZgKc = 16;
constexpr int kZgStages = 3;
constexpr int kZgLd = kZgTile + 2;  // row stride == 2 (mod 8) complex elements: conflict-free 16-byte fragment loads
constexpr int kZgThreads = 128;
#ifndef SQB_ZG_GROUP
#define SQB_ZG_GROUP 16
#endif
constexpr int kZgGroup = SQB_ZG_GROUP;  // tile rows per L2 group
constexpr int kZgStageElems = 2 * kZgKc * kZgLd;  // A chunk [kk][m] + B chunk [kk][n], interleaved (re, im)
constexpr size_t kZgSmem = (size_t)kZgStages * kZgStageElems * sizeof(c128);

__device__ __forceinline__ void zg_dmma(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(d0), "+d"(d1)
               : "d"(a), "d"(b));
}
// 16-byte asynchronous global -> shared copy (LDGSTS); src_bytes == 0 writes zeros (edge tiles)
__device__ __forceinline__ void zg_cp_async(c128* dst, const c128* src, int src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src),
               "r"(src_bytes)
               : "memory");
}
// sign flip on the integer pipe (the FP64 pipe belongs to the DMMAs)
__device__ __forceinline__ double zg_flip(double v, int mask) {
  return __hiloint2double(__double2hiint(v) ^ mask, __double2loint(v));
}

struct ZgOperand {
  const c128* p;
  int64_t rs, cs, bs;  // row / column / batch strides in elements
  int conj;
};

// tile element q of thread `tid`: (outer index in the 64-long dimension, kk in the K chunk); `kfast` picks the
// mapping whose consecutive threads walk the unit-stride direction of the operand
__device__ __forceinline__ void zg_map(int idx, bool kfast, int& o, int& kk) {
  if (kfast) {
    o = idx >> 4;
    kk = idx & (kZgKc - 1);
  } else {
    kk = idx >> 6;
    o = idx & 63;
  }
}

// K chunks of 16 travel global -> shared as 16-byte cp.async copies through a 3-stage ring (one block barrier per
// chunk; nothing is staged in registers), operands stay interleaved (re, im) in shared memory so that a fragment is
// ONE 16-byte load, and conjugation is a sign flip of the imaginary fragment on the integer pipe.
__global__ void __launch_bounds__(kZgThreads, 2)
zgemm_kernel(int m, int n, int k, ZgOperand A, ZgOperand B, c128* __restrict__ c, int64_t ldc, int64_t c_bs,
             c128 alpha, c128 beta, int64_t batch0) {
  extern __shared__ __align__(16) c128 zg_smem[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t b = batch0 + blockIdx.z;
  // grouped tile order: consecutive CTAs walk kZgGroup tile rows column by column, so the CTAs resident at the same
  // time share a few A panels and a few B panels in L2 instead of streaming all of B once per tile row
  const int tiles_m = (m + kZgTile - 1) / kZgTile, tiles_n = (n + kZgTile - 1) / kZgTile;
  const int width = kZgGroup * tiles_n;
  const int first_m = ((int)blockIdx.x / width) * kZgGroup;
  const int gsize = min(tiles_m - first_m, kZgGroup);
  const int i0 = (first_m + ((int)blockIdx.x % width) % gsize) * kZgTile;
  const int j0 = (((int)blockIdx.x % width) / gsize) * kZgTile;
  const c128* a = A.p + b * A.bs;
  const c128* bb = B.p + b * B.bs;
  const bool a_kfast = A.cs == 1, b_kfast = B.rs == 1 && B.cs != 1;
  const int a_mask = A.conj ? (int)0x80000000 : 0, b_mask = B.conj ? (int)0x80000000 : 0;

  const int wm = (warp & 1) * 32, wn = (warp >> 1) * 32;
  const int frow = lane >> 2, fcol = lane & 3;
  double cr[4][4][2], ci[4][4][2];
#pragma unroll
  for (int x = 0; x < 4; ++x)
#pragma unroll
    for (int y = 0; y < 4; ++y) cr[x][y][0] = cr[x][y][1] = ci[x][y][0] = ci[x][y][1] = 0.0;

  const int nchunks = (k + kZgKc - 1) / kZgKc;
  auto issue = [&](int ch) {  // every thread; always closes a commit group (empty past the last chunk)
    if (ch < nchunks) {
      const int kc = ch * kZgKc;
      c128* As = zg_smem + (size_t)(ch % kZgStages) * kZgStageElems;
      c128* Bs = As + kZgKc * kZgLd;
#pragma unroll
      for (int q = 0; q < (kZgTile * kZgKc) / kZgThreads; ++q) {
        const int idx = tid + kZgThreads * q;
        int o, kk;
        zg_map(idx, a_kfast, o, kk);
        bool ok = i0 + o < m && kc + kk < k;
        zg_cp_async(As + kk * kZgLd + o, ok ? a + (int64_t)(i0 + o) * A.rs + (int64_t)(kc + kk) * A.cs : a, ok ? 16 : 0);
        zg_map(idx, b_kfast, o, kk);
        ok = j0 + o < n && kc + kk < k;
        zg_cp_async(Bs + kk * kZgLd + o, ok ? bb + (int64_t)(kc + kk) * B.rs + (int64_t)(j0 + o) * B.cs : bb, ok ? 16 : 0);
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  issue(0);
  issue(1);
  for (int ch = 0; ch < nchunks; ++ch) {
    asm volatile("cp.async.wait_group 1;" ::: "memory");  // chunk ch has landed (only the newest group may be pending)
    __syncthreads();                                      // ... for every thread, and stage (ch - 1) % 3 is free
    issue(ch + 2);
    const c128* As = zg_smem + (size_t)(ch % kZgStages) * kZgStageElems;
    const c128* Bs = As + kZgKc * kZgLd;
#pragma unroll
    for (int k4 = 0; k4 < kZgKc; k4 += 4) {
      double ar[4], ai[4], nai[4], br[4], bi[4];
#pragma unroll
      for (int x = 0; x < 4; ++x) {
        const c128 v = As[(k4 + fcol) * kZgLd + wm + x * 8 + frow];
        ar[x] = v.x;
        ai[x] = zg_flip(v.y, a_mask);
        nai[x] = zg_flip(v.y, a_mask ^ (int)0x80000000);
      }
#pragma unroll
      for (int y = 0; y < 4; ++y) {
        const c128 v = Bs[(k4 + fcol) * kZgLd + wn + y * 8 + frow];
        br[y] = v.x;
        bi[y] = zg_flip(v.y, b_mask);
      }
#pragma unroll
      for (int x = 0; x < 4; ++x)
#pragma unroll
        for (int y = 0; y < 4; ++y) {
          zg_dmma(cr[x][y][0], cr[x][y][1], ar[x], br[y]);
          zg_dmma(ci[x][y][0], ci[x][y][1], ar[x], bi[y]);
          zg_dmma(cr[x][y][0], cr[x][y][1], nai[x], bi[y]);
          zg_dmma(ci[x][y][0], ci[x][y][1], ai[x], br[y]);
        }
    }
  }

  c128* cb = c + b * c_bs;
  const bool use_beta = beta.x != 0.0 || beta.y != 0.0;
#pragma unroll
  for (int x = 0; x < 4; ++x) {
    const int i = i0 + wm + x * 8 + frow;
    if (i >= m) continue;
    c128* crow = cb + (int64_t)i * ldc;
#pragma unroll
    for (int y = 0; y < 4; ++y) {
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int j = j0 + wn + y * 8 + fcol * 2 + h;
        if (j >= n) continue;
        c128 v = cmul(alpha, make_double2(cr[x][y][h], ci[x][y][h]));
        if (use_beta) cfma(v, beta, crow[j]);
        crow[j] = v;
      }
    }
  }
}

constexpr int kRowThreads = 256;

// out[r] = sum_i conj(x[r, i]) * y[r, i]
__global__ void __launch_bounds__(kRowThreads)
rowwise_vdot_kernel(int64_t len, const c128* __restrict__ x, int64_t x_rs, const c128* __restrict__ y, int64_t y_rs,
                    c128* __restrict__ out) {
  __shared__ double red[2][kRowThreads / 32];
  const int64_t r = blockIdx.x;
  const c128* xr = x + r * x_rs;
  const c128* yr = y + r * y_rs;
  c128 acc = make_double2(0.0, 0.0);
  for (int64_t i = threadIdx.x; i < len; i += kRowThreads) cfmac(acc, xr[i], yr[i]);
  acc.x = warp_sum(acc.x);
  acc.y = warp_sum(acc.y);
  if ((threadIdx.x & 31) == 0) {
    red[0][threadIdx.x >> 5] = acc.x;
    red[1][threadIdx.x >> 5] = acc.y;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double sx = 0.0, sy = 0.0;
    for (int w = 0; w < kRowThreads / 32; ++w) {
      sx += red[0][w];
      sy += red[1][w];
    }
    out[r] = make_double2(sx, sy);
  }
}

// out[r] = sum_i w[i] |x[r, i]|^2: the expectation of an operator that is diagonal in the basis x is written in
__global__ void __launch_bounds__(kRowThreads)
rowwise_weighted_abs2_kernel(int64_t len, const c128* __restrict__ x, int64_t x_rs, const c128* __restrict__ w,
                             c128* __restrict__ out) {
  __shared__ double red[2][kRowThreads / 32];
  const int64_t r = blockIdx.x;
  const c128* xr = x + r * x_rs;
  double ax = 0.0, ay = 0.0;
  for (int64_t i = threadIdx.x; i < len; i += kRowThreads) {
    const c128 z = xr[i], wi = w[i];
    const double p = fma(z.x, z.x, z.y * z.y);
    ax = fma(wi.x, p, ax);
    ay = fma(wi.y, p, ay);
  }
  ax = warp_sum(ax);
  ay = warp_sum(ay);
  if ((threadIdx.x & 31) == 0) {
    red[0][threadIdx.x >> 5] = ax;
    red[1][threadIdx.x >> 5] = ay;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double sx = 0.0, sy = 0.0;
    for (int k = 0; k < kRowThreads / 32; ++k) {
      sx += red[0][k];
      sy += red[1][k];
    }
    out[r] = make_double2(sx, sy);
  }
}

// out[r, i] = w[i] * x[r, i]: a diagonal operator applied to every state of a list
__global__ void __launch_bounds__(kRowThreads)
scale_columns_kernel(int64_t rows, int64_t len, const c128* __restrict__ x, int64_t x_rs, const c128* __restrict__ w,
                     c128* __restrict__ out) {
  const int64_t stride = (int64_t)gridDim.x * kRowThreads;
  for (int64_t r = blockIdx.y; r < rows; r += gridDim.y)
    for (int64_t i = (int64_t)blockIdx.x * kRowThreads + threadIdx.x; i < len; i += stride)
      out[r * len + i] = cmul(w[i], x[r * x_rs + i]);
}

// out[i] = |x[i]|^2
__global__ void __launch_bounds__(kRowThreads)
abs2_kernel(int64_t count, const c128* __restrict__ x, double* __restrict__ out) {
  const int64_t stride = (int64_t)gridDim.x * kRowThreads;
  for (int64_t i = (int64_t)blockIdx.x * kRowThreads + threadIdx.x; i < count; i += stride) {
    const c128 z = x[i];
    out[i] = fma(z.x, z.x, z.y * z.y);
  }
}

// out[r, i] = x[r, i] / sqrt(|norm2[r]|)  (normalize_each divides by sqrt of <psi|psi>, state/_state.py:242-262)
__global__ void __launch_bounds__(kRowThreads)
normalize_rows_kernel(int64_t rows, int64_t len, const c128* __restrict__ x, const c128* __restrict__ norm2, c128* __restrict__ out) {
  const int64_t stride = (int64_t)gridDim.x * kRowThreads;
  for (int64_t r = blockIdx.y; r < rows; r += gridDim.y) {
    const double s = rsqrt(hypot(norm2[r].x, norm2[r].y));
    for (int64_t i = (int64_t)blockIdx.x * kRowThreads + threadIdx.x; i < len; i += stride)
      out[r * len + i] = cscale(x[r * len + i], s);
  }
}

}  // namespace

extern "C" int sqb_zgemm_batched(int m, int n, int k, int64_t batch, const sqb_c128* a, int64_t a_row_stride,
                                 int64_t a_col_stride, int64_t a_batch_stride, int conj_a, const sqb_c128* b,
                                 int64_t b_row_stride, int64_t b_col_stride, int64_t b_batch_stride, int conj_b,
                                 const double* alpha, const double* beta, sqb_c128* c, int64_t ldc,
                                 int64_t c_batch_stride, void* stream) {
  if (m < 0 || n < 0 || k < 0 || batch < 0 || !alpha || !beta) return SQB_E_BADARG;
  if (m == 0 || n == 0 || batch == 0) return SQB_OK;
  if (!c || ldc < n || (k > 0 && (!a || !b))) return SQB_E_BADARG;
  ZgOperand A{reinterpret_cast<const c128*>(a), a_row_stride, a_col_stride, a_batch_stride, conj_a != 0};
  ZgOperand B{reinterpret_cast<const c128*>(b), b_row_stride, b_col_stride, b_batch_stride, conj_b != 0};
  const c128 al = make_double2(alpha[0], alpha[1]), be = make_double2(beta[0], beta[1]);
  const int64_t tiles = (int64_t)((n + kZgTile - 1) / kZgTile) * ((m + kZgTile - 1) / kZgTile);
  if (tiles > 0x7fffffff) return SQB_E_UNSUPPORTED;
  cudaError_t attr = cudaFuncSetAttribute(zgemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kZgSmem);
  if (attr != cudaSuccess) return (int)attr;
  for (int64_t b0 = 0; b0 < batch; b0 += 65535) {
    const unsigned gz = (unsigned)sqb_min64(65535, batch - b0);
    zgemm_kernel<<<dim3((unsigned)tiles, 1, gz), kZgThreads, kZgSmem, sqb_stream(stream)>>>(
        m, n, k, A, B, reinterpret_cast<c128*>(c), ldc, c_batch_stride, al, be, b0);
    SQB_LAUNCH_CHECK();
  }
  return SQB_OK;
}

extern "C" int sqb_rowwise_vdot(int64_t rows, int64_t len, const sqb_c128* x, int64_t x_row_stride, const sqb_c128* y,
                                int64_t y_row_stride, sqb_c128* out, void* stream) {
  if (rows < 0 || len < 0) return SQB_E_BADARG;
  if (rows == 0) return SQB_OK;
  if (!out || (len > 0 && (!x || !y))) return SQB_E_BADARG;
  for (int64_t r0 = 0; r0 < rows; r0 += 0x7fffffff) {
    const unsigned g = (unsigned)sqb_min64(0x7fffffff, rows - r0);
    rowwise_vdot_kernel<<<g, kRowThreads, 0, sqb_stream(stream)>>>(
        len, reinterpret_cast<const c128*>(x) + r0 * x_row_stride, x_row_stride,
        reinterpret_cast<const c128*>(y) + r0 * y_row_stride, y_row_stride, reinterpret_cast<c128*>(out) + r0);
    SQB_LAUNCH_CHECK();
  }
  return SQB_OK;
}

extern "C" int sqb_rowwise_weighted_abs2(int64_t rows, int64_t len, const sqb_c128* x, int64_t x_row_stride,
                                        const sqb_c128* weights, sqb_c128* out, void* stream) {
  if (rows < 0 || len < 0) return SQB_E_BADARG;
  if (rows == 0) return SQB_OK;
  if (!out || (len > 0 && (!x || !weights))) return SQB_E_BADARG;
  for (int64_t r0 = 0; r0 < rows; r0 += 0x7fffffff) {
    const unsigned g = (unsigned)sqb_min64(0x7fffffff, rows - r0);
    rowwise_weighted_abs2_kernel<<<g, kRowThreads, 0, sqb_stream(stream)>>>(
        len, reinterpret_cast<const c128*>(x) + r0 * x_row_stride, x_row_stride,
        reinterpret_cast<const c128*>(weights), reinterpret_cast<c128*>(out) + r0);
    SQB_LAUNCH_CHECK();
  }
  return SQB_OK;
}

extern "C" int sqb_scale_columns(int64_t rows, int64_t len, const sqb_c128* x, int64_t x_row_stride,
                                const sqb_c128* weights, sqb_c128* out, void* stream) {
  if (rows < 0 || len < 0) return SQB_E_BADARG;
  if (rows == 0 || len == 0) return SQB_OK;
  if (!x || !weights || !out) return SQB_E_BADARG;
  const unsigned gx = (unsigned)sqb_min64((len + kRowThreads - 1) / kRowThreads, 148 * 8);
  scale_columns_kernel<<<dim3(gx, (unsigned)sqb_min64(rows, 65535)), kRowThreads, 0, sqb_stream(stream)>>>(
      rows, len, reinterpret_cast<const c128*>(x), x_row_stride, reinterpret_cast<const c128*>(weights),
      reinterpret_cast<c128*>(out));
  SQB_LAUNCH_CHECK();
  return SQB_OK;
}

extern "C" int sqb_abs2(int64_t count, const sqb_c128* x, double* out, void* stream) {
  if (count < 0) return SQB_E_BADARG;
  if (count == 0) return SQB_OK;
  if (!x || !out) return SQB_E_BADARG;
  const unsigned g = (unsigned)sqb_min64((count + kRowThreads - 1) / kRowThreads, 148 * 8);
  abs2_kernel<<<g, kRowThreads, 0, sqb_stream(stream)>>>(count, reinterpret_cast<const c128*>(x), out);
  SQB_LAUNCH_CHECK();
  return SQB_OK;
}

extern "C" int sqb_normalize_rows(int64_t rows, int64_t len, const sqb_c128* x, const sqb_c128* norm2, sqb_c128* out,
                                  void* stream) {
  if (rows < 0 || len < 0) return SQB_E_BADARG;
  if (rows == 0 || len == 0) return SQB_OK;
  if (!x || !norm2 || !out) return SQB_E_BADARG;
  const unsigned gx = (unsigned)sqb_min64((len + kRowThreads - 1) / kRowThreads, 148 * 8);
  normalize_rows_kernel<<<dim3(gx, (unsigned)sqb_min64(rows, 65535)), kRowThreads, 0, sqb_stream(stream)>>>(
      rows, len, reinterpret_cast<const c128*>(x), reinterpret_cast<const c128*>(norm2), reinterpret_cast<c128*>(out));
  SQB_LAUNCH_CHECK();
  return SQB_OK;
}
