// K6: observables of position-basis state lists.
//
// Replaces the diagonal expectations behind slate_quantum.operator.measure.{all_x, all_periodic_x,
// all_variance_x, all_coherent_width} (operator/_measure.py:191-305): every one of them is
//   sum_x f(x_axis) |psi_t(x)|^2 / sum_x |psi_t(x)|^2
// with f = x (optionally shifted by a per-state offset and wrapped into [-L/2, L/2)), f = x^2, or
// f = exp(+2 pi i n_axis / N_axis) (the fundamental "scattering" operator whose angle is the periodic <x>).
// One kernel returns the five sums per state; the reference builds a dense diagonal operator and an einsum
// per call instead.  HBM bound: 16 B read per element, one pass (the [T, M] array is read ONCE for
// x, x^2 and the scattering phase together).
//
// Orthogonal axes only: x along `axis` = n_axis * spacing (for a non-orthogonal cell the Cartesian component
// mixes lattice axes; SQB_E_UNSUPPORTED is the caller's job to raise, the kernel takes the spacing it is given).
#include <math.h>

#include "sqb_common.cuh"

namespace {

constexpr int kMeasureThreads = 512;
constexpr int kMeasureMaxTable = 4096;

// states [lead, M], M = outer * len * inner with `len` the extent of the measured axis
__global__ void __launch_bounds__(kMeasureThreads)
measure_moments_kernel(int64_t m_total, int len, int64_t inner, double spacing, double length,
                       const double* __restrict__ offsets, int wrapped, const c128* __restrict__ states,
                       double* __restrict__ out) {
  extern __shared__ double measure_smem[];
  double* xs = measure_smem;  // [len] (wrapped) position of index i along the axis
  double* cs = xs + len;      // [len]
  double* sn = cs + len;      // [len]
  __shared__ double red[5][kMeasureThreads / 32];
  const int64_t t = blockIdx.x;
  const double off = offsets ? offsets[t] : 0.0;
  const bool table = len <= kMeasureMaxTable;  // longer axes (cfg2: 131072 points) evaluate x / sincospi per element
  for (int i = threadIdx.x; table && i < len; i += kMeasureThreads) {
    double x = (double)i * spacing + off;
    if (wrapped) x -= length * floor(x / length + 0.5);
    xs[i] = x;
    double s, c;
    sincospi(2.0 * (double)i / (double)len, &s, &c);
    cs[i] = c;
    sn[i] = s;
  }
  __syncthreads();
  const c128* row = states + t * m_total;
  double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0, a4 = 0.0;
  const bool small = m_total < (1LL << 31);
  for (int64_t m = threadIdx.x; m < m_total; m += kMeasureThreads) {
    const c128 z = row[m];
    const double w = fma(z.x, z.x, z.y * z.y);
    int i;
    if (small) i = (int)(((uint32_t)m / (uint32_t)inner) % (uint32_t)len);
    else i = (int)((m / inner) % len);
    double x, c, sv;
    if (table) {
      x = xs[i];
      c = cs[i];
      sv = sn[i];
    } else {
      x = (double)i * spacing + off;
      if (wrapped) x -= length * floor(x / length + 0.5);
      sincospi(2.0 * (double)i / (double)len, &sv, &c);
    }
    a0 += w;
    a1 = fma(w, x, a1);
    a2 = fma(w * x, x, a2);
    a3 = fma(w, c, a3);
    a4 = fma(w, sv, a4);
  }
  double v[5] = {a0, a1, a2, a3, a4};
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < 5; ++k) {
    v[k] = warp_sum(v[k]);
    if (lane == 0) red[k][warp] = v[k];
  }
  __syncthreads();
  if (warp == 0) {
#pragma unroll
    for (int k = 0; k < 5; ++k) {
      double r = lane < kMeasureThreads / 32 ? red[k][lane] : 0.0;
      r = warp_sum(r);
      if (lane == 0) out[t * 5 + k] = r;
    }
  }
}

}  // namespace

extern "C" int sqb_measure_moments(int nd, const int* grid_host, int64_t lead, const sqb_c128* states, int axis,
                                   double spacing, double length, const double* offsets, int wrapped, double* out,
                                   void* stream) {
  if (nd < 1 || nd > SQB_MAX_DIMS || !grid_host || !states || !out || axis < 0 || axis >= nd || lead < 0 ||
      !(length > 0.0))
    return SQB_E_BADARG;
  int64_t m_total = 1, inner = 1;
  for (int a = 0; a < nd; ++a) {
    if (grid_host[a] < 1) return SQB_E_BADARG;
    m_total *= grid_host[a];
    if (a > axis) inner *= grid_host[a];
  }
  if (lead == 0) return SQB_OK;
  if (lead >= (1LL << 31)) return SQB_E_UNSUPPORTED;
  const size_t smem = grid_host[axis] <= kMeasureMaxTable ? (size_t)3 * grid_host[axis] * sizeof(double) : 0;
  cudaError_t e = cudaFuncSetAttribute(measure_moments_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  measure_moments_kernel<<<(unsigned)lead, kMeasureThreads, smem, sqb_stream(stream)>>>(
      m_total, grid_host[axis], inner, spacing, length, offsets, wrapped, reinterpret_cast<const c128*>(states), out);
  SQB_LAUNCH_CHECK();
  return SQB_OK;
}
