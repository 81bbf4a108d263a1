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

// General (non-orthogonal) cells: the Cartesian component `axis` of the position mixes every lattice axis,
//   x = sum_a frac_a A[a][axis],  frac_a = n_a / N_a + off_t * finv[a]   (folded into [-1/2, 1/2) when wrapped),
// finv = column `axis` of (A^T)^-1, i.e. the fractional coordinates of a unit Cartesian offset along `axis`
// (slate_core fundamental_stacked_x_points semantics as restated by oracle.stacked_x_points).  The scattering
// phase only depends on the index along `axis`.  Every element decodes its nd indices: still one 16 B read per
// element, HBM bound.
struct LatticeGeom {
  int nd;
  int grid[SQB_MAX_DIMS];
  double row[SQB_MAX_DIMS];   // A[a][axis]
  double finv[SQB_MAX_DIMS];  // ((A^T)^-1)[a][axis]
};

__global__ void __launch_bounds__(kMeasureThreads)
measure_moments_lattice_kernel(int64_t m_total, LatticeGeom g, int axis, const double* __restrict__ offsets,
                               int wrapped, const c128* __restrict__ states, double* __restrict__ out) {
  __shared__ double red[5][kMeasureThreads / 32];
  const int64_t t = blockIdx.x;
  const double off = offsets ? offsets[t] : 0.0;
  const c128* row = states + t * m_total;
  double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0, a4 = 0.0;
  for (int64_t m = threadIdx.x; m < m_total; m += kMeasureThreads) {
    const c128 z = row[m];
    const double w = fma(z.x, z.x, z.y * z.y);
    int64_t rem = m;
    double x = 0.0;
    int i_axis = 0;
#pragma unroll
    for (int a = SQB_MAX_DIMS - 1; a >= 0; --a) {
      if (a < g.nd) {
        const int ia = (int)(rem % g.grid[a]);
        rem /= g.grid[a];
        if (a == axis) i_axis = ia;
        double frac = (double)ia / (double)g.grid[a] + off * g.finv[a];
        if (wrapped) frac -= floor(frac + 0.5);
        x = fma(frac, g.row[a], x);
      }
    }
    double sv, c;
    sincospi(2.0 * (double)i_axis / (double)g.grid[axis], &sv, &c);
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

extern "C" int sqb_measure_moments_lattice(int nd, const int* grid_host, int64_t lead, const sqb_c128* states, int axis,
                                           const double* delta_x_host, const double* offsets, int wrapped,
                                           double* out, void* stream) {
  if (nd < 1 || nd > SQB_MAX_DIMS || !grid_host || !states || !out || !delta_x_host || axis < 0 || axis >= nd ||
      lead < 0)
    return SQB_E_BADARG;
  LatticeGeom g;
  g.nd = nd;
  int64_t m_total = 1;
  for (int a = 0; a < SQB_MAX_DIMS; ++a) {
    g.grid[a] = a < nd ? grid_host[a] : 1;
    g.row[a] = 0.0;
    g.finv[a] = 0.0;
    if (g.grid[a] < 1) return SQB_E_BADARG;
    m_total *= g.grid[a];
  }
  // finv = solution of A^T f = e_axis (Gauss-Jordan with partial pivoting on the nd x nd system, host side)
  double mat[SQB_MAX_DIMS][SQB_MAX_DIMS + 1];
  for (int i = 0; i < nd; ++i) {
    for (int a = 0; a < nd; ++a) mat[i][a] = delta_x_host[a * nd + i];  // (A^T)[i][a] = A[a][i]
    mat[i][nd] = i == axis ? 1.0 : 0.0;
  }
  for (int col = 0; col < nd; ++col) {
    int piv = col;
    for (int i = col + 1; i < nd; ++i)
      if (fabs(mat[i][col]) > fabs(mat[piv][col])) piv = i;
    if (mat[piv][col] == 0.0) return SQB_E_BADARG;  // singular cell
    for (int k = 0; k <= nd; ++k) {
      const double tmp = mat[col][k];
      mat[col][k] = mat[piv][k];
      mat[piv][k] = tmp;
    }
    for (int i = 0; i < nd; ++i) {
      if (i == col) continue;
      const double f = mat[i][col] / mat[col][col];
      for (int k = col; k <= nd; ++k) mat[i][k] -= f * mat[col][k];
    }
  }
  for (int a = 0; a < nd; ++a) {
    g.finv[a] = mat[a][nd] / mat[a][a];
    g.row[a] = delta_x_host[a * nd + axis];
  }
  if (lead == 0) return SQB_OK;
  if (lead >= (1LL << 31)) return SQB_E_UNSUPPORTED;
  measure_moments_lattice_kernel<<<(unsigned)lead, kMeasureThreads, 0, sqb_stream(stream)>>>(
      m_total, g, axis, offsets, wrapped, reinterpret_cast<const c128*>(states), out);
  SQB_LAUNCH_CHECK();
  return SQB_OK;
}

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
