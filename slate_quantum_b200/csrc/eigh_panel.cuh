// K2, stage 1 for 64 < n <= 256: BLOCKED Householder tridiagonalisation (LAPACK zhetrd / zlatrd, lower triangle),
// one CTA (512 threads) per matrix.  Numpy model: tests/eigh_model.py::tridiagonalise_blocked.
//
// Same conventions and outputs as tridiag_kernel (eigh.cu): the reduction runs in the column-major view of the
// row-major input (i.e. on conj(H); element (r, c) of that view is A[c * n + r]), reflector i goes to Y[i * n + r]
// (r >= i + 1, v[i + 1] = 1), d / e / tau as zhetd2.
//
// Inside a panel of NB columns the trailing matrix is only READ:
//   * column i is materialised from the panel's pending update  a -= V conj(W[i, :]) + W conj(V[i, :]);
//   * p = A22 v uses the matrix of the panel start (16 B read per lower-triangle element and column instead of a
//     16 B read + 16 B write) and is corrected with the panel's V and W (two thin products);
//   * the rank-2NB update  A22 -= V W^H + W V^H  runs ONCE per panel as a complex GEMM on the FP64 tensor cores
//     (DMMA.8x8x4, operands = the V / W panels in shared memory, accumulators in registers in the transposed
//     fragment layout so that a lane touches 32 contiguous bytes of a column).
// The trailing columns c >= jres live in shared memory for the whole kernel (packed lower triangle, every column
// start aligned to 128 B); the leading columns are streamed from L2 and only touched by the first jres steps.
//
// Matrix-vector product.  The lower triangle is cut into 32 x 32 units (row chunk Rc, column block Cb <= Rc).  A
// lane owns a 4 x 8 sub-tile of a unit (rows i + 8 rho, columns j + 4 q; i = lane & 7, j = lane >> 3), so both
// halves of the Hermitian product accumulate in registers - y[rho] += a v[c] and z[q] += conj(a) v[r] - and only
// 4 + 8 partial sums per lane are combined across lanes (transpose-reduce butterflies: 3 exchange steps for y
// per unit, 7 for z per strip segment) instead of one 32-lane reduction per column.  Units are dealt to the 16
// warps as contiguous ranges of the column-major unit list, weighted (diagonal units count half).
#pragma once
#include "sqb_common.cuh"

#ifdef SQB_PANEL_PROFILE
#include <stdio.h>
#endif

namespace {

#ifdef SQB_PANEL_PROFILE
// developer build only (nvcc -DSQB_PANEL_PROFILE): cycles of block 0 per phase, printed by the launcher
__device__ unsigned long long pn_prof[16];
__device__ unsigned long long pn_prof2[4];
__device__ unsigned long long pn_prof3[32];  // per warp: cycles inside the units / whole of phase (b)
#define PN_TICK(slot)                                          \
  do {                                                         \
    if (blockIdx.x == 0 && threadIdx.x == 0) {                 \
      const long long now__ = clock64();                       \
      pn_prof[slot] += (unsigned long long)(now__ - pn_t0__);  \
      pn_t0__ = now__;                                         \
    }                                                          \
  } while (0)
#else
#define PN_TICK(slot) \
  do {                \
  } while (0)
#endif

#ifndef SQB_PN_THREADS
#define SQB_PN_THREADS 256
#endif
constexpr int kPnThreads = SQB_PN_THREADS;          // 256: one CTA per SM (255 registers); 128: two CTAs per SM
constexpr int kPnCtasPerSm = 256 / kPnThreads;
constexpr int kPnWarps = kPnThreads / 32;
constexpr unsigned kPnFull = 0xffffffffu;
constexpr int kPnMaxN = 512;
// cost model of the matrix-vector product's work list, in quarter-kilocycles (see phase (b))
#ifndef SQB_PN_W_SF
#define SQB_PN_W_SF 4   // streamed full unit
#define SQB_PN_W_SD 4   // streamed diagonal unit
#define SQB_PN_W_RF 4   // resident full unit
#define SQB_PN_W_RD 4   // resident diagonal unit
#define SQB_PN_W_DOT 1  // one correction dot product
#endif
constexpr int kPnWDot = SQB_PN_W_DOT;
__device__ __forceinline__ int pn_wfull(bool resident) { return resident ? SQB_PN_W_RF : SQB_PN_W_SF; }
__device__ __forceinline__ int pn_wdiag(bool resident) { return resident ? SQB_PN_W_RD : SQB_PN_W_SD; }

__device__ __forceinline__ void pn_dmma(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(d0), "+d"(d1)
               : "d"(a), "d"(b));
}

struct PanelLayout {  // byte offsets into the dynamic shared memory, identical on host and device
  int n8, ldp, rc;
  size_t vp, wp, red, tmp, scal, cbase, res;
};

__host__ __device__ inline PanelLayout panel_layout(int n, int nb) {
  PanelLayout L;
  L.n8 = (n + 7) & ~7;
  L.ldp = L.n8 + 2;  // panel row stride == 2 (mod 8) complex: conflict-free DMMA fragment loads
  L.rc = (n + 31) / 32;
  size_t o = 0;
  L.vp = o;
  o += (size_t)nb * L.ldp * 16;
  L.wp = o;
  o += (size_t)nb * L.ldp * 16;
  L.red = o;  // [kPnWarps][n8]: every warp's partial sums of the matrix-vector product (only that warp writes it)
  o += (size_t)(kPnThreads / 32) * L.n8 * 16;
  L.tmp = o;  // [2][nb] W^H v, V^H v
  o += (size_t)2 * nb * 16;
  L.scal = o;  // 128 doubles: reduction scratch + scalars
  o += 128 * 8;
  L.cbase = o;  // [n8] ints
  o += (size_t)L.n8 * 4;
  o = (o + 127) & ~(size_t)127;
  L.res = o;
  return L;
}

// padded length of resident column c: rows floor8(c) .. n8-1
__host__ __device__ inline int panel_col_len(int n8, int c) { return n8 - (c & ~7); }

// block-wide sums with ONE barrier: every warp publishes its partial sum, then every thread adds the 16 partials
// itself in the same order (identical result in all threads, no single-warp second stage).  `buf` (>= 32 doubles)
// must not be the buffer of the previous reduction unless a barrier separates the two.
__device__ __forceinline__ double pn_block_sum(double v, double* buf) {
  v = warp_sum(v);
  if ((threadIdx.x & 31) == 0) buf[threadIdx.x >> 5] = v;
  __syncthreads();
  double t = 0.0;
#pragma unroll
  for (int w = 0; w < kPnWarps; ++w) t += buf[w];
  return t;
}
__device__ __forceinline__ double2 pn_block_sum2(double a, double b, double* buf) {
  a = warp_sum(a);
  b = warp_sum(b);
  if ((threadIdx.x & 31) == 0) {
    buf[threadIdx.x >> 5] = a;
    buf[kPnWarps + (threadIdx.x >> 5)] = b;
  }
  __syncthreads();
  double ta = 0.0, tb = 0.0;
#pragma unroll
  for (int w = 0; w < kPnWarps; ++w) {
    ta += buf[w];
    tb += buf[kPnWarps + w];
  }
  return make_double2(ta, tb);
}

// one transpose-reduce exchange: lanes l and l ^ off combine; the lane whose bit `b` is 0 keeps `lo` and sends
// `hi`, its partner keeps `hi` and sends `lo`.  Result (own kept value + partner's sent value) in `lo`.
__device__ __forceinline__ void pn_xchg(c128& lo, const c128& hi, bool b, int off) {
  c128 send, keep;
  send.x = b ? lo.x : hi.x;
  send.y = b ? lo.y : hi.y;
  keep.x = b ? hi.x : lo.x;
  keep.y = b ? hi.y : lo.y;
  send.x = __shfl_xor_sync(kPnFull, send.x, off);
  send.y = __shfl_xor_sync(kPnFull, send.y, off);
  lo.x = keep.x + send.x;
  lo.y = keep.y + send.y;
}

// 16-byte PREDICATED loads as volatile asm (zero when the predicate is off).  The compiler keeps volatile asm in
// program order, so a unit issues ALL its 32 loads back to back before the arithmetic (one exposed L2 / shared
// memory latency per unit instead of one per column); masked elements cost no bandwidth and need no select.
__device__ __forceinline__ c128 pn_ldg(const c128* p, bool pred) {
  c128 v;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %3, 0;\n"
      "mov.f64 %0, 0d0000000000000000;\n"
      "mov.f64 %1, 0d0000000000000000;\n"
      "@p ld.global.v2.f64 {%0, %1}, [%2];\n"
      "}\n"
      : "=d"(v.x), "=d"(v.y)
      : "l"(p), "r"((int)pred));
  return v;
}
__device__ __forceinline__ c128 pn_lds(const c128* p, bool pred) {
  c128 v;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %3, 0;\n"
      "mov.f64 %0, 0d0000000000000000;\n"
      "mov.f64 %1, 0d0000000000000000;\n"
      "@p ld.shared.v2.f64 {%0, %1}, [%2];\n"
      "}\n"
      : "=d"(v.x), "=d"(v.y)
      : "r"((uint32_t)__cvta_generic_to_shared(p)), "r"((int)pred));
  return v;
}

// One 32 x 32 unit of the Hermitian matrix-vector product (rows r0.., columns c0..) for one warp.
// Lane (li = lane & 7, lj = lane >> 3) owns rows r0 + li + 8 rho (rho < 4) and columns c0 + lj + 4 q (q < 8):
//   y[rho] += a v[c]   (this unit's contribution to p[rows]; returned reduced over the 4 column lanes: the lane
//                       holds the sum for row r0 + li + 8 (2 b3 + b4), b3 / b4 = lane bits 3 / 4)
//   z[q]   += conj(a) v[r]   (mirrored half, rows > columns; accumulated in the caller's registers over a strip)
// RES: the column block is resident in shared memory (packed lower triangle) - otherwise streamed from global
// memory.  Elements outside the trailing lower triangle are never loaded (predicated loads returning 0).
template <bool RES, bool DIAG>
__device__ __forceinline__ c128 pn_unit(int n, int lo, int r0, int c0, int lane, const c128* __restrict__ vn,
                                        const c128* res, const int* cbase, const c128* A, c128 (&z)[8]) {
  const int li = lane & 7, lj = lane >> 3;
  const int rb = r0 + li;
  c128 a[8][4];
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const int c = c0 + lj + 4 * q;
    const bool cok = c >= lo && c < n;
    const int cc = min(c, n - 1);
    const c128* colp = RES ? res + cbase[cc] : A + (int64_t)cc * n;
#pragma unroll
    for (int rho = 0; rho < 4; ++rho) {
      const int rr = rb + 8 * rho;
      const bool ok = cok && rr < n && (!DIAG || rr >= c);  // masked elements: no load, value 0
      a[q][rho] = RES ? pn_lds(colp + rr, ok) : pn_ldg(colp + rr, ok);
    }
  }
  c128 vr[4], y[4];
#pragma unroll
  for (int rho = 0; rho < 4; ++rho) {
    vr[rho] = vn[min(rb + 8 * rho, n - 1)];
    y[rho] = make_double2(0.0, 0.0);
  }
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const int c = c0 + lj + 4 * q;
    const c128 vc = vn[min(c, n - 1)];
#pragma unroll
    for (int rho = 0; rho < 4; ++rho) {
      c128 e = a[q][rho];
      cfma(y[rho], e, vc);
      if (DIAG && rb + 8 * rho == c) e = make_double2(0.0, 0.0);  // the diagonal has no mirrored partner
      cfmac(z[q], e, vr[rho]);
    }
  }
  // y: 4 values per lane, summed over the 4 column lanes lj (lane bits 3, 4) -> one row per lane
  const bool b3 = (lane >> 3) & 1, b4 = (lane >> 4) & 1;
  pn_xchg(y[0], y[2], b3, 8);
  pn_xchg(y[1], y[3], b3, 8);
  pn_xchg(y[0], y[1], b4, 16);
  return y[0];
}

// Off-diagonal unit (all rows > all columns): NO per-element predicates.  Addresses are clamped into the live part
// of the matrix (columns to [lo, n), rows to < n; a clamped lane re-reads an element another lane needs, so no extra
// traffic) and the masking rides on v: vn[c] is zero for c < lo and for the padding rows >= n (index clamped to n8,
// a padding entry), so dead columns add nothing to y, dead rows nothing to z, and the sums a dead lane accumulates
// are never stored.  Unconditional volatile loads: no destination zeroing, no predicate set-up (~half the
// instructions of the predicated unit, which matters with two warps per scheduler).
__device__ __forceinline__ c128 pn_ldg_u(const c128* p) {
  c128 v;
  asm volatile("ld.global.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
  return v;
}
__device__ __forceinline__ c128 pn_lds_u(const c128* p) {
  c128 v;
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"((uint32_t)__cvta_generic_to_shared(p)));
  return v;
}
template <bool RES>
__device__ __forceinline__ c128 pn_unit_full(int n, int n8, int lo, int r0, int c0, int lane,
                                             const c128* __restrict__ vn, const c128* res, const int* cbase,
                                             const c128* A, c128 (&z)[8]) {
  const int li = lane & 7, lj = lane >> 3;
  const int rb = r0 + li;
  int roff[4];
#pragma unroll
  for (int rho = 0; rho < 4; ++rho) roff[rho] = min(rb + 8 * rho, n - 1);
  c128 a[8][4];
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const int cc = min(max(c0 + lj + 4 * q, lo), n - 1);
    const c128* colp = RES ? res + cbase[cc] : A + (int64_t)cc * n;
#pragma unroll
    for (int rho = 0; rho < 4; ++rho) a[q][rho] = RES ? pn_lds_u(colp + roff[rho]) : pn_ldg_u(colp + roff[rho]);
  }
  c128 vr[4], y[4];
#pragma unroll
  for (int rho = 0; rho < 4; ++rho) {
    vr[rho] = vn[min(rb + 8 * rho, n8)];
    y[rho] = make_double2(0.0, 0.0);
  }
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const c128 vc = vn[min(c0 + lj + 4 * q, n8)];
#pragma unroll
    for (int rho = 0; rho < 4; ++rho) {
      cfma(y[rho], a[q][rho], vc);
      cfmac(z[q], a[q][rho], vr[rho]);
    }
  }
  const bool b3 = (lane >> 3) & 1, b4 = (lane >> 4) & 1;
  pn_xchg(y[0], y[2], b3, 8);
  pn_xchg(y[1], y[3], b3, 8);
  pn_xchg(y[0], y[1], b4, 16);
  return y[0];
}

template <int NB, int NR>
__global__ void __launch_bounds__(kPnThreads, kPnCtasPerSm)
tridiag_panel_kernel(int n, int jres, c128* __restrict__ A_all, double* __restrict__ d_all, double* __restrict__ e_all,
                     c128* __restrict__ tau_all, c128* __restrict__ Y_all) {
  extern __shared__ __align__(128) unsigned char pn_smem[];
  const PanelLayout L = panel_layout(n, NB);
  const int n8 = L.n8, ldp = L.ldp, RC = L.rc;
  c128* Vp = reinterpret_cast<c128*>(pn_smem + L.vp);   // [NB][ldp]  Vp[k][r]
  c128* Wp = reinterpret_cast<c128*>(pn_smem + L.wp);   // [NB][ldp]
  c128* part_w = reinterpret_cast<c128*>(pn_smem + L.red) + (size_t)(threadIdx.x >> 5) * n8;  // this warp's slot
  const c128* part_all = reinterpret_cast<const c128*>(pn_smem + L.red);
  c128* tmp1 = reinterpret_cast<c128*>(pn_smem + L.tmp);  // W^H v
  c128* tmp2 = tmp1 + NB;                                  // V^H v
  double* scratch = reinterpret_cast<double*>(pn_smem + L.scal);  // [0..71] reductions, [80..] scalars
  int* cbase = reinterpret_cast<int*>(pn_smem + L.cbase);
  c128* res = reinterpret_cast<c128*>(pn_smem + L.res);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t b = blockIdx.x;
  c128* A = A_all + b * (int64_t)n * n;
  c128* Y = Y_all + b * (int64_t)n * n;
  double* dd = d_all + b * n;
  double* ee = e_all + b * n;
  c128* tt = tau_all + b * n;

  // ---- set-up: resident column offsets, panels zeroed (rows >= n of a panel row stay zero for ever)
  for (int c = tid; c < n8; c += kPnThreads) {
    int v = -1;
    if (c >= jres && c < n) {
      const int g0 = jres >> 3, G = c >> 3;
      const int S = 8 * ((G - g0) * n8 - 8 * (G * (G - 1) / 2 - g0 * (g0 - 1) / 2)) + (c - 8 * G) * (n8 - 8 * G);
      v = S - 8 * G;  // element (r, c) sits at res[cbase[c] + r]
    }
    cbase[c] = v;
  }
  for (int idx = tid; idx < 2 * NB * ldp; idx += kPnThreads) Vp[idx] = make_double2(0.0, 0.0);  // Vp and Wp are adjacent
  __syncthreads();
  for (int c = jres + warp; c < n; c += kPnWarps) {
    c128* dst = res + cbase[c];
    const c128* src = A + (int64_t)c * n;
    for (int r = (c & ~7) + lane; r < n; r += 32) dst[r] = src[r];
  }
  __syncthreads();

#ifdef SQB_PANEL_PROFILE
  long long pn_t0__ = clock64();
#endif
  PN_TICK(0);
#if defined(SQB_PANEL_PROFILE) && SQB_PANEL_PROFILE > 1
  long long pb_setup = 0, pb_flush = 0, pb_dots = 0, pb_units = 0, pb_t = 0, pb_count = 0;
#define PB_START() pb_t = clock64()
#define PB_ADD(var) do { const long long n__ = clock64(); var += n__ - pb_t; pb_t = n__; } while (0)
#else
#define PB_START() do {} while (0)
#define PB_ADD(var) do {} while (0)
#endif
  // column phases: thread t owns rows t + kPnThreads * s, s < NR  (n <= NR * kPnThreads)
  c128 a_pref[NR];
#pragma unroll
  for (int s = 0; s < NR; ++s) a_pref[s] = make_double2(0.0, 0.0);
  for (int i0 = 0; i0 < n - 1; i0 += NB) {
    const int width = min(NB, n - 1 - i0);
    if (width < NB) {  // last, narrower panel: the unused panel rows must not carry the previous panel
      for (int idx = tid; idx < (NB - width) * ldp; idx += kPnThreads) {
        Vp[width * ldp + idx] = make_double2(0.0, 0.0);
        Wp[width * ldp + idx] = make_double2(0.0, 0.0);
      }
      __syncthreads();
    }
    for (int jj = 0; jj < width; ++jj) {
      const int i = i0 + jj;
      const int lo = i + 1;
      c128* vn = Vp + jj * ldp;
      // ---------------- (a) column i with the panel's pending update; Householder reflector
      c128 colv[NR];
      double part = 0.0;
#pragma unroll
      for (int s = 0; s < NR; ++s) {
        const int r = tid + s * kPnThreads;
        colv[s] = make_double2(0.0, 0.0);
        if (r >= i && r < n) {
          // column i was fetched during the previous step's matrix-vector product (the matrix is read-only inside
          // a panel); the first column of a panel is fetched here, after the trailing update
          c128 a = jj > 0 ? a_pref[s] : (i >= jres ? res[cbase[i] + r] : A[(int64_t)i * n + r]);
          c128 t = make_double2(0.0, 0.0);
#pragma unroll
          for (int k = 0; k < NB - 1; ++k) {
            if (k < jj) {
              const c128 cw = cconj(Wp[k * ldp + i]), cv = cconj(Vp[k * ldp + i]);
              cfma(t, Vp[k * ldp + r], cw);
              cfma(t, Wp[k * ldp + r], cv);
            }
          }
          a = csub(a, t);
          colv[s] = a;
          if (r >= i + 2) part += a.x * a.x + a.y * a.y;
          if (r == i) dd[i] = a.x;
          if (r == i + 1) {
            scratch[80] = a.x;
            scratch[81] = a.y;
          }
        }
      }
      const double xnorm2 = pn_block_sum(part, scratch);  // scratch[0..15]; also publishes scratch[80..81]
      const double ar = scratch[80], ai = scratch[81];
      double beta, tau_r, tau_i;
      c128 scal;
      if (xnorm2 == 0.0 && ai == 0.0) {
        beta = ar;
        tau_r = 0.0;
        tau_i = 0.0;
        scal = make_double2(0.0, 0.0);
      } else {
        beta = -copysign(sqrt(ar * ar + ai * ai + xnorm2), ar);
        tau_r = (beta - ar) / beta;
        tau_i = -ai / beta;
        const double pr = ar - beta, pi = ai;
        const double rden = 1.0 / (pr * pr + pi * pi);
        scal = make_double2(pr * rden, -pi * rden);
      }
      const c128 tau = make_double2(tau_r, tau_i);
      c128 myv[NR];
#pragma unroll
      for (int s = 0; s < NR; ++s) {
        const int r = tid + s * kPnThreads;
        myv[s] = make_double2(0.0, 0.0);
        if (r < n) {
          if (r == i + 1) myv[s] = make_double2(1.0, 0.0);
          else if (r >= i + 2) myv[s] = cmul(colv[s], scal);
          vn[r] = myv[s];
          if (r >= i + 1) Y[(int64_t)i * n + r] = myv[s];
        }
      }
      if (tid == 0) {
        ee[i] = beta;
        tt[i] = tau;
      }
      __syncthreads();
      PN_TICK(1);

      // ---------------- (b) p = A22 v from the panel-start matrix (read only) + the correction dot products
      {
        PB_START();
        // next column of the panel for phase (a) of the next step (rows >= lo; hidden behind the units below)
#pragma unroll
        for (int s = 0; s < NR; ++s) {
          const int r = tid + s * kPnThreads;
          const bool want = jj + 1 < width && r >= lo && r < n;
          if (lo >= jres) a_pref[s] = pn_lds(res + cbase[min(lo, n - 1)] + min(r, n - 1), want);
          else a_pref[s] = pn_ldg(A + (int64_t)lo * n + min(r, n - 1), want);
        }
        // this warp's partial-sum slot, zeroed over the trailing range
        for (int x = (lo & ~31) + lane; x < n8; x += 32) part_w[x] = make_double2(0.0, 0.0);
        __syncwarp();
        const int X0 = lo >> 5;
        // Units are dealt to the warps as contiguous ranges of the column-major unit list, by COST: a unit is
        // latency bound, not flop bound, so a diagonal unit costs as much as a full one (clock64, N = 256: streamed
        // full 2.5 k cycles, streamed diagonal 2.8 k, resident full 1.5 k, resident diagonal 1.9 k, one correction
        // dot product ~0.5 k); the 2 jj correction dot products are the tail of the same list.
        int W = 0;
        for (int Cb = X0; Cb < RC; ++Cb) W += pn_wdiag(Cb * 32 >= jres) + (RC - Cb - 1) * pn_wfull(Cb * 32 >= jres);
        const int W_units = W;
        W += 2 * jj * kPnWDot;
        const int w_lo = (warp * W + kPnWarps - 1) / kPnWarps, w_hi = ((warp + 1) * W + kPnWarps - 1) / kPnWarps;
        int pos = 0;
        PB_ADD(pb_setup);
        for (int Cb = X0; Cb < RC && pos < w_hi; ++Cb) {
          const int wd = pn_wdiag(Cb * 32 >= jres), wf = pn_wfull(Cb * 32 >= jres);
          const int strip = wd + (RC - Cb - 1) * wf;
          if (pos + strip <= w_lo) {
            pos += strip;
            continue;
          }
          const int c0 = Cb * 32;
          bool open = false;
          c128 z[8];
#pragma unroll
          for (int q = 0; q < 8; ++q) z[q] = make_double2(0.0, 0.0);
          // first unit of the strip whose start position is >= w_lo: unit Rc starts at pos (Rc == Cb) or at
          // pos + wd + (Rc - Cb - 1) wf
          int Rc = Cb, start = pos;
          if (pos < w_lo) {
            const int k = max(0, (w_lo - pos - wd + wf - 1) / wf);
            Rc = Cb + 1 + k;
            start = pos + wd + k * wf;
          }
          pos += strip;
          for (; Rc < RC && start < w_hi; ++Rc) {
            start += (Rc == Cb) ? wd : wf;
            open = true;
#if defined(SQB_PANEL_PROFILE) && SQB_PANEL_PROFILE > 1
            ++pb_count;
#endif
            c128 y;
            const int r0 = Rc * 32;
            if (c0 >= jres) {
              if (Rc == Cb) y = pn_unit<true, true>(n, lo, r0, c0, lane, vn, res, cbase, A, z);
              else y = pn_unit_full<true>(n, n8, lo, r0, c0, lane, vn, res, cbase, A, z);
            } else {
              if (Rc == Cb) y = pn_unit<false, true>(n, lo, r0, c0, lane, vn, res, cbase, A, z);
              else y = pn_unit_full<false>(n, n8, lo, r0, c0, lane, vn, res, cbase, A, z);
            }
            const int rr = Rc * 32 + (lane & 7) + 8 * (((lane >> 3) & 1) * 2 + ((lane >> 4) & 1));
            if (rr < n) part_w[rr] = cadd(part_w[rr], y);
            __syncwarp();
          }
          PB_ADD(pb_units);
          if (open) {
            // z: 8 values per lane, summed over the 8 row lanes li (lane bits 0, 1, 2) -> one column per lane
            const bool b0 = lane & 1, b1 = (lane >> 1) & 1, b2 = (lane >> 2) & 1;
            pn_xchg(z[0], z[4], b0, 1);
            pn_xchg(z[1], z[5], b0, 1);
            pn_xchg(z[2], z[6], b0, 1);
            pn_xchg(z[3], z[7], b0, 1);
            pn_xchg(z[0], z[2], b1, 2);
            pn_xchg(z[1], z[3], b1, 2);
            pn_xchg(z[0], z[1], b2, 4);
            const int c = c0 + (lane >> 3) + 4 * ((b0 ? 4 : 0) + (b1 ? 2 : 0) + (b2 ? 1 : 0));
            if (c < n) part_w[c] = cadd(part_w[c], z[0]);
            __syncwarp();
          }
          PB_ADD(pb_flush);
        }
        PB_ADD(pb_units);
        // correction dot products: tmp1[k] = W_k^H v, tmp2[k] = V_k^H v over rows >= lo
        for (int q = 0; q < 2 * jj; ++q) {
          const int start = W_units + q * kPnWDot;
          if (start < w_lo || start >= w_hi) continue;
          const int k = q >> 1;
          const c128* P = (q & 1) ? Vp + k * ldp : Wp + k * ldp;
          c128 acc = make_double2(0.0, 0.0);
          for (int rr = lo + lane; rr < n; rr += 32) cfmac(acc, P[rr], vn[rr]);
          acc.x = warp_sum(acc.x);
          acc.y = warp_sum(acc.y);
          if (lane == 0) ((q & 1) ? tmp2 : tmp1)[k] = acc;
        }
        PB_ADD(pb_dots);
      }
      PN_TICK(2);
      __syncthreads();
      PN_TICK(3);

      // ---------------- (d) w = tau (p - V tmp1 - W tmp2) - (tau/2)(p^H v) v
      c128 p[NR];
      double dr = 0.0, di = 0.0;
#pragma unroll
      for (int s = 0; s < NR; ++s) {
        const int r = tid + s * kPnThreads;
        p[s] = make_double2(0.0, 0.0);
        if (r >= lo && r < n) {
          c128 sum = part_all[r];
#pragma unroll
          for (int w = 1; w < kPnWarps; ++w) sum = cadd(sum, part_all[(size_t)w * n8 + r]);
          c128 t = make_double2(0.0, 0.0);
#pragma unroll
          for (int k = 0; k < NB - 1; ++k) {
            if (k < jj) {
              cfma(t, Vp[k * ldp + r], tmp1[k]);
              cfma(t, Wp[k * ldp + r], tmp2[k]);
            }
          }
          p[s] = cmul(tau, csub(sum, t));
          dr += p[s].x * myv[s].x + p[s].y * myv[s].y;  // conj(p) * v
          di += p[s].x * myv[s].y - p[s].y * myv[s].x;
        }
      }
      const double2 dot = pn_block_sum2(dr, di, scratch + 32);  // scratch[32..63]
      const c128 al = cscale(cmul(tau, dot), -0.5);
#pragma unroll
      for (int s = 0; s < NR; ++s) {
        const int r = tid + s * kPnThreads;
        if (r < n) {
          c128 w = make_double2(0.0, 0.0);
          if (r >= lo) w = cadd(p[s], cmul(al, myv[s]));
          Wp[jj * ldp + r] = w;
        }
      }
      __syncthreads();
      PN_TICK(4);
    }

    // ---------------- (e) trailing update A22 -= V W^H + W V^H on the FP64 tensor cores (rows >= cols >= t0)
    const int t0 = i0 + width;
    {
      const int g = lane >> 2, q = lane & 3;
      const int u0 = t0 >> 4, U = (n + 15) >> 4;
      int idx = 0;
      for (int u = u0; u < U; ++u) {
        for (int s = u >> 1; s < RC; ++s, ++idx) {
          if ((idx & (kPnWarps - 1)) != warp) continue;
          const int c0 = u * 16, r0 = s * 32;
          double cre[2][4][2], cim[2][4][2];
#pragma unroll
          for (int s2 = 0; s2 < 2; ++s2)
#pragma unroll
            for (int t4 = 0; t4 < 4; ++t4) {
              cre[s2][t4][0] = cre[s2][t4][1] = 0.0;
              cim[s2][t4][0] = cim[s2][t4][1] = 0.0;
            }
          // the C tile is fetched BEFORE the MMA loop (predicated volatile loads keep their place in the instruction
          // stream), so its L2 / shared-memory latency hides behind the 128 DMMAs instead of following them
          c128 cpre[2][4][2];
          c128* ccol[2];
#pragma unroll
          for (int s2 = 0; s2 < 2; ++s2) {
            const int c = c0 + 8 * s2 + g;
            const bool cok = c >= t0 && c < n;
            const bool resident = c0 + 8 * s2 >= jres;
            const int cc = min(c, n - 1);
            ccol[s2] = resident ? res + cbase[cc] : A + (int64_t)cc * n;
#pragma unroll
            for (int t4 = 0; t4 < 4; ++t4)
#pragma unroll
              for (int e = 0; e < 2; ++e) {
                const int rr = r0 + 8 * t4 + 2 * q + e;
                const bool ok = cok && rr >= c && rr < n;
                cpre[s2][t4][e] = resident ? pn_lds(ccol[s2] + rr, ok) : pn_ldg(ccol[s2] + rr, ok);
              }
          }
#pragma unroll
          for (int kc = 0; kc < (2 * NB) / 4; ++kc) {
            // L = [V W] (row side), R = [W V] (column side)
            const c128* Rs = (4 * kc < NB ? Wp : Vp) + ((4 * kc) % NB + q) * ldp;
            const c128* Ls = (4 * kc < NB ? Vp : Wp) + ((4 * kc) % NB + q) * ldp;
            double ar[2], ai[2], nai[2], br[4], bi[4];
#pragma unroll
            for (int s2 = 0; s2 < 2; ++s2) {
              const c128 v = Rs[min(c0 + 8 * s2 + g, n8 - 1)];
              ar[s2] = v.x;
              ai[s2] = v.y;
              nai[s2] = -v.y;
            }
#pragma unroll
            for (int t4 = 0; t4 < 4; ++t4) {
              const c128 v = Ls[min(r0 + 8 * t4 + g, n8 - 1)];
              br[t4] = v.x;
              bi[t4] = v.y;
            }
#pragma unroll
            for (int s2 = 0; s2 < 2; ++s2)
#pragma unroll
              for (int t4 = 0; t4 < 4; ++t4) {
                pn_dmma(cre[s2][t4][0], cre[s2][t4][1], ar[s2], br[t4]);
                pn_dmma(cre[s2][t4][0], cre[s2][t4][1], ai[s2], bi[t4]);
                pn_dmma(cim[s2][t4][0], cim[s2][t4][1], ar[s2], bi[t4]);
                pn_dmma(cim[s2][t4][0], cim[s2][t4][1], nai[s2], br[t4]);
              }
          }
#pragma unroll
          for (int s2 = 0; s2 < 2; ++s2) {
            const int c = c0 + 8 * s2 + g;
            if (c < t0 || c >= n) continue;
#pragma unroll
            for (int t4 = 0; t4 < 4; ++t4) {
#pragma unroll
              for (int e = 0; e < 2; ++e) {
                const int rr = r0 + 8 * t4 + 2 * q + e;
                if (rr >= c && rr < n) {
                  c128 a = cpre[s2][t4][e];
                  a.x -= cre[s2][t4][e];
                  a.y -= cim[s2][t4][e];
                  ccol[s2][rr] = a;
                }
              }
            }
          }
        }
      }
    }
    __syncthreads();
    PN_TICK(5);
  }
#if defined(SQB_PANEL_PROFILE) && SQB_PANEL_PROFILE > 1
  if (blockIdx.x == 0 && tid == 0) {
    pn_prof[14] += (unsigned long long)pb_setup;
    pn_prof[15] += (unsigned long long)pb_flush;
    pn_prof[4 + 8] += 0;
    pn_prof[3] += 0;
    atomicAdd(&pn_prof2[0], (unsigned long long)pb_units);
    atomicAdd(&pn_prof2[1], (unsigned long long)pb_dots);
  }
  if (blockIdx.x == 0 && lane == 0) {
    pn_prof3[warp] += (unsigned long long)pb_units;
    pn_prof3[8 + warp] += (unsigned long long)(pb_setup + pb_flush + pb_dots);
    pn_prof3[16 + warp] += (unsigned long long)pb_count;
  }
#endif
  if (tid == 0) {
    const int l = n - 1;
    const c128 a = l >= jres ? res[cbase[l] + l] : A[(int64_t)l * n + l];
    dd[l] = a.x;
    ee[l] = 0.0;
    tt[l] = make_double2(0.0, 0.0);
  }
}

// resident columns: the largest multiple-of-8 suffix of columns that fits beside the fixed arrays
inline int panel_jres(int n, int nb, size_t budget, size_t* smem_out) {
  const PanelLayout L = panel_layout(n, nb);
  int jres = L.rc * 32;  // multiples of 32: every 32 x 32 unit is wholly resident or wholly streamed
  size_t elems = 0;
  while (jres >= 32) {
    size_t add = 0;
    for (int c = jres - 32; c < jres && c < n; ++c) add += (size_t)panel_col_len(L.n8, c);
    if (L.res + (elems + add) * 16 > budget) break;
    elems += add;
    jres -= 32;
  }
  *smem_out = L.res + elems * 16;
  return jres;
}

template <int NB, int NR>
int launch_tridiag_panel(int n, int64_t count, c128* A, double* d, double* e, c128* tau, c128* y, cudaStream_t st) {
  size_t smem = 0;
  const int jres = panel_jres(n, NB, (size_t)(233472 / kPnCtasPerSm) - 1024 - (kPnCtasPerSm > 1 ? 1024 : 0), &smem);
  cudaError_t err =
      cudaFuncSetAttribute(tridiag_panel_kernel<NB, NR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (err != cudaSuccess) return (int)err;
  tridiag_panel_kernel<NB, NR><<<(unsigned)count, kPnThreads, smem, st>>>(n, jres, A, d, e, tau, y);
  SQB_LAUNCH_CHECK();
#ifdef SQB_PANEL_PROFILE
  {
    unsigned long long h[16];
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(h, pn_prof, sizeof(h));
    fprintf(stderr, "[panel NB=%d n=%d jres=%d smem=%zu] cycles: setup %llu | (a) %llu | (b) own work %llu | (b) barrier wait %llu | (d) %llu | (e) gemm %llu\n",
            NB, n, jres, smem, h[0], h[1], h[2], h[3], h[4], h[5]);
    fprintf(stderr, "   warp 0 units: streamed full %llu x %llu cyc, streamed diag %llu x %llu, resident full %llu x %llu, resident diag %llu x %llu\n",
            h[10], h[10] ? h[6] / h[10] : 0, h[11], h[11] ? h[7] / h[11] : 0, h[12], h[12] ? h[8] / h[12] : 0, h[13], h[13] ? h[9] / h[13] : 0);
#if SQB_PANEL_PROFILE > 1
    unsigned long long h2[4];
    cudaMemcpyFromSymbol(h2, pn_prof2, sizeof(h2));
    fprintf(stderr, "   warp 0 inside (b): setup %llu | units %llu | strip flush %llu | dots %llu\n", h[14], h2[0], h[15], h2[1]);
    memset(h2, 0, sizeof(h2));
    cudaMemcpyToSymbol(pn_prof2, h2, sizeof(h2));
    unsigned long long h3[32];
    cudaMemcpyFromSymbol(h3, pn_prof3, sizeof(h3));
    for (int w = 0; w < kPnWarps; ++w)
      fprintf(stderr, "   warp %d: %llu units, %llu cycles in units, %llu in set-up / flush / dots\n", w, h3[16 + w], h3[w], h3[8 + w]);
    memset(h3, 0, sizeof(h3));
    cudaMemcpyToSymbol(pn_prof3, h3, sizeof(h3));
#endif
    memset(h, 0, sizeof(h));
    cudaMemcpyToSymbol(pn_prof, h, sizeof(h));
  }
#endif
  return SQB_OK;
}

}  // namespace
