// K2: batched complex128 Hermitian eigensolver for the Bloch k-blocks.
//
// Replaces np.linalg.eigh per block (slate_core.linalg.into_diagonal_hermitian, call site
// slate_quantum/operator/_linalg.py:56).  Four kernels per chunk of matrices:
//
//   (1) tridiag_kernel     one CTA per matrix.  Householder reduction H -> real symmetric tridiagonal T
//                          (LAPACK zhetd2 recurrences, lower/column form) carried out in the column-major
//                          view of the row-major input, i.e. on conj(H).  The rank-2 update of step i-1 is
//                          deferred and fused with the matrix-vector product of step i, so every step
//                          streams the LOWER TRIANGLE of the trailing block exactly once (16 B read + 16 B
//                          write per element, Hermitian symmetry supplies the other half of the product);
//                          the block lives in L2.  Reflectors go to workspace Y, so A is dead afterwards.
//   (2) ql_kernel          one warp per matrix.  Implicit-shift QL (EISPACK imtql2 recurrences) on (d, e):
//                          latency-bound scalar chain run by lane 0 for thousands of matrices at once; the
//                          Givens rotations are RECORDED (c, s) instead of applied.  Also ranks eigenvalues.
//   (3) replay_kernel      one warp per (matrix, 32-row strip).  Replays the recorded rotations on a strip of
//                          the identity held in shared memory (rows are independent) -> real eigenvectors
//                          Q_T of T, written with columns in ascending-eigenvalue order.
//   (4) backtransform_kernel  one warp per 2 eigenvectors, vector in registers.  Applies the n-1 Householder
//                          reflectors, conjugates (undoing the conj(H) view) and writes rows = eigenvectors
//                          into A.
//
// Stages (2)+(3) are the fallback (environment SQB_EIGH_TRIDIAG_SOLVER=ql); by default the tridiagonal
// eigenproblem is solved by the divide-and-conquer kernels of eigh_dc.cuh (DMMA merge GEMMs), which write the
// same Q_T layout.  For 64 < n <= 512 stage (4) is the blocked compact-WY back-transformation of eigh_wy.cuh
// (three DMMA GEMMs per block of 32 / 16 reflectors) unless SQB_BACKTRANSFORM=reflector.  The algorithms are modelled step by step in tests/eigh_model.py and tests/dc_model.py.
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <mutex>

#include "sqb_common.cuh"
#include "eigh_dc.cuh"
#include "eigh_wy.cuh"
#include "eigh_panel.cuh"

namespace {

constexpr int kMaxN = 512;
constexpr int kQlMaxIter = 60;

__host__ __device__ inline int64_t rot_capacity(int n) { return (int64_t)5 * n * n / 2 + 64 * (int64_t)n + 64; }
__host__ __device__ inline int sweep_capacity(int n) { return 8 * n + 64; }

struct EighWs {  // per-chunk workspace pointers (each array is [chunk, ...])
  double* d;      // [chunk, n]
  double* e;      // [chunk, n]
  c128* tau;      // [chunk, n]
  int* rank;      // [chunk, n]
  int* nsweeps;   // [chunk]
  int4* sweeps;   // [chunk, sweep_capacity]  {l, m, count, offset}
  double2* rot;   // [chunk, rot_capacity]    {c, s}
  c128* y;        // [chunk, n, n]  reflector i at y[i*n + r]
  double* qt;     // [chunk, n, n]  qt[e*n + k]
  DcWs dc;        // divide-and-conquer arrays (dc.q[0] == qt)
  c128* wy_t;     // [chunk, wy_blocks(n), NB, NB+1] T factors of the blocked back-transformation (padded rows)
};

// blocked (compact WY, DMMA) back-transformation: 64 < n <= 512; SQB_BACKTRANSFORM=reflector disables it
bool eigh_use_wy(int n) {
  static int cached = -1;
  if (cached < 0) {
    const char* v = getenv("SQB_BACKTRANSFORM");
    cached = (v && strcmp(v, "reflector") == 0) ? 0 : 1;
  }
  return cached == 1 && n > 64 && n <= 512;
}

// tridiagonalisation: blocked panel kernel (eigh_panel.cuh) for 64 < n <= 256 unless SQB_TRIDIAG=column;
// SQB_TRIDIAG_NB = 4 | 8 | 16 selects the panel width (default 8)
int eigh_panel_nb(int n) {
  static int cached = -2;
  if (cached == -2) {
    const char* v = getenv("SQB_TRIDIAG");
    const char* nb = getenv("SQB_TRIDIAG_NB");
    cached = (v && strcmp(v, "column") == 0) ? 0 : (nb ? atoi(nb) : 8);
    if (cached != 0 && cached != 4 && cached != 8 && cached != 16) cached = 8;
  }
  return (n > 64 && n <= kPnMaxN) ? cached : 0;
}

// tridiagonal solver: 1 = divide and conquer (default), 0 = implicit QL + rotation replay
int eigh_use_dc() {
  static int cached = -1;
  if (cached < 0) {
    const char* v = getenv("SQB_EIGH_TRIDIAG_SOLVER");
    cached = (v && strcmp(v, "ql") == 0) ? 0 : 1;
  }
  return cached;
}

// ------------------------------------------------------------------------------------------------
// block-wide sum of a double over `nthreads` (multiple of 32) threads; result broadcast to all.
// scratch: >= 33 doubles of shared memory.  Contains two __syncthreads().
__device__ __forceinline__ double block_sum(double v, double* scratch, int nwarps) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) scratch[w] = v;
  __syncthreads();
  if (w == 0) {
    double t = lane < nwarps ? scratch[lane] : 0.0;
    t = warp_sum(t);
    if (lane == 0) scratch[32] = t;
  }
  __syncthreads();
  return scratch[32];
}

// two sums at once (one barrier pair instead of two); scratch: >= 66 doubles
__device__ __forceinline__ double2 block_sum2(double a, double b, double* scratch, int nwarps) {
  a = warp_sum(a);
  b = warp_sum(b);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) {
    scratch[w] = a;
    scratch[32 + w] = b;
  }
  __syncthreads();
  if (w == 0) {
    double ta = lane < nwarps ? scratch[lane] : 0.0;
    double tb = lane < nwarps ? scratch[32 + lane] : 0.0;
    ta = warp_sum(ta);
    tb = warp_sum(tb);
    if (lane == 0) {
      scratch[64] = ta;
      scratch[65] = tb;
    }
  }
  __syncthreads();
  return make_double2(scratch[64], scratch[65]);
}

// ------------------------------------------------------------------------------------------------
// (1) Householder tridiagonalisation.  Threads = 256*RG; warp w: row group rg = w % RG, column group
// cg = w / RG (8 column groups).  Lane rows: r = (c*RG + rg)*32 + lane, c < NR.
//
// Column j is touched in j steps, so the TRAILING columns carry most of the traffic: the columns j >= jres are
// kept resident in shared memory (packed lower triangle, as many as fit next to the work vectors, chosen by
// the host) and only the leading columns j < jres are streamed from global memory during the first jres
// steps.  N <= ~150: everything is resident; N = 256: jres ~ 107, 37 % of the former global traffic.
__device__ __forceinline__ int64_t tri_off(int n, int jres, int j) {
  // offset (in elements) of resident column j's row 0 inside the packed store: column q holds rows q..n-1
  return (int64_t)(j - jres) * n - ((int64_t)j * (j - 1) / 2 - (int64_t)jres * (jres - 1) / 2) - j;
}

template <int NR, int RG>
__global__ void __launch_bounds__(256 * RG, 1)
tridiag_kernel(int n, int jres, c128* __restrict__ A_all, EighWs ws) {
  constexpr int kThreads = 256 * RG;
  constexpr int kWarps = kThreads / 32;
  extern __shared__ double smem_d[];
  c128* vo = reinterpret_cast<c128*>(smem_d);  // pending update vectors (rows >= i)
  c128* wo = vo + n;
  c128* vn = wo + n;                            // new reflector
  c128* red = vn + n;                           // [8][n] partial A22*v sums per column group (rows >= column)
  c128* redc = red + 8 * n;                     // [RG][n] mirrored (upper triangle) contributions per column
  double* scratch = reinterpret_cast<double*>(redc + RG * n);  // [40]
  double* scratch2 = scratch + 40;                              // [72] block_sum2
  c128* res = reinterpret_cast<c128*>(scratch + 112);           // packed lower triangle of columns >= jres
  // scalars: scratch[34..39]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int rg = warp % RG, cg = warp / RG;
  const int64_t b = blockIdx.x;
  c128* A = A_all + b * (int64_t)n * n;
  c128* Y = ws.y + b * (int64_t)n * n;
  double* dd = ws.d + b * n;
  double* ee = ws.e + b * n;
  c128* tt = ws.tau + b * n;

  for (int r = tid; r < n; r += kThreads) {
    vo[r] = make_double2(0.0, 0.0);
    wo[r] = make_double2(0.0, 0.0);
  }
  // load the resident columns (rows j..n-1 of column j), a warp per column, coalesced along rows
  for (int j = jres + warp; j < n; j += kWarps) {
    c128* dst = res + tri_off(n, jres, j);
    const c128* src = A + (int64_t)j * n;
    for (int r = j + lane; r < n; r += 32) dst[r] = src[r];
  }
  __syncthreads();

  for (int i = 0; i < n - 1; ++i) {
    // ---- phase A: materialise column i (rows >= i); one row per thread (kThreads >= n)
    const int r = tid;
    c128 colv = make_double2(0.0, 0.0);
    double part = 0.0;
    if (r >= i && r < n) {
      const c128 cw = cconj(wo[i]), cv = cconj(vo[i]);
      c128 a = i >= jres ? res[tri_off(n, jres, i) + r] : A[(int64_t)i * n + r];
      {
        c128 t = cmul(vo[r], cw);  // the correction is formed independently of `a`: one ADD on its chain
        cfma(t, wo[r], cv);
        a = csub(a, t);
      }
      colv = a;
      if (r >= i + 2) part = a.x * a.x + a.y * a.y;
      if (r == i) dd[i] = a.x;
      if (r == i + 1) {
        scratch[34] = a.x;
        scratch[35] = a.y;
      }
    }
    const double xnorm2 = block_sum(part, scratch, kWarps);  // also publishes scratch[34..35]
    const double ar = scratch[34], ai = scratch[35];
    // every thread computes the same scalars (cheaper than another broadcast round)
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
      // scal = 1 / (alpha - beta)
      const double pr = ar - beta, pi = ai;
      const double den = pr * pr + pi * pi;
      scal = make_double2(pr / den, -pi / den);
    }
    const c128 tau = make_double2(tau_r, tau_i);
    if (r < n) {
      c128 v = make_double2(0.0, 0.0);
      if (r == i + 1) v = make_double2(1.0, 0.0);
      else if (r >= i + 2) v = cmul(colv, scal);
      vn[r] = v;
      if (r >= i + 1) Y[(int64_t)i * n + r] = v;
    }
    if (tid == 0) {
      ee[i] = beta;
      tt[i] = tau;
    }
    __syncthreads();

    // ---- phase B: apply the pending update to the LOWER triangle of the trailing block (rows r >= column j)
    // and accumulate p = A22 * vn using Hermitian symmetry: element a = A(r,j) contributes a*vn[j] to p[r]
    // and, for r > j, conj(a)*vn[r] to p[j].  Only the lower triangle is ever read or written, which halves
    // the L2/HBM traffic of the reduction (it is bandwidth bound).
    c128 acc[NR];
    int rows[NR];
#pragma unroll
    for (int c = 0; c < NR; ++c) {
      acc[c] = make_double2(0.0, 0.0);
      rows[c] = (c * RG + rg) * 32 + lane;
    }
    const int lo = i + 1;
    int j = lo + cg;
    // (B1) streamed columns j < jres: register double buffering over the column loop
    if (j < jres) {
      c128 cur[NR], nxt[NR];
#pragma unroll
      for (int c = 0; c < NR; ++c) {
        cur[c] = make_double2(0.0, 0.0);
        if (rows[c] >= j && rows[c] < n) cur[c] = A[(int64_t)j * n + rows[c]];
      }
      for (; j < jres; j += 8) {
        const int jn = j + 8;
#pragma unroll
        for (int c = 0; c < NR; ++c) {
          nxt[c] = make_double2(0.0, 0.0);
          if (jn < jres && rows[c] >= jn && rows[c] < n) nxt[c] = A[(int64_t)jn * n + rows[c]];
        }
        const c128 cw = cconj(wo[j]), cv = cconj(vo[j]), vj = vn[j];
        c128 dotj = make_double2(0.0, 0.0);
#pragma unroll
        for (int c = 0; c < NR; ++c) {
          if (rows[c] >= j && rows[c] < n) {
            c128 a = cur[c];
            c128 t = cmul(vo[rows[c]], cw);
            cfma(t, wo[rows[c]], cv);
            a = csub(a, t);
            A[(int64_t)j * n + rows[c]] = a;
            cfma(acc[c], a, vj);
            if (rows[c] > j) cfmac(dotj, a, vn[rows[c]]);
          }
        }
        dotj.x = warp_sum(dotj.x);
        dotj.y = warp_sum(dotj.y);
        if (lane == 0) redc[rg * n + j] = dotj;
#pragma unroll
        for (int c = 0; c < NR; ++c) cur[c] = nxt[c];
      }
    }
    // (B2) resident columns: the same update out of shared memory; the lane's rows of vo / wo / vn are hoisted
    // into registers so that an element costs one 16 B shared load and one 16 B shared store
    if (j < n) {
      // NR <= 4: hoisted (48 registers); NR = 8 (n > 256) would spill, it reads the vectors from shared memory
      constexpr bool kHoist = NR <= 4;
      c128 vor[kHoist ? NR : 1], wor[kHoist ? NR : 1], vnr[kHoist ? NR : 1];
      if (kHoist) {
#pragma unroll
        for (int c = 0; c < NR; ++c) {
          const bool in = rows[c] < n;
          vor[kHoist ? c : 0] = in ? vo[rows[c]] : make_double2(0.0, 0.0);
          wor[kHoist ? c : 0] = in ? wo[rows[c]] : make_double2(0.0, 0.0);
          vnr[kHoist ? c : 0] = in ? vn[rows[c]] : make_double2(0.0, 0.0);
        }
      }
      for (; j < n; j += 8) {
        c128* col = res + tri_off(n, jres, j);
        const c128 cw = cconj(wo[j]), cv = cconj(vo[j]), vj = vn[j];
        c128 dotj = make_double2(0.0, 0.0);
#pragma unroll
        for (int c = 0; c < NR; ++c) {
          if (rows[c] >= j && rows[c] < n) {
            c128 a = col[rows[c]];
            c128 t = cmul(kHoist ? vor[kHoist ? c : 0] : vo[rows[c]], cw);
            cfma(t, kHoist ? wor[kHoist ? c : 0] : wo[rows[c]], cv);
            a = csub(a, t);
            col[rows[c]] = a;
            cfma(acc[c], a, vj);
            if (rows[c] > j) cfmac(dotj, a, kHoist ? vnr[kHoist ? c : 0] : vn[rows[c]]);
          }
        }
        dotj.x = warp_sum(dotj.x);
        dotj.y = warp_sum(dotj.y);
        if (lane == 0) redc[rg * n + j] = dotj;
      }
    }
#pragma unroll
    for (int c = 0; c < NR; ++c)
      if (rows[c] < n) red[cg * n + rows[c]] = acc[c];
    __syncthreads();

    // ---- p = tau * A22 vn ; w = p - (tau/2)(p^H vn) vn
    c128 p = make_double2(0.0, 0.0);
    double dr = 0.0, di = 0.0;
    if (r >= i + 1 && r < n) {
      c128 s = red[r];
#pragma unroll
      for (int g = 1; g < 8; ++g) s = cadd(s, red[g * n + r]);
#pragma unroll
      for (int g = 0; g < RG; ++g) s = cadd(s, redc[g * n + r]);
      p = cmul(tau, s);
      const c128 v = vn[r];
      // conj(p) * v
      dr = p.x * v.x + p.y * v.y;
      di = p.x * v.y - p.y * v.x;
    }
    const double2 dot_ri = block_sum2(dr, di, scratch2, kWarps);
    const double dot_r = dot_ri.x, dot_i = dot_ri.y;
    if (r < n) {
      // alpha2 = -0.5 * tau * dot
      const c128 al = cscale(cmul(tau, make_double2(dot_r, dot_i)), -0.5);
      c128 w = make_double2(0.0, 0.0), v = make_double2(0.0, 0.0);
      if (r >= i + 1) {
        v = vn[r];
        w = cadd(p, cmul(al, v));
      }
      vo[r] = v;
      wo[r] = w;
    }
    __syncthreads();
  }
  if (tid == 0) {
    const int l = n - 1;
    c128 a = l >= jres ? res[tri_off(n, jres, l) + l] : A[(int64_t)l * n + l];
    a = csub(a, cmul(vo[l], cconj(wo[l])));
    a = csub(a, cmul(wo[l], cconj(vo[l])));
    dd[l] = a.x;
    ee[l] = 0.0;
    tt[l] = make_double2(0.0, 0.0);
  }
}

// ------------------------------------------------------------------------------------------------
// (2) implicit QL with recorded rotations; one warp per matrix.
__global__ void __launch_bounds__(32)
ql_kernel(int n, int64_t count, EighWs ws, double* __restrict__ w_out, int* __restrict__ info) {
  extern __shared__ double smem_d[];
  double* d = smem_d;            // [n]
  double* e = d + n;             // [n]
  double2* rbuf = reinterpret_cast<double2*>(d + 2 * n);  // [n], 16 B aligned since 2n doubles precede it
  const int lane = threadIdx.x;
  const int64_t b = blockIdx.x;
  if (b >= count) return;
  const double* dg = ws.d + b * n;
  const double* eg = ws.e + b * n;
  for (int i = lane; i < n; i += 32) {
    d[i] = dg[i];
    e[i] = eg[i];
  }
  __syncwarp();
  if (lane == 0) e[n - 1] = 0.0;
  __syncwarp();

  double2* rot = ws.rot + b * rot_capacity(n);
  int4* sweeps = ws.sweeps + b * (int64_t)sweep_capacity(n);
  const int64_t rcap = rot_capacity(n);
  const int scap = sweep_capacity(n);
  int64_t nrot = 0;
  int nsw = 0;
  int fails = 0;
  const double eps = 2.220446049250313e-16;

  for (int l = 0; l < n; ++l) {
    int iter = 0;
    while (true) {
      // warp-parallel search for the first negligible off-diagonal at or after l
      int m = n - 1;
      for (int base = l; base < n - 1; base += 32) {
        const int idx = base + lane;
        bool small = false;
        if (idx < n - 1) {
          const double ddv = fabs(d[idx]) + fabs(d[idx + 1]);
          small = fabs(e[idx]) <= eps * ddv;
        }
        const unsigned mask = __ballot_sync(0xffffffffu, small);
        if (mask) {
          m = base + __ffs(mask) - 1;
          break;
        }
      }
      if (m == l) break;
      if (iter == kQlMaxIter) {
        ++fails;
        break;
      }
      ++iter;
      int cnt = 0;
      if (lane == 0) {
        double g = (d[l + 1] - d[l]) / (2.0 * e[l]);
        double r = hypot(g, 1.0);
        g = d[m] - d[l] + e[l] / (g + copysign(r, g));
        double s = 1.0, c = 1.0, p = 0.0;
        int i = m - 1;
        bool underflow = false;
        for (; i >= l; --i) {
          const double ei = e[i];
          const double f = s * ei;
          const double bb = c * ei;
          r = sqrt(f * f + g * g);
          e[i + 1] = r;
          if (r == 0.0) {
            d[i + 1] -= p;
            e[m] = 0.0;
            underflow = true;
            break;
          }
          const double rinv = 1.0 / r;
          s = f * rinv;
          c = g * rinv;
          g = d[i + 1] - p;
          r = (d[i] - g) * s + 2.0 * c * bb;
          p = s * r;
          d[i + 1] = g + p;
          g = c * r - bb;
          rbuf[cnt++] = make_double2(c, s);
        }
        if (!underflow) {
          d[l] -= p;
          e[l] = g;
          e[m] = 0.0;
        }
      }
      cnt = __shfl_sync(0xffffffffu, cnt, 0);
      __syncwarp();
      // flush this sweep's rotations (coalesced) and its header; every sweep starts at a multiple of 8 in
      // the record so that the replay kernel's 8-rotation blocks never wrap inside its ring
      nrot = (nrot + 7) & ~(int64_t)7;
      if (nsw < scap && nrot + cnt <= rcap) {
        for (int k = lane; k < cnt; k += 32) rot[nrot + k] = rbuf[k];
        if (lane == 0) sweeps[nsw] = make_int4(l, m, cnt, (int)nrot);
        nrot += cnt;
        ++nsw;
      } else {
        fails = n + 1;  // out of rotation storage
      }
      __syncwarp();
    }
  }
  // rank eigenvalues (ascending, stable) and emit sorted
  int* rank = ws.rank + b * n;
  double* wo = w_out + b * n;
  bool bad = false;
  for (int j = lane; j < n; j += 32) {
    const double dj = d[j];
    if (!(dj == dj)) bad = true;
    int rk = 0;
    for (int k = 0; k < n; ++k) {
      const double dk = d[k];
      rk += (dk < dj) || (dk == dj && k < j);
    }
    rank[j] = rk;
    wo[rk] = dj;
  }
  const unsigned anybad = __ballot_sync(0xffffffffu, bad);
  if (lane == 0) {
    ws.nsweeps[b] = nsw;
    info[b] = anybad ? n + 2 : fails;
  }
}

// ------------------------------------------------------------------------------------------------
// (3) rotation replay on a 32-row strip of the identity; one warp per (strip, matrix).
// A lane owns one row; the strip lives in shared memory as zs[col][lane] (conflict free).  Within a QL
// sweep the rotations form a chain with ONE dependent FMA per rotation, so a single sweep is bound by the
// FP64 FMA latency, not by throughput.  The kernel therefore keeps TWO sweeps in flight as a rolling
// pipeline: a leader and, when the next sweep ends at the same column m, a follower that trails the leader
// by >= 16 rotations.  The follower then only reads columns the leader finished in an earlier block, so a
// "dual block" (8 rotations of each) is straight-line code with two independent chains and all its
// shared-memory loads hoisted in front.  The (c, s) record is one contiguous stream; leader and follower
// read it through ONE shared ring (2 cursors, at most n entries apart) refilled 32 entries at a time with
// four refills in flight.
constexpr int kRotBlock = 8;
constexpr int kFollowLag = 16;

struct SweepState {
  int m, cnt, r;       // upper column, rotations, next rotation to apply
  int off;             // first rotation in the record (multiple of 8)
  double carry;        // pending value of column m - r (per lane)
  bool active;
};

__global__ void __launch_bounds__(32)
replay_kernel(int n, int ring_size, EighWs ws) {
  extern __shared__ double smem_d[];
  double* zs = smem_d;                                            // [n][32]  zs[col*32 + lane]
  double2* ring = reinterpret_cast<double2*>(zs + (size_t)n * 32);  // [ring_size], index = record index & mask
  const int mask = ring_size - 1;
  const int lane = threadIdx.x;
  const int r0 = blockIdx.x * 32;
  const int64_t b = blockIdx.y;
  const double2* rot = ws.rot + b * rot_capacity(n);
  const int4* sweeps = ws.sweeps + b * (int64_t)sweep_capacity(n);
  const int nsw = ws.nsweeps[b];
  const int* rank = ws.rank + b * n;
  double* qt = ws.qt + b * (int64_t)n * n;

  for (int col = 0; col < n; ++col) zs[col * 32 + lane] = (col == r0 + lane) ? 1.0 : 0.0;

  int64_t total = 0;
  if (nsw > 0) {
    const int4 last = sweeps[nsw - 1];
    total = (int64_t)last.w + last.z;
  }
  // The ring holds record entries [loaded - ring_size, loaded).  Refills go global -> shared with cp.async
  // (no registers, no waiting moves): 32 entries per commit group, and the most recent kPend groups may
  // still be in flight, i.e. entries below `loaded - 32*kPend` are readable after cp.async.wait_group kPend.
  // The record lives in HBM (far larger than L2), so the prefetch distance has to cover DRAM latency.
  constexpr int kPend = 3;
  const int rcap = (int)rot_capacity(n);
  int loaded = 0;
  auto refill = [&]() {
    const int idx = loaded + lane;
    const uint32_t dst = (uint32_t)__cvta_generic_to_shared(ring + (idx & mask));
    const double2* src = rot + (idx < rcap ? idx : rcap - 1);  // beyond `total` the values are never used
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
    asm volatile("cp.async.commit_group;" ::: "memory");
    loaded += 32;
  };
  for (int q = 0; q < 6; ++q) refill();
  (void)total;

  SweepState L, F;
  L.active = false;
  F.active = false;
  L.m = L.cnt = L.r = 0; L.off = 0; L.carry = 0.0;
  F = L;
  int next_sw = 0;
  int4 hn = nsw > 0 ? sweeps[0] : make_int4(0, 0, 0, 0);  // header of sweep `next_sw`, prefetched

  auto take_next = [&](SweepState& s) {
    s.m = hn.y;
    s.cnt = hn.z;
    s.off = hn.w;
    s.r = 0;
    s.active = true;
    s.carry = zs[s.m * 32 + lane];
    ++next_sw;
    if (next_sw < nsw) hn = sweeps[next_sw];
  };

  while (true) {
    if (!L.active) {
      if (F.active) {
        L = F;
        F.active = false;
      } else if (next_sw < nsw) {
        take_next(L);
      } else {
        break;
      }
    }
    if (!F.active && next_sw < nsw && hn.y == L.m && L.r >= kFollowLag) take_next(F);

    // keep the ring ahead of the furthest cursor (never overwriting the leader's unread entries):
    // `target` is a prefetch goal (the follower's cursor, or the start of the next sweep when there is no
    // follower yet), `need` is what this iteration will certainly read.
    {
      const int lead_pos = L.off + L.r;
      const int target = (F.active ? F.off + F.r : L.off + L.cnt) + kRotBlock + 32 + 32 * kPend;
      int need = lead_pos + kRotBlock;
      if (F.active) need = max(need, F.off + F.r + kRotBlock);
      need += 32 * kPend;
#pragma unroll
      for (int rep = 0; rep < 2; ++rep)
        if (target > loaded) refill();
      while (need > loaded) refill();
      asm volatile("cp.async.wait_group 3;" ::: "memory");
      __syncwarp();
    }

    const int left_l = L.cnt - L.r;
    if (F.active && left_l >= kRotBlock && F.cnt - F.r >= kRotBlock && L.r >= F.r + kFollowLag) {
      // ---- dual block: two independent chains, all loads first
      double2 cl[kRotBlock], cf[kRotBlock];
      double zl[kRotBlock], zf[kRotBlock];
      const double2* rl = ring + ((L.off + L.r) & mask);  // 8-aligned: no wrap inside the block
      const double2* rf = ring + ((F.off + F.r) & mask);
      const int coll = L.m - 1 - L.r, colf = F.m - 1 - F.r;
#pragma unroll
      for (int j = 0; j < kRotBlock; ++j) {
        cl[j] = rl[j];
        cf[j] = rf[j];
        zl[j] = zs[(coll - j) * 32 + lane];
        zf[j] = zs[(colf - j) * 32 + lane];
      }
      double ca = L.carry, cb = F.carry;
#pragma unroll
      for (int j = 0; j < kRotBlock; ++j) {
        const double tl = cl[j].x * zl[j];
        const double tf = cf[j].x * zf[j];
        zs[(coll - j + 1) * 32 + lane] = fma(cl[j].y, zl[j], cl[j].x * ca);
        zs[(colf - j + 1) * 32 + lane] = fma(cf[j].y, zf[j], cf[j].x * cb);
        ca = fma(-cl[j].y, ca, tl);
        cb = fma(-cf[j].y, cb, tf);
      }
      L.carry = ca;
      F.carry = cb;
      L.r += kRotBlock;
      F.r += kRotBlock;
    } else if (left_l >= kRotBlock) {
      // ---- single full block on the leader
      double2 cl[kRotBlock];
      double zl[kRotBlock];
      const double2* rl = ring + ((L.off + L.r) & mask);
      const int coll = L.m - 1 - L.r;
#pragma unroll
      for (int j = 0; j < kRotBlock; ++j) {
        cl[j] = rl[j];
        zl[j] = zs[(coll - j) * 32 + lane];
      }
      double ca = L.carry;
#pragma unroll
      for (int j = 0; j < kRotBlock; ++j) {
        const double tl = cl[j].x * zl[j];
        zs[(coll - j + 1) * 32 + lane] = fma(cl[j].y, zl[j], cl[j].x * ca);
        ca = fma(-cl[j].y, ca, tl);
      }
      L.carry = ca;
      L.r += kRotBlock;
    } else {
      // ---- leader tail (< 8 rotations) and finalisation
      double ca = L.carry;
      for (int k = 0; k < left_l; ++k) {
        const double2 c1 = ring[((L.off + L.r) & mask) + k];
        const int col = L.m - 1 - L.r - k;
        const double z1 = zs[col * 32 + lane];
        zs[(col + 1) * 32 + lane] = fma(c1.y, z1, c1.x * ca);
        ca = fma(-c1.y, ca, c1.x * z1);
      }
      zs[(L.m - L.cnt) * 32 + lane] = ca;
      L.active = false;
    }
  }
  __syncwarp();
  if (r0 + lane < n) {
    for (int col = 0; col < n; ++col) qt[(int64_t)rank[col] * n + r0 + lane] = zs[col * 32 + lane];
  }
}

// ------------------------------------------------------------------------------------------------
// (4) Householder back-transformation.  Warp = 32/LPE eigenvectors; lane kg = lane % LPE owns components
// k = kk*LPE + kg (kk < KPT) of eigenvector e, all in registers (LPE = 16 lanes per eigenvector, 32 for
// n > 256).  The reflectors are streamed through a kBtStages-deep shared-memory ring by 1-D bulk (TMA)
// copies issued by a dedicated producer warp; consumer warps never touch global memory inside the loop.
constexpr int kBtStages = 8;
constexpr int kBtConsumerWarps = 8;

__device__ __forceinline__ uint32_t bt_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void bt_mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bt_smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void bt_mbar_arrive_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bt_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bt_mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bt_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bt_mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(bt_smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void bt_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(bt_smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(bt_smem_u32(bar))
      : "memory");
}

template <int KPT, int LPE>
__global__ void __launch_bounds__(32 * (kBtConsumerWarps + 1), 2)
backtransform_kernel(int n, c128* __restrict__ A_all, EighWs ws) {
  extern __shared__ __align__(16) unsigned char bt_smem[];
  c128* ring = reinterpret_cast<c128*>(bt_smem);                        // [stages][n]
  uint64_t* full = reinterpret_cast<uint64_t*>(ring + kBtStages * n);   // [stages]
  uint64_t* empty = full + kBtStages;                                   // [stages]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t b = blockIdx.y;
  const c128* Y = ws.y + b * (int64_t)n * n;
  const c128* tau = ws.tau + b * n;
  const int n_refl = n - 1;  // reflectors n-2 .. 0, iteration it <-> reflector i = n-2-it

  if (threadIdx.x == 0) {
    for (int s = 0; s < kBtStages; ++s) {
      bt_mbar_init(&full[s], 1);
      bt_mbar_init(&empty[s], kBtConsumerWarps);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  if (warp == kBtConsumerWarps) {
    // ---- producer warp: one lane issues the bulk copies
    if (lane == 0) {
      for (int it = 0; it < n_refl; ++it) {
        const int s = it % kBtStages;
        const uint32_t ph = (uint32_t)((it / kBtStages) & 1);
        bt_mbar_wait(&empty[s], ph ^ 1u);
        const int i = n - 2 - it;
        const uint32_t bytes = (uint32_t)((n - i - 1) * sizeof(c128));
        bt_mbar_arrive_tx(&full[s], bytes);
        bt_bulk_g2s(ring + s * n + i + 1, Y + (int64_t)i * n + i + 1, bytes, &full[s]);
      }
    }
    return;
  }

  const int kg = lane % LPE;
  const int e = (blockIdx.x * kBtConsumerWarps + warp) * (32 / LPE) + (lane / LPE);
  const double* qt = ws.qt + b * (int64_t)n * n;
  c128* out = A_all + b * (int64_t)n * n;
  const bool valid = e < n;

  c128 z[KPT];
#pragma unroll
  for (int kk = 0; kk < KPT; ++kk) {
    const int k = kk * LPE + kg;
    z[kk] = make_double2((valid && k < n) ? qt[(int64_t)e * n + k] : 0.0, 0.0);
  }
  // Group reflectors by q = (i+1)/LPE so that the chunk loop [q, KPT) has compile-time bounds
  // (registers stay statically indexed) while skipping the rows above the reflector's support.
  int it = 0;
#pragma unroll
  for (int q = KPT - 1; q >= 0; --q) {
    const int i_hi = min(n - 2, q * LPE + LPE - 2);  // i+1 <= q*LPE + LPE-1
    const int i_lo = max(0, q * LPE - 1);            // i+1 >= q*LPE
    for (int i = i_hi; i >= i_lo; --i, ++it) {
      const int s = it % kBtStages;
      const uint32_t ph = (uint32_t)((it / kBtStages) & 1);
      const c128 t = tau[i];
      bt_mbar_wait(&full[s], ph);
      if (t.x != 0.0 || t.y != 0.0) {
        const c128* yi = ring + s * n;
        c128 dot = make_double2(0.0, 0.0);
#pragma unroll
        for (int kk = q; kk < KPT; ++kk) {
          const int k = kk * LPE + kg;
          const c128 v = (k >= i + 1 && k < n) ? yi[k] : make_double2(0.0, 0.0);
          cfmac(dot, v, z[kk]);
        }
#pragma unroll
        for (int o = LPE / 2; o > 0; o >>= 1) {
          dot.x += __shfl_xor_sync(0xffffffffu, dot.x, o);
          dot.y += __shfl_xor_sync(0xffffffffu, dot.y, o);
        }
        const c128 sc = cmul(t, dot);
        const c128 ms = make_double2(-sc.x, -sc.y);
        // second sweep re-reads the reflector from shared memory (cheaper than 2*KPT more registers:
        // keeps the kernel at 2 CTAs / SM)
#pragma unroll
        for (int kk = q; kk < KPT; ++kk) {
          const int k = kk * LPE + kg;
          const c128 v = (k >= i + 1 && k < n) ? yi[k] : make_double2(0.0, 0.0);
          cfma(z[kk], v, ms);
        }
      }
      __syncwarp();
      if (lane == 0) bt_mbar_arrive(&empty[s]);
    }
  }
  if (valid) {
#pragma unroll
    for (int kk = 0; kk < KPT; ++kk) {
      const int k = kk * LPE + kg;
      if (k < n) out[(int64_t)e * n + k] = cconj(z[kk]);
    }
  }
}

// ------------------------------------------------------------------------------------------------
size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// ring of the replay kernel: a power of two >= n + 240 record entries (two cursors at most n apart + refills)
int replay_ring_size(int n) {
  int r = 256;
  while (r < n + 240) r *= 2;
  return r;
}

size_t dc_per_matrix_bytes(int n) {
  size_t s = 0;
  s += 256;                                                    // scale (amortised)
  s += align_up((size_t)n * 8, 256) * 3;                       // lam[2], invnorm
  s += align_up((size_t)n * 4, 256) * 3;                       // src_row, out_row, kcount
  s += align_up((size_t)n * n * 8, 256) * 2;                   // q[1], u   (q[0] is qt)
  return s;
}

size_t per_matrix_bytes(int n) {
  size_t s = 0;
  s += align_up((size_t)n * 8, 256) * 2;                       // d, e
  s += align_up((size_t)n * 16, 256);                          // tau
  s += align_up((size_t)n * n * 16, 256);                      // y
  s += align_up((size_t)n * n * 8, 256);                       // qt
  if (eigh_use_wy(n)) s += align_up(wy_t_elems(n) * 16, 256);  // wy_t
  if (eigh_use_dc()) {
    s += dc_per_matrix_bytes(n);
  } else {
    s += align_up((size_t)n * 4, 256);                         // rank
    s += 256;                                                  // nsweeps (amortised)
    s += align_up((size_t)sweep_capacity(n) * 16, 256);        // sweeps
    s += align_up((size_t)rot_capacity(n) * 16, 256);          // rot
  }
  return s;
}

struct Carver {
  char* p;
  char* take(size_t bytes_per, int64_t cnt) {
    char* r = p;
    p += align_up(bytes_per * (size_t)cnt, 256);
    return r;
  }
};

void carve_dc(DcWs& dc, Carver& cv, int n, int64_t chunk, const double* d, const double* e, double* q0) {
  dc.d = d;
  dc.e = e;
  dc.q[0] = q0;
  dc.scale = reinterpret_cast<double*>(cv.take(8, chunk));
  dc.lam[0] = reinterpret_cast<double*>(cv.take((size_t)n * 8, chunk));
  dc.lam[1] = reinterpret_cast<double*>(cv.take((size_t)n * 8, chunk));
  dc.invnorm = reinterpret_cast<double*>(cv.take((size_t)n * 8, chunk));
  dc.src_row = reinterpret_cast<int*>(cv.take((size_t)n * 4, chunk));
  dc.out_row = reinterpret_cast<int*>(cv.take((size_t)n * 4, chunk));
  dc.kcount = reinterpret_cast<int*>(cv.take((size_t)n * 4, chunk));
  dc.q[1] = reinterpret_cast<double*>(cv.take((size_t)n * n * 8, chunk));
  dc.u = reinterpret_cast<double*>(cv.take((size_t)n * n * 8, chunk));
}

void carve(EighWs& ws, void* base, int n, int64_t chunk) {
  Carver cv{reinterpret_cast<char*>(base)};
  ws.d = reinterpret_cast<double*>(cv.take((size_t)n * 8, chunk));
  ws.e = reinterpret_cast<double*>(cv.take((size_t)n * 8, chunk));
  ws.tau = reinterpret_cast<c128*>(cv.take((size_t)n * 16, chunk));
  ws.y = reinterpret_cast<c128*>(cv.take((size_t)n * n * 16, chunk));
  ws.qt = reinterpret_cast<double*>(cv.take((size_t)n * n * 8, chunk));
  ws.wy_t = eigh_use_wy(n) ? reinterpret_cast<c128*>(cv.take(wy_t_elems(n) * 16, chunk)) : nullptr;
  ws.rank = nullptr;
  ws.nsweeps = nullptr;
  ws.sweeps = nullptr;
  ws.rot = nullptr;
  if (eigh_use_dc()) {
    carve_dc(ws.dc, cv, n, chunk, ws.d, ws.e, ws.qt);
  } else {
    ws.rank = reinterpret_cast<int*>(cv.take((size_t)n * 4, chunk));
    ws.nsweeps = reinterpret_cast<int*>(cv.take(4, chunk));
    ws.sweeps = reinterpret_cast<int4*>(cv.take((size_t)sweep_capacity(n) * 16, chunk));
    ws.rot = reinterpret_cast<double2*>(cv.take((size_t)rot_capacity(n) * 16, chunk));
  }
}

template <int NR, int RG>
int launch_tridiag(int n, int64_t count, c128* A, const EighWs& ws, cudaStream_t st) {
  const size_t vec = (size_t)(3 + 8 + RG) * n * sizeof(c128) + 112 * sizeof(double);
  // resident columns: as many trailing columns as fit in the 227 KB opt-in limit next to the vectors; small
  // matrices take only what they need so that several CTAs share an SM
  const size_t budget = 232448 - 1024;
  int jres = n;
  size_t tri = 0;
  while (jres > 0 && vec + (tri + (size_t)(n - (jres - 1))) * sizeof(c128) <= budget) {
    --jres;
    tri += (size_t)(n - jres);
  }
  const size_t smem = vec + tri * sizeof(c128);
  cudaError_t e = cudaFuncSetAttribute(tridiag_kernel<NR, RG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  tridiag_kernel<NR, RG><<<(unsigned)count, 256 * RG, smem, st>>>(n, jres, A, ws);
  SQB_LAUNCH_CHECK();
  return SQB_OK;
}

template <int KPT, int LPE>
int launch_back(int n, int64_t count, c128* A, const EighWs& ws, cudaStream_t st) {
  const int per_cta = kBtConsumerWarps * (32 / LPE);
  dim3 grid((unsigned)((n + per_cta - 1) / per_cta), (unsigned)count);
  const size_t smem = (size_t)kBtStages * n * sizeof(c128) + 2 * kBtStages * sizeof(uint64_t);
  cudaError_t e = cudaFuncSetAttribute(backtransform_kernel<KPT, LPE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  backtransform_kernel<KPT, LPE><<<grid, 32 * (kBtConsumerWarps + 1), smem, st>>>(n, A, ws);
  SQB_LAUNCH_CHECK();
  return SQB_OK;
}

// Two internal streams + fork / join events per device, created lazily ONCE (not per call) and kept for the
// life of the process.  nullptr when any creation fails: the caller then runs single-lane on its own stream.
struct LanePool {
  cudaStream_t lane[2] = {nullptr, nullptr};
  cudaEvent_t fork = nullptr, join[2] = {nullptr, nullptr};
  bool ok = false;
  std::mutex use;  // held while one call enqueues on the lanes (fork / join events are shared)
};

LanePool* lane_pool() {
  static std::mutex mu;
  static LanePool pools[64];
  static bool tried[64] = {false};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
  std::lock_guard<std::mutex> lock(mu);
  LanePool& p = pools[dev];
  if (!tried[dev]) {
    tried[dev] = true;
    bool ok = cudaEventCreateWithFlags(&p.fork, cudaEventDisableTiming) == cudaSuccess;
    for (int l = 0; l < 2 && ok; ++l) {
      ok = cudaStreamCreateWithFlags(&p.lane[l], cudaStreamNonBlocking) == cudaSuccess &&
           cudaEventCreateWithFlags(&p.join[l], cudaEventDisableTiming) == cudaSuccess;
    }
    if (!ok) {
      for (int l = 0; l < 2; ++l) {
        if (p.lane[l]) cudaStreamDestroy(p.lane[l]);
        if (p.join[l]) cudaEventDestroy(p.join[l]);
        p.lane[l] = nullptr;
        p.join[l] = nullptr;
      }
      if (p.fork) cudaEventDestroy(p.fork);
      p.fork = nullptr;
      (void)cudaGetLastError();
    }
    p.ok = ok;
  }
  return p.ok ? &p : nullptr;
}

}  // namespace

extern "C" size_t sqb_eigh_workspace_bytes(int n, int64_t batch) {
  if (n < 1 || n > kMaxN || batch < 0) return 0;
  // recommended: whole batch in one chunk up to 8192 matrices (the QL stage is latency bound, so it wants
  // thousands of matrices in flight); sqb_eigh_batched accepts anything >= one matrix' worth.
  const int64_t chunk = batch < 8192 ? (batch < 1 ? 1 : batch) : 8192;
  return per_matrix_bytes(n) * (size_t)chunk + 4096;
}

extern "C" int sqb_eigh_batched(int n, int64_t batch, sqb_c128* A_inout, double* w, int* info, void* workspace,
                                size_t workspace_bytes, void* stream) {
  if (n < 1 || batch < 0 || !A_inout || !w || !info) return SQB_E_BADARG;
  if (n > kMaxN) return SQB_E_UNSUPPORTED;
  if (batch == 0) return SQB_OK;
  const size_t per = per_matrix_bytes(n);
  if (!workspace || workspace_bytes < per + 4096) return SQB_E_WORKSPACE;
  int64_t chunk = (int64_t)((workspace_bytes - 4096) / per);
  if (chunk > batch) chunk = batch;
  if (chunk > 65535) chunk = 65535;  // grid.y limit of the strip kernels
  cudaStream_t st = sqb_stream(stream);
  c128* A = reinterpret_cast<c128*>(A_inout);
  void* base = reinterpret_cast<void*>(align_up(reinterpret_cast<size_t>(workspace), 256));

  {
    cudaError_t e = cudaFuncSetAttribute(replay_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         n * 32 * 8 + replay_ring_size(n) * 16);
    if (e != cudaSuccess) return (int)e;
  }
  // Two chunks are kept in flight on two internal streams (each with its own half of the workspace): the
  // latency-bound QL stage of one chunk then overlaps the bandwidth / FP64-bound stages of the other.
  bool two_lanes = batch >= 512 && chunk >= 256;
  cudaStream_t lanes[2] = {st, st};
  LanePool* pool = two_lanes ? lane_pool() : nullptr;  // created once per device; nullptr = creation failed
  if (!pool) two_lanes = false;                        // clean single-lane fallback on the caller's stream
  std::unique_lock<std::mutex> pool_lock;
  if (two_lanes) pool_lock = std::unique_lock<std::mutex>(pool->use);
  const int n_lanes = two_lanes ? 2 : 1;
  (void)n_lanes;
  if (two_lanes) chunk = sqb_min64(chunk / 2, (batch + 1) / 2);
  const size_t lane_bytes = align_up(per * (size_t)chunk, 256);
  if (two_lanes) {
    // fork: both lanes start after everything already enqueued on the caller's stream
    cudaError_t e = cudaEventRecord(pool->fork, st);
    for (int l = 0; l < 2 && e == cudaSuccess; ++l) {
      lanes[l] = pool->lane[l];
      e = cudaStreamWaitEvent(lanes[l], pool->fork, 0);
    }
    if (e != cudaSuccess) return (int)e;
  }
  int status = SQB_OK;
  int which = 0;
  for (int64_t b0 = 0; b0 < batch && status == SQB_OK; b0 += chunk, which ^= 1) {
    const int64_t cnt = sqb_min64(chunk, batch - b0);
    const int l = two_lanes ? which : 0;
    cudaStream_t ls = lanes[l];
    EighWs ws;
    carve(ws, reinterpret_cast<char*>(base) + (size_t)l * lane_bytes, n, chunk);
    c128* Ab = A + b0 * (int64_t)n * n;
    int rc;
    const int pnb = eigh_panel_nb(n);
    constexpr int kR1 = 256 / kPnThreads, kR2 = 512 / kPnThreads;  // rows per thread for n <= 256 / n <= 512
    if (pnb == 4 && n <= 256) rc = launch_tridiag_panel<4, kR1>(n, cnt, Ab, ws.d, ws.e, ws.tau, ws.y, ls);
    else if (pnb == 8 && n <= 256) rc = launch_tridiag_panel<8, kR1>(n, cnt, Ab, ws.d, ws.e, ws.tau, ws.y, ls);
    else if (pnb == 16 && n <= 256) rc = launch_tridiag_panel<16, kR1>(n, cnt, Ab, ws.d, ws.e, ws.tau, ws.y, ls);
    else if (pnb == 4) rc = launch_tridiag_panel<4, kR2>(n, cnt, Ab, ws.d, ws.e, ws.tau, ws.y, ls);
    else if (pnb != 0) rc = launch_tridiag_panel<8, kR2>(n, cnt, Ab, ws.d, ws.e, ws.tau, ws.y, ls);
    else if (n <= 32) rc = launch_tridiag<1, 1>(n, cnt, Ab, ws, ls);
    else if (n <= 64) rc = launch_tridiag<2, 1>(n, cnt, Ab, ws, ls);
    else if (n <= 128) rc = launch_tridiag<2, 2>(n, cnt, Ab, ws, ls);
    else if (n <= 256) rc = launch_tridiag<4, 2>(n, cnt, Ab, ws, ls);
    else rc = launch_tridiag<8, 2>(n, cnt, Ab, ws, ls);
    if (rc) { status = rc; break; }

    if (eigh_use_dc()) {
      rc = dc_solve(n, cnt, ws.dc, w + b0 * n, info + b0, ls);
      if (rc) { status = rc; break; }
    } else {
      const size_t ql_smem = (size_t)(2 * n + 2) * 8 + (size_t)n * 16;
      ql_kernel<<<(unsigned)cnt, 32, ql_smem, ls>>>(n, cnt, ws, w + b0 * n, info + b0);
      if (cudaGetLastError() != cudaSuccess) { status = (int)cudaErrorLaunchFailure; break; }

      dim3 rgrid((unsigned)((n + 31) / 32), (unsigned)cnt);
      replay_kernel<<<rgrid, 32, (size_t)n * 32 * 8 + (size_t)replay_ring_size(n) * 16, ls>>>(n, replay_ring_size(n), ws);
      if (cudaGetLastError() != cudaSuccess) { status = (int)cudaErrorLaunchFailure; break; }
    }

    if (eigh_use_wy(n)) {
      // n <= 256: 16-reflector blocks and 8 eigenvectors per CTA keep the CTA below half an SM (two CTAs overlap
      // each other's staging / reduction phases: 132.4 vs 135.0 ms at cfg3); n <= 512: 16 eigenvectors per CTA
      static const int wy_variant = getenv("SQB_WY_VARIANT") ? atoi(getenv("SQB_WY_VARIANT")) : 1;
      if (n <= 128 && wy_variant == 1) rc = launch_back_wy<16, 2, 1>(n, cnt, Ab, ws.qt, ws.y, ws.tau, ws.wy_t, ls);
      else if (n <= 128) rc = launch_back_wy<32, 2, 2>(n, cnt, Ab, ws.qt, ws.y, ws.tau, ws.wy_t, ls);
      else if (n <= 256 && wy_variant == 1) rc = launch_back_wy<16, 4, 1>(n, cnt, Ab, ws.qt, ws.y, ws.tau, ws.wy_t, ls);
      else if (n <= 256) rc = launch_back_wy<32, 4, 2>(n, cnt, Ab, ws.qt, ws.y, ws.tau, ws.wy_t, ls);
      else rc = launch_back_wy<16, 8, 2>(n, cnt, Ab, ws.qt, ws.y, ws.tau, ws.wy_t, ls);
    } else if (n <= 64) rc = launch_back<4, 16>(n, cnt, Ab, ws, ls);
    else if (n <= 128) rc = launch_back<8, 16>(n, cnt, Ab, ws, ls);
    else if (n <= 256) rc = launch_back<16, 16>(n, cnt, Ab, ws, ls);
    else if (n <= 384) rc = launch_back<12, 32>(n, cnt, Ab, ws, ls);
    else rc = launch_back<16, 32>(n, cnt, Ab, ws, ls);
    if (rc) { status = rc; break; }
  }
  if (two_lanes) {
    // join: the caller's stream continues only after both lanes have drained (also on an error path, so that
    // the caller's workspace cannot be reused while a lane still runs)
    for (int l = 0; l < 2; ++l) {
      cudaError_t e = cudaEventRecord(pool->join[l], lanes[l]);
      if (e == cudaSuccess) e = cudaStreamWaitEvent(st, pool->join[l], 0);
      if (e != cudaSuccess) {
        cudaStreamSynchronize(lanes[l]);  // last resort: order by blocking the host
        if (status == SQB_OK) status = (int)e;
      }
    }
  }
  return status;
}

// real symmetric tridiagonal eigenproblems on their own (the divide-and-conquer stage of sqb_eigh_batched)
extern "C" size_t sqb_tridiag_eigh_workspace_bytes(int n, int64_t batch) {
  if (n < 1 || n > kMaxN || batch < 0) return 0;
  return dc_per_matrix_bytes(n) * (size_t)(batch < 1 ? 1 : batch) + 4096;
}

extern "C" int sqb_tridiag_eigh_batched(int n, int64_t batch, const double* d, const double* e, double* w,
                                        double* q_rows, int* info, void* workspace, size_t workspace_bytes,
                                        void* stream) {
  if (n < 1 || batch < 0 || !d || !e || !w || !q_rows || !info) return SQB_E_BADARG;
  if (n > kMaxN) return SQB_E_UNSUPPORTED;
  if (batch == 0) return SQB_OK;
  if (batch > 65535) return SQB_E_UNSUPPORTED;  // grid.y / grid.z limit of the merge kernels
  if (!workspace || workspace_bytes < sqb_tridiag_eigh_workspace_bytes(n, batch)) return SQB_E_WORKSPACE;
  Carver cv{reinterpret_cast<char*>(align_up(reinterpret_cast<size_t>(workspace), 256))};
  DcWs dc;
  carve_dc(dc, cv, n, batch, d, e, q_rows);
  return dc_solve(n, batch, dc, w, info, sqb_stream(stream));
}
