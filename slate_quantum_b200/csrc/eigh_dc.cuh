// K2, stage 2: batched divide-and-conquer eigensolver for the real symmetric tridiagonal matrices produced by
// tridiag_kernel (Cuppen's method with the deflation rules of LAPACK dlaed2, a dlaed4-style secular solver and
// the Gu / Eisenstat re-computation of z).  Included by eigh.cu (one translation unit).
//
// Why: eigenvectors of T by implicit QL cost ~6 N^3 flop of Givens rotations per matrix, applied as chains
// that are bound by FP64 FMA latency (replay_kernel: 95 ms for 4096 x N=256).  Divide and conquer turns the
// eigenvector work into dense products  Qe_new = U^T . Qe  (<= (8/3) N^3 flop, usually far less after
// deflation) that run on the FP64 tensor cores (DMMA.8x8x4), and the O(N^2) secular / deflation part is
// parallel over roots.  The numpy model of this file is tests/dc_model.py.
//
// Tree: all leaves at depth L = min{ L : ceil(n / 2^L) <= 32 }, node i of level l covers
// [floor(i n / 2^l), floor((i+1) n / 2^l)).  The tears of every level are applied to the diagonal up front
// (d[s-1] -= |e[s-1]|, d[s] -= |e[s-1]| at every node boundary s), so all leaves start at once.
//
// Eigenvectors are stored as ROWS (Qe[e][r], block diagonal inside an n x n array per matrix), ping-ponging
// between two buffers per level; every stage emits eigenvalues in ascending order with its rows permuted
// accordingly, so the last merge writes the final sorted Q_T straight into the back-transform's input.
#pragma once

namespace {

constexpr int kDcLeaf = 32;
constexpr int kDcMaxSecularIter = 40;
constexpr int kDcQlMaxIter = 60;
constexpr double kDcEps = 2.220446049250313e-16;

__host__ __device__ inline int dc_depth(int n) {
  int depth = 0;
  while (((n + (1 << depth) - 1) >> depth) > kDcLeaf) ++depth;
  return depth;
}
__host__ __device__ inline int dc_start(int n, int level, int i) { return (int)(((int64_t)i * n) >> level); }

struct DcWs {             // per-chunk arrays
  const double* d;        // [chunk, n]   tridiagonal diagonal (unscaled input)
  const double* e;        // [chunk, n]   off-diagonal, e[i] couples i and i+1
  double* scale;          // [chunk]      max-norm of (d, e); the solve runs on T / scale
  double* lam[2];         // [chunk, n]   eigenvalues of the nodes of the current / next level
  double* q[2];           // [chunk, n, n] eigenvector rows; q[0] receives the final result
  double* u;              // [chunk, n, n] secular eigenvector matrices U[pole i][root j] (un-normalised)
  double* invnorm;        // [chunk, n]   1 / |U[:, j]|
  int* src_row;           // [chunk, n]   K index of the merge GEMM -> row of the old buffer (relative to node)
  int* out_row;           // [chunk, n]   M index -> row of the new buffer (ascending eigenvalue rank)
  int* kcount;            // [chunk, n]   at the node's start: number of non-deflated poles
};

// ------------------------------------------------------------------------------------------------
// leaves: one warp per (matrix, leaf); implicit QL (EISPACK imtql2 recurrences) with the rotations applied at
// once to the leaf's eigenvector matrix, lane = row, held in shared memory.  The scalar recurrence is
// evaluated redundantly by all lanes (same cost as one lane, no broadcasts).
constexpr int kDcLeafWarps = 4;

__global__ void __launch_bounds__(32 * kDcLeafWarps)
dc_leaf_kernel(int n, int depth, int64_t count, DcWs ws, int buf, double* __restrict__ w_out, int* __restrict__ info) {
  __shared__ double zs_all[kDcLeafWarps][32 * 32];
  __shared__ double d_all[kDcLeafWarps][32];
  __shared__ double e_all[kDcLeafWarps][32];
  __shared__ int rk_all[kDcLeafWarps][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nleaf = 1 << depth;
  const int64_t gw = (int64_t)blockIdx.x * kDcLeafWarps + warp;
  if (gw >= count * nleaf) return;
  const int64_t b = gw / nleaf;
  const int leaf = (int)(gw - b * nleaf);
  const int s0 = dc_start(n, depth, leaf), s1 = dc_start(n, depth, leaf + 1);
  const int m = s1 - s0;
  double* zs = zs_all[warp];
  double* d = d_all[warp];
  double* e = e_all[warp];
  int* rk = rk_all[warp];
  const double* dg = ws.d + b * n;
  const double* eg = ws.e + b * n;

  double mx = 0.0;
  for (int i = lane; i < n; i += 32) {
    mx = fmax(mx, fabs(dg[i]));
    if (i < n - 1) mx = fmax(mx, fabs(eg[i]));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  const double scale = mx > 0.0 ? mx : 1.0;
  const double inv = 1.0 / scale;
  if (leaf == 0 && lane == 0) ws.scale[b] = scale;

  if (lane < m) {
    double dv = dg[s0 + lane] * inv;
    if (lane == 0 && s0 > 0) dv -= fabs(eg[s0 - 1] * inv);
    if (lane == m - 1 && s1 < n) dv -= fabs(eg[s1 - 1] * inv);
    d[lane] = dv;
    e[lane] = lane < m - 1 ? eg[s0 + lane] * inv : 0.0;
  }
  for (int col = 0; col < m; ++col) zs[col * 32 + lane] = (col == lane) ? 1.0 : 0.0;
  __syncwarp();

  int fails = 0;
  for (int l = 0; l < m; ++l) {
    int iter = 0;
    while (true) {
      bool small = false;
      const int idx = l + lane;
      if (idx < m - 1) small = fabs(e[idx]) <= kDcEps * (fabs(d[idx]) + fabs(d[idx + 1]));
      const unsigned mask = __ballot_sync(0xffffffffu, small);
      const int mm = mask ? l + __ffs(mask) - 1 : m - 1;
      if (mm == l) break;
      if (iter == kDcQlMaxIter) {
        ++fails;
        break;
      }
      ++iter;
      double g = (d[l + 1] - d[l]) / (2.0 * e[l]);
      double r = hypot(g, 1.0);
      g = d[mm] - d[l] + e[l] / (g + copysign(r, g));
      double s = 1.0, c = 1.0, p = 0.0;
      double carry = zs[mm * 32 + lane];
      bool underflow = false;
      int i = mm - 1;
      for (; i >= l; --i) {
        const double ei = e[i];
        const double f = s * ei;
        const double bb = c * ei;
        r = sqrt(f * f + g * g);
        __syncwarp();
        e[i + 1] = r;
        if (r == 0.0) {
          d[i + 1] -= p;
          e[mm] = 0.0;
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
        const double zi = zs[i * 32 + lane];
        zs[(i + 1) * 32 + lane] = fma(s, zi, c * carry);
        carry = fma(-s, carry, c * zi);
      }
      zs[(i + 1) * 32 + lane] = carry;  // column l after a full sweep, column i+1 after an underflow exit
      if (!underflow) {
        d[l] -= p;
        e[l] = g;
        e[mm] = 0.0;
      }
      __syncwarp();
    }
  }
  __syncwarp();
  bool bad = false;
  if (lane < m) {
    const double dj = d[lane];
    bad = !(dj == dj);
    int rank = 0;
    for (int k = 0; k < m; ++k) {
      const double dk = d[k];
      rank += (dk < dj) || (dk == dj && k < lane);
    }
    rk[lane] = rank;
    if (depth == 0) w_out[b * n + s0 + rank] = dj * scale;
    else ws.lam[buf][b * n + s0 + rank] = dj;
  }
  __syncwarp();
  const unsigned anybad = __ballot_sync(0xffffffffu, bad);
  if (lane == 0 && (anybad || fails)) atomicMax(info + b, anybad ? n + 2 : fails);
  double* q = ws.q[buf] + b * (int64_t)n * n;
  if (lane < m)
    for (int col = 0; col < m; ++col) q[(int64_t)(s0 + rk[col]) * n + s0 + lane] = zs[col * 32 + lane];
}

// ------------------------------------------------------------------------------------------------
// secular equation  1/rho + sum_i z2_i / (dk_i - lambda) = 0, root j in (dk_j, dk_{j+1}) (last root: right of
// dk_{k-1}).  Works in tau = lambda - dk[origin] with the origin at the closer pole, so that every difference
// dk_i - lambda = (dk_i - dk[origin]) - tau is accurate.  Start: the two neighbouring poles exact, the rest
// frozen at the interval midpoint (dlaed4); iteration: two-pole rational interpolation ("middle way") with a
// Newton fallback and a shrinking bracket.  tests/dc_model.py::secular_root is the line-by-line model.
__device__ void dc_secular_root(int j, int k, const double* __restrict__ dk, const double* __restrict__ z2, double rho,
                                int* origin_out, double* tau_out) {
  if (k == 1) {
    *origin_out = 0;
    *tau_out = rho * z2[0];
    return;
  }
  const double rhoinv = 1.0 / rho;
  int origin, jlo, jhi;
  bool last;
  double lo, hi, tau;
  if (j < k - 1) {
    last = false;
    jlo = j;
    jhi = j + 1;
    const double dj = dk[j];
    const double gap = dk[j + 1] - dj;
    const double half = 0.5 * gap;
    double c = rhoinv;
    for (int i = 0; i < j; ++i) c += z2[i] / ((dk[i] - dj) - half);
    for (int i = j + 2; i < k; ++i) c += z2[i] / ((dk[i] - dj) - half);
    const double w = c + z2[j] / (-half) + z2[j + 1] / (gap - half);
    if (w > 0.0) {
      origin = j;
      lo = 0.0;
      hi = half;
      const double a = c * gap + z2[j] + z2[j + 1];
      const double bq = z2[j] * gap;
      const double den = a + sqrt(fabs(a * a - 4.0 * bq * c));
      tau = den > 0.0 ? 2.0 * bq / den : hi;
    } else {
      origin = j + 1;
      lo = -half;
      hi = 0.0;
      const double a = -c * gap + z2[j] + z2[j + 1];
      const double bq = z2[j + 1] * gap;
      const double den = a + sqrt(fabs(a * a + 4.0 * bq * c));
      tau = den > 0.0 ? -2.0 * bq / den : lo;
    }
  } else {
    last = true;
    jlo = k - 2;
    jhi = k - 1;
    origin = k - 1;
    const double dorig = dk[k - 1];
    double zsum = 0.0;
    for (int i = 0; i < k; ++i) zsum += z2[i];
    lo = 0.0;
    hi = rho * zsum * (1.0 + 8.0 * kDcEps) + 2.2250738585072014e-308;
    const double mid = 0.5 * hi;
    double c = rhoinv;
    for (int i = 0; i < k - 2; ++i) c += z2[i] / ((dk[i] - dorig) - mid);
    const double gap = dorig - dk[k - 2];
    const double w = c + z2[k - 2] / (-gap - mid) + z2[k - 1] / (-mid);
    const double a = -c * gap + z2[k - 2] + z2[k - 1];
    const double bq = z2[k - 1] * gap;
    if (c > 0.0) {
      const double disc = sqrt(fabs(a * a + 4.0 * bq * c));
      tau = a < 0.0 ? 2.0 * bq / (disc - a) : (a + disc) / (2.0 * c);
    } else {
      tau = hi;
    }
    if (w > 0.0) hi = mid;
    else lo = mid;
  }
  if (!(lo < tau && tau < hi)) tau = 0.5 * (lo + hi);
  const double dorig = dk[origin];
  const int split = jhi;  // terms [0, split) form psi, [split, k) phi (interior: split = j+1; last: k-1)
  for (int it = 0; it < kDcMaxSecularIter; ++it) {
    double psi = 0.0, dpsi = 0.0, phi = 0.0, dphi = 0.0;
    for (int i = 0; i < split; ++i) {
      const double rinv = 1.0 / ((dk[i] - dorig) - tau);
      const double t = z2[i] * rinv;
      psi += t;
      dpsi = fma(t, rinv, dpsi);
    }
    for (int i = split; i < k; ++i) {
      const double rinv = 1.0 / ((dk[i] - dorig) - tau);
      const double t = z2[i] * rinv;
      phi += t;
      dphi = fma(t, rinv, dphi);
    }
    const double w = rhoinv + psi + phi;
    const double erretm = 8.0 * (fabs(phi) + fabs(psi)) + 2.0 * rhoinv + fabs(tau) * (dpsi + dphi);
    if (fabs(w) <= kDcEps * erretm) break;
    if (w > 0.0) hi = tau;
    else lo = tau;
    const double dl = (dk[jlo] - dorig) - tau, du = (dk[jhi] - dorig) - tau;
    double c = w - dl * dpsi - du * dphi;
    const double a = (dl + du) * w - dl * du * (dpsi + dphi);
    const double bq = dl * du * w;
    double eta;
    if (last) {
      if (c < 0.0) c = fabs(c);
      if (c == 0.0) {
        eta = hi - tau;
      } else {
        const double disc = sqrt(fabs(a * a - 4.0 * bq * c));
        eta = a >= 0.0 ? (a + disc) / (2.0 * c) : 2.0 * bq / (a - disc);
      }
    } else {
      if (c == 0.0) {
        eta = a != 0.0 ? bq / a : 0.0;
      } else {
        const double disc = sqrt(fabs(a * a - 4.0 * bq * c));
        eta = a <= 0.0 ? (a - disc) / (2.0 * c) : 2.0 * bq / (a + disc);
      }
    }
    if (w * eta >= 0.0) eta = -w / (dpsi + dphi);
    double nw = tau + eta;
    if (!(lo < nw && nw < hi)) nw = 0.5 * (lo + hi);
    if (nw == tau || nw == lo || nw == hi) break;
    tau = nw;
  }
  *origin_out = origin;
  *tau_out = tau;
}

// ------------------------------------------------------------------------------------------------
// merge set-up: one CTA per (node, matrix).  Sort, deflate, solve the secular equation, rebuild z, write the
// (un-normalised) secular eigenvector matrix U and the row maps for dc_merge_gemm_kernel.
constexpr int kDcSetupThreads = 256;

__host__ __device__ inline size_t dc_setup_smem_bytes(int nmax) {
  return (size_t)nmax * (8 * sizeof(double) + 3 * sizeof(int) + sizeof(int2) + sizeof(double2)) + 64 * sizeof(double);
}

__global__ void __launch_bounds__(kDcSetupThreads)
dc_merge_setup_kernel(int n, int level, int nmax, DcWs ws, int src, double* __restrict__ w_out, int* __restrict__ info) {
  extern __shared__ __align__(16) unsigned char dc_smem[];
  double* d = reinterpret_cast<double*>(dc_smem);  // [nmax] children's eigenvalues (modified by deflation)
  double* z = d + nmax;                            // [nmax]
  double* dk = z + nmax;                           // [nmax] non-deflated poles, ascending
  double* zk = dk + nmax;
  double* z2 = zk + nmax;
  double* mu = z2 + nmax;                          // root j = dk[origin[j]] + mu[j]
  double* zhat = mu + nmax;
  double* lamn = zhat + nmax;                      // merged eigenvalues: k roots, then the deflated ones
  double* scratch = lamn + nmax;                   // [64]
  double2* gcs = reinterpret_cast<double2*>(scratch + 64);  // [nmax] deflation rotations (c, s)
  int2* gp = reinterpret_cast<int2*>(gcs + nmax);           // [nmax] rows (pj, j)
  int* order = reinterpret_cast<int*>(gp + nmax);           // [nmax] ascending order of d
  int* ndl = order + nmax;                                  // [nmax] non-deflated (front) / deflated (back) indices
  int* origin = ndl + nmax;                                 // [nmax]
  int* iscr = reinterpret_cast<int*>(scratch + 48);         // k, number of rotations

  const int tid = threadIdx.x;
  const int node = blockIdx.x;
  const int64_t b = blockIdx.y;
  const int s0 = dc_start(n, level, node), s1 = dc_start(n, level, node + 1);
  const int sm = dc_start(n, level + 1, 2 * node + 1);
  const int nn = s1 - s0, n1 = sm - s0;
  const double inv = 1.0 / ws.scale[b];
  const double beta = ws.e[b * n + sm - 1] * inv;
  const double sgn = beta >= 0.0 ? 1.0 : -1.0;
  const double rho = 2.0 * fabs(beta);
  const double* lam_in = ws.lam[src] + b * n + s0;
  double* qold = ws.q[src] + b * (int64_t)n * n;

  double dmax = 0.0, zmax = 0.0;
  for (int i = tid; i < nn; i += kDcSetupThreads) {
    const double di = lam_in[i];
    const double zi = (i < n1 ? qold[(int64_t)(s0 + i) * n + sm - 1] : sgn * qold[(int64_t)(s0 + i) * n + sm]) *
                      0.70710678118654752440;
    d[i] = di;
    z[i] = zi;
    order[i] = i;  // keeps every slot a valid index even if NaNs break the merge below
    dmax = fmax(dmax, fabs(di));
    zmax = fmax(zmax, fabs(zi));
  }
  // block max of max(|d|, |z|)
  {
    double v = fmax(dmax, zmax);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    if ((tid & 31) == 0) scratch[tid >> 5] = v;
  }
  __syncthreads();
  // merged ascending order of the two sorted halves (ties: left half first)
  for (int i = tid; i < nn; i += kDcSetupThreads) {
    const double di = d[i];
    int cnt = 0;
    if (i < n1) {
      for (int x = n1; x < nn; ++x) cnt += d[x] < di;
      order[i + cnt] = i;
    } else {
      for (int x = 0; x < n1; ++x) cnt += d[x] <= di;
      order[i - n1 + cnt] = i;
    }
  }
  __syncthreads();
  if (tid == 0) {
    double mxv = 0.0;
    for (int w = 0; w < kDcSetupThreads / 32; ++w) mxv = fmax(mxv, scratch[w]);
    const double tol = 8.0 * kDcEps * mxv;
    // sequential deflation scan (dlaed2): negligible z, then close pole pairs by a Givens rotation
    int k = 0, ndefl = 0, ng = 0, pj = -1;
    for (int t = 0; t < nn; ++t) {
      const int j = order[t];
      if (rho * fabs(z[j]) <= tol) {
        ndl[nn - 1 - ndefl++] = j;
        continue;
      }
      if (pj < 0) {
        pj = j;
        continue;
      }
      double s = z[pj], c = z[j];
      const double tau = hypot(c, s);
      const double tt = d[j] - d[pj];
      c /= tau;
      s = -s / tau;
      if (fabs(tt * c * s) <= tol) {
        z[j] = tau;
        z[pj] = 0.0;
        gp[ng] = make_int2(pj, j);
        gcs[ng] = make_double2(c, s);
        ++ng;
        const double t2 = d[pj] * c * c + d[j] * s * s;
        d[j] = d[pj] * s * s + d[j] * c * c;
        d[pj] = t2;
        ndl[nn - 1 - ndefl++] = pj;
        pj = j;
      } else {
        ndl[k++] = pj;
        pj = j;
      }
    }
    if (pj >= 0) ndl[k++] = pj;
    iscr[0] = k;
    iscr[1] = ng;
  }
  __syncthreads();
  const int k = iscr[0], ng = iscr[1];
  // deflation rotations on the eigenvector rows (a thread owns the same columns in every rotation: no syncs)
  for (int g = 0; g < ng; ++g) {
    const int2 pr = gp[g];
    const double2 cs = gcs[g];
    double* rp = qold + (int64_t)(s0 + pr.x) * n + s0;
    double* rj = qold + (int64_t)(s0 + pr.y) * n + s0;
    for (int r = tid; r < nn; r += kDcSetupThreads) {
      const double a = rp[r], bq = rj[r];
      rp[r] = cs.x * a + cs.y * bq;
      rj[r] = -cs.y * a + cs.x * bq;
    }
  }
  for (int i = tid; i < k; i += kDcSetupThreads) {
    const int src_i = ndl[i];
    const double zi = z[src_i];
    dk[i] = d[src_i];
    zk[i] = zi;
    z2[i] = zi * zi;
  }
  __syncthreads();
  for (int j = tid; j < k; j += kDcSetupThreads) dc_secular_root(j, k, dk, z2, rho, &origin[j], &mu[j]);
  __syncthreads();
  // Gu / Eisenstat: zhat_i^2 ~ prod_j (dk_i - lam_j) / prod_{j != i} (dk_i - dk_j), every difference accurate
  for (int i = tid; i < k; i += kDcSetupThreads) {
    const double di = dk[i];
    double w = (di - dk[origin[i]]) - mu[i];
    for (int j = 0; j < k; ++j) {
      if (j == i) continue;
      w *= ((di - dk[origin[j]]) - mu[j]) / (di - dk[j]);
    }
    zhat[i] = copysign(sqrt(fabs(w)), zk[i]);
  }
  __syncthreads();
  double* u = ws.u + b * (int64_t)n * n;
  for (int j = tid; j < k; j += kDcSetupThreads) {
    const double doj = dk[origin[j]], muj = mu[j];
    double nrm = 0.0;
    for (int i = 0; i < k; ++i) {
      const double v = zhat[i] / ((dk[i] - doj) - muj);
      nrm = fma(v, v, nrm);
      u[(int64_t)(s0 + i) * n + s0 + j] = v;
    }
    ws.invnorm[b * n + s0 + j] = 1.0 / sqrt(nrm);
    lamn[j] = doj + muj;
  }
  for (int t = k + tid; t < nn; t += kDcSetupThreads) lamn[t] = d[ndl[t]];
  __syncthreads();
  bool bad = false;
  const double out_scale = level == 0 ? ws.scale[b] : 1.0;
  double* lam_out = level == 0 ? w_out + b * n : ws.lam[src ^ 1] + b * n + s0;
  for (int j = tid; j < nn; j += kDcSetupThreads) {
    const double lj = lamn[j];
    bad |= !(lj == lj);
    int rank = 0;
    for (int x = 0; x < nn; ++x) {
      const double lx = lamn[x];
      rank += (lx < lj) || (lx == lj && x < j);
    }
    ws.out_row[b * n + s0 + j] = rank;
    ws.src_row[b * n + s0 + j] = ndl[j];
    lam_out[rank] = lj * out_scale;
  }
  if (bad) atomicMax(info + b, n + 2);
  if (tid == 0) ws.kcount[b * n + s0] = k;
}

// ------------------------------------------------------------------------------------------------
// merge GEMM on the FP64 tensor cores: for node rows j < k
//   Qnew[out_row[j]][r] = invnorm[j] * sum_{i < k} U[i][j] * Qold[src_row[i]][r]
// and a plain row copy for the deflated rows j >= k.  CTA tile 64 (j) x 64 (r), 4 warps of 32 x 32, K chunks
// of 16 staged through shared memory with a register prefetch of the next chunk.
constexpr int kDgTile = 64;
constexpr int kDgKc = 16;
constexpr int kDgLd = kDgTile + 8;  // == 8 (mod 32): conflict-free fragment loads As[k][m], Bs[k][n]
constexpr int kDgThreads = 128;

__device__ __forceinline__ void dc_dmma(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(d0), "+d"(d1)
               : "d"(a), "d"(b));
}

__global__ void __launch_bounds__(kDgThreads)
dc_merge_gemm_kernel(int n, int level, DcWs ws, int src) {
  __shared__ double As[2][kDgKc * kDgLd];
  __shared__ double Bs[2][kDgKc * kDgLd];
  __shared__ int srow[kDgTile * 8];  // src_row of the K range (n <= 512)
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int node = blockIdx.y;
  const int64_t b = blockIdx.z;
  const int s0 = dc_start(n, level, node), s1 = dc_start(n, level, node + 1);
  const int nn = s1 - s0;
  const int tiles = (nn + kDgTile - 1) / kDgTile;
  if ((int)blockIdx.x >= tiles * tiles) return;
  const int j0 = ((int)blockIdx.x / tiles) * kDgTile;
  const int r0 = ((int)blockIdx.x % tiles) * kDgTile;
  const int k = ws.kcount[b * n + s0];
  const int* src_row = ws.src_row + b * n + s0;
  const int* out_row = ws.out_row + b * n + s0;
  const double* qold = ws.q[src] + b * (int64_t)n * n;
  double* qnew = ws.q[src ^ 1] + b * (int64_t)n * n;
  const double* u = ws.u + b * (int64_t)n * n;

  if (j0 < k) {
    for (int i = tid; i < k; i += kDgThreads) srow[i] = src_row[i];
    __syncthreads();
    const int wm = (warp & 1) * 32, wn = (warp >> 1) * 32;
    const int frow = lane >> 2, fcol = lane & 3;
    double acc[4][4][2];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int c = 0; c < 4; ++c) acc[a][c][0] = acc[a][c][1] = 0.0;
    const int nchunks = (k + kDgKc - 1) / kDgKc;
    // staging: element idx = tid + 128 q, q < 8: kk = idx / 64, col = idx % 64
    double ra[8], rb[8];
    auto fetch = [&](int kc) {
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const int idx = tid + kDgThreads * q;
        const int kk = kc * kDgKc + (idx >> 6), col = idx & 63;
        double va = 0.0, vb = 0.0;
        if (kk < k) {
          if (j0 + col < k) va = u[(int64_t)(s0 + kk) * n + s0 + j0 + col];
          if (r0 + col < nn) vb = qold[(int64_t)(s0 + srow[kk]) * n + s0 + r0 + col];
        }
        ra[q] = va;
        rb[q] = vb;
      }
    };
    auto stash = [&](int s) {
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const int idx = tid + kDgThreads * q;
        As[s][(idx >> 6) * kDgLd + (idx & 63)] = ra[q];
        Bs[s][(idx >> 6) * kDgLd + (idx & 63)] = rb[q];
      }
    };
    fetch(0);
    stash(0);
    __syncthreads();
    for (int kc = 0; kc < nchunks; ++kc) {
      const int s = kc & 1;
      if (kc + 1 < nchunks) fetch(kc + 1);
#pragma unroll
      for (int k4 = 0; k4 < kDgKc; k4 += 4) {
        double af[4], bf[4];
#pragma unroll
        for (int a = 0; a < 4; ++a) af[a] = As[s][(k4 + fcol) * kDgLd + wm + a * 8 + frow];
#pragma unroll
        for (int c = 0; c < 4; ++c) bf[c] = Bs[s][(k4 + fcol) * kDgLd + wn + c * 8 + frow];
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
          for (int c = 0; c < 4; ++c) dc_dmma(acc[a][c][0], acc[a][c][1], af[a], bf[c]);
      }
      if (kc + 1 < nchunks) stash(s ^ 1);
      __syncthreads();
    }
    const double* invnorm = ws.invnorm + b * n + s0;
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      const int j = j0 + wm + a * 8 + frow;
      if (j >= k) continue;
      const double sc = invnorm[j];
      double* orow = qnew + (int64_t)(s0 + out_row[j]) * n + s0;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int r = r0 + wn + c * 8 + fcol * 2;
        if (r < nn) orow[r] = acc[a][c][0] * sc;
        if (r + 1 < nn) orow[r + 1] = acc[a][c][1] * sc;
      }
    }
  }
  // deflated rows of this tile: unchanged eigenvectors, moved to their sorted position
  const int jd0 = max(j0, k), jd1 = min(j0 + kDgTile, nn);
  for (int idx = tid; idx < (jd1 - jd0) * kDgTile; idx += kDgThreads) {
    const int j = jd0 + (idx >> 6), r = r0 + (idx & 63);
    if (r < nn) qnew[(int64_t)(s0 + out_row[j]) * n + s0 + r] = qold[(int64_t)(s0 + src_row[j]) * n + s0 + r];
  }
}

// ------------------------------------------------------------------------------------------------
// host driver: eigen-decomposition of `count` tridiagonals; eigenvalues ascending into w_out [count, n], rows
// of ws.q[0] = eigenvectors; info[b] != 0 on failure (QL did not converge in a leaf, NaN input).
int dc_solve(int n, int64_t count, const DcWs& ws, double* w_out, int* info, cudaStream_t st) {
  const int depth = dc_depth(n);
  cudaMemsetAsync(info, 0, (size_t)count * sizeof(int), st);
  if (depth > 0) {
    // zero both eigenvector buffers once: a node's block is block diagonal in its children, and whatever an
    // earlier level left behind lies inside diagonal blocks that are overwritten before they are read
    cudaMemsetAsync(ws.q[0], 0, (size_t)count * n * n * sizeof(double), st);
    cudaMemsetAsync(ws.q[1], 0, (size_t)count * n * n * sizeof(double), st);
  }
  const int leaf_buf = depth & 1;
  const int64_t warps = count << depth;
  dc_leaf_kernel<<<(unsigned)((warps + kDcLeafWarps - 1) / kDcLeafWarps), 32 * kDcLeafWarps, 0, st>>>(
      n, depth, count, ws, leaf_buf, w_out, info);
  SQB_LAUNCH_CHECK();
  for (int level = depth - 1; level >= 0; --level) {
    const int src = (level + 1) & 1;
    const int nmax = (n + (1 << level) - 1) >> level;
    const size_t smem = dc_setup_smem_bytes(nmax);
    cudaError_t e = cudaFuncSetAttribute(dc_merge_setup_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    dc_merge_setup_kernel<<<dim3(1u << level, (unsigned)count), kDcSetupThreads, smem, st>>>(n, level, nmax, ws, src,
                                                                                            w_out, info);
    SQB_LAUNCH_CHECK();
    const int tiles = (nmax + kDgTile - 1) / kDgTile;
    dc_merge_gemm_kernel<<<dim3((unsigned)(tiles * tiles), 1u << level, (unsigned)count), kDgThreads, 0, st>>>(n, level,
                                                                                                             ws, src);
    SQB_LAUNCH_CHECK();
  }
  return SQB_OK;
}

}  // namespace
