// K3c on UNIFORM time grids as a non-uniform FFT (type 1): included by evolve.cu (one translation unit).
//
// For t_r = t_0 + r dt the back-transformed evolution of a k-block (reference: dynamics/schrodinger.py:45-56 followed
// by StateList.with_state_basis, state/_state.py:153-165)
//
//     out[r, j] = sum_e c_e exp(-1j w_e t_r / hbar) V[e, j]
//
// is a sum of N complex exponentials in r whose angles theta_e = (w_e dt / hbar) mod 2 pi are the SAME for every
// column j of the block.  The GEMM form (evolve_bloch_3m_kernel) costs N multiply-adds per output element; here a tile
// of TT = 512 rows costs a gather of the N <= 512 sources onto an oversampled grid of NG = 1024 points (W = 14 window weights
// per source, shared by all columns of the block), one 1024-point FFT per column and a scale:
//
//     out[r0 + K m, j] = dec[m] * FFT_NG( g_j )[(m - TT/2) mod NG],                  (K = 1, 2, 4 ...: nu_interleave)
//     g_j[l]  = sum_e a_e[j] phi((l - u_e) / (W/2)),   u_e = (K theta_e mod 2 pi) NG / 2 pi,
//     a_e[j]  = c_e exp(-1j w_e t_{r0 + K TT/2} / hbar) V[e, j]       (the tile-centre phase is evaluated directly),
//
// phi(z) = exp(beta (sqrt(1 - z^2) - 1)) the "exponential of semicircle" window (Barnett, Magland, af Klinteberg,
// SIAM J. Sci. Comput. 41 (2019)), beta = 2.3 W, dec = 1 / (window transform).  Aliasing error ~1e-13 of sum_e |a_e|
// (tests/nufft_model.py is the numpy model of every step below; tests/test_nufft_model.py checks it against the
// direct sum).  ~130 flop per output element instead of 6 N = 1536 (3M GEMM) at N = 256.
//
// One CTA owns (k-block b, 8 adjacent columns j) and walks the row tiles: grid lines live in shared memory as
// G[l][8 j] (a grid point = 128 bytes = all 32 banks: every 16-byte access of a quarter-warp is conflict free whatever
// the stride in l), the FFT runs in place (decimation in frequency, radix 16 x 8 x 8, every thread owns whole
// butterflies so a stage needs one barrier), the last stage keeps only the 512 wanted frequencies and stores them
// straight to global memory in 128-byte runs.
#pragma once

namespace {

constexpr int kNuTT = 512;                    // rows per tile
constexpr int kNuNG = 1024;                   // oversampled grid points
constexpr int kNuW = 14;                      // window width in grid points
constexpr int kNuPad = 3;                     // zero weights either side (a thread gathers 4 consecutive grid points)
constexpr int kNuRow = kNuW + 2 * kNuPad;     // padded weight row
constexpr double kNuBeta = 2.3 * kNuW;
constexpr int kNuMaxN = 512;                  // sources (eigenstates) per block
constexpr int kNuChunk = 256;                 // sources resident in shared memory per gather pass
constexpr int kNuJ = 8;                       // columns per CTA
constexpr int kNuPcStride = kNuNG + 8;        // pc[x + 1] = #sources with first grid point <= x, x = -1 .. NG-1
#ifndef SQB_NU_THREADS
#define SQB_NU_THREADS 512  // developer switch (tools/build_variant.sh): 256 threads x 255 registers measured slower
#endif
constexpr int kNuThreads = SQB_NU_THREADS;
constexpr int kNuSlots = kNuThreads / kNuJ;   // threads along the grid axis

struct NuPlan {
  double* dec;          // [TT]      1 / window transform at m' = m - TT/2
  c128* tw;             // [NG/2]    exp(-2 pi i k / NG)
  unsigned long long* band;  // [2]  order-preserving keys of min / max of w dt / hbar over the call (nu_interleave)
  double* wts;          // [batch][n][kNuRow]  window weights of the SORTED sources, zero padded
  unsigned short* src;  // [batch][n]  sorted position -> eigenstate
  unsigned short* l0s;  // [batch][n]  first grid point of the window (mod NG), ascending
  unsigned short* pc;   // [batch][kNuPcStride]
};

__host__ __device__ inline size_t nufft_plan_bytes(int n, int64_t batch) {
  if (n > kNuMaxN) return 0;
  return (size_t)kNuTT * 8 + (size_t)(kNuNG / 2) * 16 + 256 +
         (size_t)batch * ((size_t)n * kNuRow * 8 + (size_t)n * 4 + (size_t)kNuPcStride * 2) + 1024;
}

inline NuPlan nufft_plan_at(void* base, int n, int64_t batch) {
  NuPlan p;
  unsigned char* q = reinterpret_cast<unsigned char*>(base);
  p.dec = reinterpret_cast<double*>(q);
  q += (size_t)kNuTT * 8;
  p.tw = reinterpret_cast<c128*>(q);
  q += (size_t)(kNuNG / 2) * 16;
  p.band = reinterpret_cast<unsigned long long*>(q);
  q += 256;
  p.wts = reinterpret_cast<double*>(q);
  q += (size_t)batch * n * kNuRow * 8;
  p.src = reinterpret_cast<unsigned short*>(q);
  q += (size_t)batch * n * 2;
  p.l0s = reinterpret_cast<unsigned short*>(q);
  q += (size_t)batch * n * 2;
  p.pc = reinterpret_cast<unsigned short*>(q);
  return p;
}

// ---- interleaved row tiles.  Physical time steps resolve the dynamics, so the angles theta_e = w_e dt / hbar of a
// call usually fill a small part of the circle (cfg3: 0.45 rad) and every source would fall on the same few grid
// points - a few threads of the gather would do all the work.  A tile therefore takes every K-th row of a super-tile
// of K * 512 rows, rows r0 + K m + m0 (m0 = the tile's index in the super-tile): its angles are K theta_e, K times
// further apart, at the same cost per tile.  K = the largest power of two that divides the call's tile count and keeps
// K * (max - min of w dt / hbar) below 0.8 * 2 pi; it is computed on the device (no host synchronisation) by every
// kernel from the min / max the band kernel leaves in the workspace.
__device__ __forceinline__ unsigned long long nu_key(double x) {
  const unsigned long long b = (unsigned long long)__double_as_longlong(x);
  return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double nu_unkey(unsigned long long k) {
  return __longlong_as_double((long long)((k >> 63) ? (k & 0x7fffffffffffffffull) : ~k));
}
__device__ __forceinline__ int nu_interleave(const unsigned long long* band, int64_t tiles) {
  const double span = nu_unkey(band[1]) - nu_unkey(band[0]);
  int64_t k_max = tiles & (-tiles);  // largest power of two dividing the tile count: super-tiles are complete
  if (k_max > 64) k_max = 64;
  int k = 1;
  while (2 * k <= k_max && 2.0 * k * span < 0.8 * 6.283185307179586) k *= 2;
  return k;
}

__global__ void __launch_bounds__(256)
nufft_band_kernel(int64_t m, const double* __restrict__ w, double dt, double hbar, unsigned long long* band) {
  __shared__ double lo_s[8], hi_s[8];
  double lo = 1.0e300, hi = -1.0e300;
  for (int64_t g = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; g < m; g += (int64_t)gridDim.x * blockDim.x) {
    const double x = (w[g] * dt) / hbar;
    lo = fmin(lo, x);
    hi = fmax(hi, x);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    lo = fmin(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    hi = fmax(hi, __shfl_xor_sync(0xffffffffu, hi, o));
  }
  if ((threadIdx.x & 31) == 0) {
    lo_s[threadIdx.x >> 5] = lo;
    hi_s[threadIdx.x >> 5] = hi;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 1; i < 8; ++i) {
      lo = fmin(lo, lo_s[i]);
      hi = fmax(hi, hi_s[i]);
    }
    if (lo <= hi) {
      atomicMin(band, nu_key(lo));
      atomicMax(band + 1, nu_key(hi));
    }
  }
}

// dec[m] = 1 / ((W/2) int_{-1}^{1} phi(z) cos(alpha_m z) dz), alpha_m = (m - TT/2) (2 pi / NG) (W/2); with z = sin t
// the integrand is smooth and (up to exp(-beta)) periodic, so the midpoint rule converges geometrically.  One warp
// per m, 256 nodes.
__global__ void __launch_bounds__(256)
nufft_deconv_kernel(double* __restrict__ dec, c128* __restrict__ tw, unsigned long long* __restrict__ band) {
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    band[0] = ~0ull;  // min key
    band[1] = 0ull;   // max key
  }
  {
    // the FFT's half twiddle table, one entry per thread of the grid (64 CTAs x 256 >= NG / 2)
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i < kNuNG / 2) {
      double s, c;
      sincospi(-(double)i / (kNuNG / 2), &s, &c);
      tw[i] = make_double2(c, s);
    }
  }
  const int m = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (m >= kNuTT) return;
  const double alpha = (double)(m - kNuTT / 2) * (6.283185307179586 / kNuNG) * (0.5 * kNuW);
  double acc = 0.0;
  for (int q = lane; q < 256; q += 32) {
    const double t = -1.5707963267948966 + ((double)q + 0.5) * (3.141592653589793 / 256.0);
    double s, c;
    sincos(t, &s, &c);
    acc += exp(kNuBeta * (c - 1.0)) * c * cos(alpha * s);
  }
  acc = warp_sum(acc);
  if (lane == 0) dec[m] = 1.0 / ((0.5 * kNuW) * acc * (3.141592653589793 / 256.0));
}

// One CTA per k-block: angles, windows, sources sorted by the first grid point of their window.
__global__ void __launch_bounds__(kNuMaxN)
nufft_plan_kernel(int n, const double* __restrict__ w, double dt, double hbar, int64_t tiles, NuPlan p, int* bsum_flag,
                  int clear_flag) {
  __shared__ int key_s[kNuMaxN];
  __shared__ unsigned short l0_sorted[kNuMaxN];
  const int64_t b = blockIdx.x;
  const int tid = threadIdx.x;
  if (clear_flag && b == 0 && tid == 0) *bsum_flag = 0;  // the transform changed: the GEMM route's Re+Im plane is stale
  double u = 0.0;
  int l0 = 0, key = 0x7fffffff;
  if (tid < n) {
    const double x = (double)nu_interleave(p.band, tiles) * ((w[b * n + tid] * dt) / hbar);  // angle per row of a tile
    const double q = rint(x / 6.283185307179586);
    double th = fma(-q, 6.283185307179586, x);
    th = fma(-q, 2.4492935982947064e-16, th);
    if (th < 0.0) th += 6.283185307179586;
    u = th * (kNuNG / 6.283185307179586);
    l0 = (int)ceil(u - 0.5 * kNuW);
    key = ((l0 + kNuNG) % kNuNG) * kNuMaxN + tid;
  }
  key_s[tid] = key;
  __syncthreads();
  if (tid < n) {
    int rank = 0;
    for (int i = 0; i < n; ++i) rank += key_s[i] < key;
    const int l0w = key / kNuMaxN;
    p.src[b * n + rank] = (unsigned short)tid;
    p.l0s[b * n + rank] = (unsigned short)l0w;
    l0_sorted[rank] = (unsigned short)l0w;
    double* row = p.wts + (b * n + rank) * kNuRow;
    for (int k = 0; k < kNuRow; ++k) {
      const int kk = k - kNuPad;
      double val = 0.0;
      if (kk >= 0 && kk < kNuW) {
        const double z = ((double)(l0 + kk) - u) / (0.5 * kNuW);
        val = exp(kNuBeta * (sqrt(fmax(0.0, 1.0 - z * z)) - 1.0));
      }
      row[k] = val;
    }
  }
  __syncthreads();
  for (int i = tid; i <= kNuNG; i += blockDim.x) {
    const int x = i - 1;
    int lo = 0, hi = n;
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if ((int)l0_sorted[mid] <= x) lo = mid + 1;
      else hi = mid;
    }
    p.pc[b * kNuPcStride + i] = (unsigned short)lo;
  }
}

// ---- register butterflies (forward transform, exp(-2 pi i n k / R)) ---------------------------------------------
__device__ __forceinline__ c128 nu_mul_mi(c128 a) { return make_double2(a.y, -a.x); }  // a * (-i)

__device__ __forceinline__ void nu_dft4(c128& a0, c128& a1, c128& a2, c128& a3) {
  const c128 b0 = cadd(a0, a2), b1 = csub(a0, a2), b2 = cadd(a1, a3), b3 = nu_mul_mi(csub(a1, a3));
  a0 = cadd(b0, b2);
  a1 = cadd(b1, b3);
  a2 = csub(b0, b2);
  a3 = csub(b1, b3);
}

// x[0..7] -> X[0..7] in place
__device__ __forceinline__ void nu_dft8(c128* x) {
  constexpr double r = 0.7071067811865476;
  c128 e0 = x[0], e1 = x[2], e2 = x[4], e3 = x[6], o0 = x[1], o1 = x[3], o2 = x[5], o3 = x[7];
  nu_dft4(e0, e1, e2, e3);
  nu_dft4(o0, o1, o2, o3);
  o1 = make_double2(r * (o1.x + o1.y), r * (o1.y - o1.x));   // * (1 - i) / sqrt 2
  o2 = nu_mul_mi(o2);
  o3 = make_double2(r * (o3.y - o3.x), -r * (o3.x + o3.y));  // * (-1 - i) / sqrt 2
  x[0] = cadd(e0, o0);
  x[1] = cadd(e1, o1);
  x[2] = cadd(e2, o2);
  x[3] = cadd(e3, o3);
  x[4] = csub(e0, o0);
  x[5] = csub(e1, o1);
  x[6] = csub(e2, o2);
  x[7] = csub(e3, o3);
}

// x[0..15] -> X[0..15] in place: 4 x 4 Cooley-Tukey, n = 4 n1 + n0, k = k0 + 4 k1
__device__ __forceinline__ void nu_dft16(c128* x) {
  constexpr double c1 = 0.9238795325112867, s1 = 0.3826834323650898, r = 0.7071067811865476;
  c128 y[4][4];  // y[n0][k0]
#pragma unroll
  for (int n0 = 0; n0 < 4; ++n0) {
    y[n0][0] = x[n0];
    y[n0][1] = x[4 + n0];
    y[n0][2] = x[8 + n0];
    y[n0][3] = x[12 + n0];
    nu_dft4(y[n0][0], y[n0][1], y[n0][2], y[n0][3]);
  }
  // twiddles w16^(n0 k0)
  y[1][1] = cmul(y[1][1], make_double2(c1, -s1));
  y[1][2] = make_double2(r * (y[1][2].x + y[1][2].y), r * (y[1][2].y - y[1][2].x));
  y[1][3] = cmul(y[1][3], make_double2(s1, -c1));
  y[2][1] = make_double2(r * (y[2][1].x + y[2][1].y), r * (y[2][1].y - y[2][1].x));
  y[2][2] = nu_mul_mi(y[2][2]);
  y[2][3] = make_double2(r * (y[2][3].y - y[2][3].x), -r * (y[2][3].x + y[2][3].y));
  y[3][1] = cmul(y[3][1], make_double2(s1, -c1));
  y[3][2] = make_double2(r * (y[3][2].y - y[3][2].x), -r * (y[3][2].x + y[3][2].y));
  y[3][3] = cmul(y[3][3], make_double2(-c1, s1));
#pragma unroll
  for (int k0 = 0; k0 < 4; ++k0) {
    nu_dft4(y[0][k0], y[1][k0], y[2][k0], y[3][k0]);  // index n0 -> k1
    x[k0] = y[0][k0];
    x[k0 + 4] = y[1][k0];
    x[k0 + 8] = y[2][k0];
    x[k0 + 12] = y[3][k0];
  }
}

constexpr size_t kNuSmemG = (size_t)kNuNG * kNuJ * sizeof(c128);        // 131072
constexpr size_t kNuSmemA = (size_t)kNuChunk * kNuJ * sizeof(c128);     // 32768
constexpr size_t kNuSmemWt = (size_t)kNuChunk * kNuRow * sizeof(double);  // 40960
constexpr size_t kNuSmemTw = (size_t)(kNuNG / 2) * sizeof(c128);        // 8192: w^k, k < NG/2 (w^(k + NG/2) = -w^k)
constexpr size_t kNuSmemDec = (size_t)kNuTT * sizeof(double);           // 4096
constexpr size_t kNuSmemPh = (size_t)kNuChunk * sizeof(c128);           // 4096
constexpr size_t kNuSmemPc = (size_t)kNuPcStride * 2;                   // 2064
constexpr size_t kNuSmemIdx = (size_t)kNuMaxN * 2;                      // 1024 (x 2: src, l0s)
constexpr size_t kNuSmem =
    kNuSmemG + kNuSmemA + kNuSmemWt + kNuSmemTw + kNuSmemDec + kNuSmemPh + kNuSmemPc + 2 * kNuSmemIdx;
constexpr int kNuVreg = kNuChunk * kNuJ / kNuThreads;  // eigenvector elements a thread keeps across the row tiles

// w_1024^k for k < NG from the half table
__device__ __forceinline__ c128 nu_tw(const c128* tw, int k) {
  const c128 t = tw[k & (kNuNG / 2 - 1)];
  return (k & (kNuNG / 2)) ? make_double2(-t.x, -t.y) : t;
}

__global__ void __launch_bounds__(kNuThreads, 1)
evolve_nufft_kernel(int n, const c128* __restrict__ V, const double* __restrict__ wv, const c128* __restrict__ cv,
                    const double* __restrict__ times, int64_t t_count, double dt, double hbar, c128* __restrict__ out,
                    int64_t out_row_stride, NuPlan p, int groups_per_cta) {
  extern __shared__ __align__(16) unsigned char nu_smem[];
  c128* G = reinterpret_cast<c128*>(nu_smem);
  c128* A = reinterpret_cast<c128*>(nu_smem + kNuSmemG);
  double* wt = reinterpret_cast<double*>(nu_smem + kNuSmemG + kNuSmemA);
  c128* tw = reinterpret_cast<c128*>(nu_smem + kNuSmemG + kNuSmemA + kNuSmemWt);
  double* decs = reinterpret_cast<double*>(nu_smem + kNuSmemG + kNuSmemA + kNuSmemWt + kNuSmemTw);
  c128* ph = reinterpret_cast<c128*>(nu_smem + kNuSmemG + kNuSmemA + kNuSmemWt + kNuSmemTw + kNuSmemDec);
  unsigned short* pcs = reinterpret_cast<unsigned short*>(nu_smem + kNuSmemG + kNuSmemA + kNuSmemWt + kNuSmemTw +
                                                         kNuSmemDec + kNuSmemPh);
  unsigned short* srcs = pcs + kNuPcStride;
  unsigned short* l0ss = srcs + kNuMaxN;

  const int tid = threadIdx.x;
  const int j = tid & (kNuJ - 1), slot = tid >> 3;
  const int64_t b = blockIdx.y;

  // blocks of more than kNuChunk eigenstates are gathered in passes of kNuChunk SORTED sources (the window weights
  // and the phase x eigenvector products of one pass fill the shared memory left beside the grid lines)
  const int n_chunks = (n + kNuChunk - 1) / kNuChunk;
  const double* wsrc = p.wts + b * (int64_t)n * kNuRow;
  // ---- global loads that feed registers first (they depend on the plan in GLOBAL memory only, so their latency
  // overlaps the shared-memory fills below): the source's eigenvalue and coefficient, and - for blocks of one pass -
  // the block's eigenvector elements of this CTA's columns, which stay in registers for every row tile
  double we = 0.0;
  c128 ce = make_double2(0.0, 0.0);
  if (tid < n) {
    const int e = p.src[b * n + tid];
    we = wv[b * n + e];
    ce = cv[b * n + e];
  }
  // ---- per-block plan and tables -> shared memory
  {
    if (n_chunks == 1) {
      const double2* w2 = reinterpret_cast<const double2*>(wsrc);  // rows of kNuRow = 20 doubles: 16-byte aligned
      double2* d2 = reinterpret_cast<double2*>(wt);
      for (int i = tid; i < n * kNuRow / 2; i += kNuThreads) d2[i] = w2[i];
    }
    for (int i = tid; i < kNuNG + 1; i += kNuThreads) pcs[i] = p.pc[b * kNuPcStride + i];
    for (int i = tid; i < n; i += kNuThreads) {
      srcs[i] = p.src[b * n + i];
      l0ss[i] = p.l0s[b * n + i];
    }
    for (int i = tid; i < kNuNG / 2; i += kNuThreads) tw[i] = p.tw[i];
    for (int i = tid; i < kNuTT; i += kNuThreads) decs[i] = p.dec[i];
  }
  __syncthreads();

  // A CTA takes `groups_per_cta` adjacent groups of 8 columns one after the other, so that the 55 KB of per-block
  // tables above are fetched once per ~8 tile-sized work items even when a call has a single row tile
  const int64_t tiles = (t_count + kNuTT - 1) / kNuTT;
  const int kint = nu_interleave(p.band, tiles);
  c128 vreg[kNuVreg];
  auto load_vreg = [&](int jbase) {
#pragma unroll
    for (int i = 0; i < kNuVreg; ++i) {
      const int idx = tid + i * kNuThreads;
      vreg[i] = make_double2(0.0, 0.0);
      if (n_chunks == 1 && idx < n * kNuJ && jbase + j < n)
        vreg[i] = __ldg(V + (b * n + srcs[idx >> 3]) * (int64_t)n + jbase + j);
    }
  };
  load_vreg(blockIdx.x * groups_per_cta * kNuJ);
  for (int gi = 0; gi < groups_per_cta; ++gi) {
  const int j0 = (blockIdx.x * groups_per_cta + gi) * kNuJ;
  if (j0 >= n) break;
  const bool col_ok = j0 + j < n;
  for (int64_t ti = 0; ti < tiles; ++ti) {
    // interleaved tile: rows r0 + kint m, m = 0 .. 511 (see nu_interleave)
    const int64_t r0 = (ti / kint) * kint * kNuTT + (ti % kint);
    for (int chunk = 0; chunk < n_chunks; ++chunk) {
      const int cs = chunk * kNuChunk, cnt = min(kNuChunk, n - cs);
      if (n_chunks > 1) {
        if (chunk > 0) __syncthreads();  // the previous pass has finished reading wt / A
        for (int i = tid; i < cnt * kNuRow; i += kNuThreads) wt[i] = wsrc[cs * kNuRow + i];
      }
      // ---- tile-centre phases, then A[s][j] = c_e exp(-1j w_e t_c / hbar) V[e][j0 + j]
      if (tid >= cs && tid < cs + cnt) {
        const double tc = times[r0] + (double)(kint * (kNuTT / 2)) * dt;
        double s, c;
        sincos(-(we * tc) / hbar, &s, &c);
        ph[tid - cs] = cmul(ce, make_double2(c, s));
      }
      __syncthreads();
      if (n_chunks == 1) {
#pragma unroll
        for (int i = 0; i < kNuVreg; ++i) {
          const int idx = tid + i * kNuThreads;
          if (idx < n * kNuJ) A[idx] = cmul(ph[idx >> 3], vreg[i]);
        }
        // last row tile of this column group: the registers are free, fetch the next group's eigenvector elements
        // now so that their latency hides behind this tile's gather and FFT
        if (ti + 1 == tiles && gi + 1 < groups_per_cta) load_vreg(j0 + kNuJ);
      } else {
        for (int idx = tid; idx < cnt * kNuJ; idx += kNuThreads) {
          c128 v = make_double2(0.0, 0.0);
          if (col_ok) v = __ldg(V + (b * n + srcs[cs + (idx >> 3)]) * (int64_t)n + j0 + j);
          A[idx] = cmul(ph[idx >> 3], v);
        }
      }
      __syncthreads();

      // ---- gather: a thread owns 4 consecutive grid points of column j
      for (int quad = slot; quad < kNuNG / 4; quad += kNuSlots) {
        const int l = 4 * quad;
        c128 acc[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[i] = chunk == 0 ? make_double2(0.0, 0.0) : G[(l + i) * kNuJ + j];
#pragma unroll 1
        for (int wrap = 0; wrap < 2; ++wrap) {
          if (wrap == 1 && l >= kNuW + kNuPad) break;
          const int base = l + wrap * kNuNG;
          const int lo = base - kNuW, hi = base + 3;
          const int s0 = max(cs, (int)pcs[(lo < -1 ? -1 : (lo > kNuNG - 1 ? kNuNG - 1 : lo)) + 1]);
          const int s1 = min(cs + cnt, (int)pcs[(hi > kNuNG - 1 ? kNuNG - 1 : hi) + 1]);
          for (int s = s0; s < s1; ++s) {
            const double* wr = wt + (s - cs) * kNuRow + kNuPad + (base - (int)l0ss[s]);
            const c128 a = A[(s - cs) * kNuJ + j];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const double wgt = wr[i];
              acc[i].x = fma(wgt, a.x, acc[i].x);
              acc[i].y = fma(wgt, a.y, acc[i].y);
            }
          }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) G[(l + i) * kNuJ + j] = acc[i];
      }
    }
    __syncthreads();

    // ---- FFT stage 1: radix 16 over n1 (stride 64), twiddle w_1024^(n2 k1)
    for (int n2 = slot; n2 < 64; n2 += kNuSlots) {
      c128 x[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) x[i] = G[(64 * i + n2) * kNuJ + j];
      nu_dft16(x);
#pragma unroll
      for (int i = 1; i < 16; ++i) x[i] = cmul(x[i], nu_tw(tw, n2 * i));
#pragma unroll
      for (int i = 0; i < 16; ++i) G[(64 * i + n2) * kNuJ + j] = x[i];
    }
    __syncthreads();
    // ---- stage 2: radix 8 over n2a (stride 8) inside every block of 64, twiddle w_64^(n2b k2)
    for (int idx = slot; idx < 128; idx += kNuSlots) {
      const int k1 = idx >> 3, n2b = idx & 7;
      c128* g = G + (64 * k1 + n2b) * kNuJ + j;
      c128 x[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) x[i] = g[8 * i * kNuJ];
      nu_dft8(x);
#pragma unroll
      for (int i = 1; i < 8; ++i) x[i] = cmul(x[i], nu_tw(tw, 16 * n2b * i));
#pragma unroll
      for (int i = 0; i < 8; ++i) g[8 * i * kNuJ] = x[i];
    }
    __syncthreads();
    // ---- stage 3: radix 8 over n2b (stride 1); frequency k = k1 + 16 k2 + 128 k3, rows m = k + 256 (k3 = 0, 1) and
    // m = k - 768 (k3 = 6, 7) are the tile; scale and store (128-byte runs, streaming)
    for (int idx = slot; idx < 128; idx += kNuSlots) {
      const int k1 = idx >> 3, k2 = idx & 7;
      const c128* g = G + (64 * k1 + 8 * k2) * kNuJ + j;
      c128 x[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) x[i] = g[i * kNuJ];
      nu_dft8(x);
      const int mlow = k1 + 16 * k2;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        // q = 0, 1: k3 = 6, 7 -> rows mlow, mlow + 128;  q = 2, 3: k3 = 0, 1 -> rows mlow + 256, mlow + 384
        const int m = mlow + 128 * q;
        const c128 val = cscale(x[q < 2 ? 6 + q : q - 2], decs[m]);
        const int64_t r = r0 + (int64_t)kint * m;
        if (r < t_count && col_ok) __stcs(out + r * out_row_stride + b * (int64_t)n + j0 + j, val);
      }
    }
    // the two barriers of the next tile's phase / A stages order these reads of G before the next gather
  }
  }  // column groups
}

int launch_evolve_nufft(int n, int64_t batch, const sqb_c128* transform, const double* w, const sqb_c128* c,
                        const double* times, int64_t t_count, double dt, double hbar, sqb_c128* out,
                        int64_t out_row_stride, void* plan_base, int* bsum_flag, int clear_flag, cudaStream_t st) {
  cudaError_t e =
      cudaFuncSetAttribute(evolve_nufft_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kNuSmem);
  if (e != cudaSuccess) return (int)e;
  const NuPlan p = nufft_plan_at(plan_base, n, batch);
  const int64_t tiles = (t_count + kNuTT - 1) / kNuTT;
  nufft_deconv_kernel<<<kNuTT / 8, 256, 0, st>>>(p.dec, p.tw, p.band);
  SQB_LAUNCH_CHECK();
  nufft_band_kernel<<<(unsigned)sqb_min64((batch * n + 255) / 256, 148 * 4), 256, 0, st>>>(batch * n, w, dt, hbar, p.band);
  SQB_LAUNCH_CHECK();
  nufft_plan_kernel<<<(unsigned)batch, n <= 256 ? 256 : kNuMaxN, 0, st>>>(n, w, dt, hbar, tiles, p, bsum_flag,
                                                                         clear_flag);
  SQB_LAUNCH_CHECK();
  const int groups = (n + kNuJ - 1) / kNuJ;
  // ~8 tile-sized work items per CTA, as long as the grid keeps >= 8 CTAs per SM (measured at cfg3's block shape, 512
  // rows per call: 11.6 ms with one group per CTA, 9.9 with four)
  int groups_per_cta = tiles >= 8 ? 1 : (tiles >= 4 ? 2 : (tiles >= 2 ? 4 : 8));
  groups_per_cta = (int)sqb_min64(groups_per_cta, sqb_min64(groups, (groups * batch + 1183) / 1184));
  if (groups_per_cta < 1) groups_per_cta = 1;
  const unsigned jg = (unsigned)((groups + groups_per_cta - 1) / groups_per_cta);
  for (int64_t b0 = 0; b0 < batch; b0 += 65535) {
    const int64_t cnt = sqb_min64(65535, batch - b0);
    NuPlan q = p;
    q.wts += b0 * (int64_t)n * kNuRow;
    q.src += b0 * n;
    q.l0s += b0 * n;
    q.pc += b0 * kNuPcStride;
    evolve_nufft_kernel<<<dim3(jg, (unsigned)cnt), kNuThreads, kNuSmem, st>>>(
        n, reinterpret_cast<const c128*>(transform) + b0 * (int64_t)n * n, w + b0 * n,
        reinterpret_cast<const c128*>(c) + b0 * n, times, t_count, dt, hbar, reinterpret_cast<c128*>(out) + b0 * n,
        out_row_stride, q, groups_per_cta);
    SQB_LAUNCH_CHECK();
  }
  return SQB_OK;
}

}  // namespace
