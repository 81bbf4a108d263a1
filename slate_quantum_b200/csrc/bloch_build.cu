// K1: Bloch block builder + kinetic diagonals, and K5: Bloch index permutation.
//
// K1 writes H[b,k,k'] = E_b[k] delta_kk' + vtable[(k'-k) mod shape]/sqrt(N) for every k-block of a launch.
// Pure HBM-write kernel: 16*N^2 algorithmic bytes per block, each thread stores 16 B with a warp covering
// 512 contiguous bytes of one row; the (<= 64 KB) potential table sits in shared memory.
//
// Reference arithmetic restated (file:line in Matt-Ord/slate_quantum):
//   kinetic diagonal  operator/_build/_hamiltonian.py:32-37 (wrap), :40-48 (bloch phase),
//                     :72-94 (stacked k points), :121-146 (energy)
//   block order       bloch/build.py:40-42, 138-143
//   densification     bloch/build.py:123-128 (closed form of the slate_core SplitBasis conversion)
#include <math.h>

#include "sqb_common.cuh"

namespace {

struct BuildParams {
  SqbDims d;
  int n;                               // prod(shape)
  double dk[SQB_MAX_DIMS * SQB_MAX_DIMS];  // row i = reciprocal vector i
  double half_delta[SQB_MAX_DIMS];     // |N_j dk_j| / 2   (norm of delta_k row j, halved)
  double mass, hbar;
  double inv_sqrt_n_den;               // sqrt(N)
  int use_fraction;                    // 1: every block uses `fraction` instead of its own k-point
  double fraction[SQB_MAX_DIMS];
};

// FFT-order signed integer of array index i on an axis of length n: [0,1,..,ceil(n/2)-1,-floor(n/2),..,-1]
__device__ __forceinline__ int fft_int(int i, int n) { return i < (n + 1) / 2 ? i : i - n; }

// np.mod(a, b) for b > 0 (python-style result in [0, b))
__device__ __forceinline__ double np_mod_pos(double a, double b) {
  double m = fmod(a, b);
  if (m != 0.0) {
    if (m < 0.0) m = __dadd_rn(m, b);
  } else {
    m = 0.0;
  }
  return m;
}

// Kinetic energy of in-cell state p (multi index pidx) in block with multi index bidx.
// No FMA contraction: mirror numpy's separate multiply / add roundings.
__device__ double kinetic_value(const BuildParams& P, const int* pidx, const int* bidx) {
  const int nd = P.d.nd;
  double e = 0.0;
  for (int j = 0; j < nd; ++j) {  // Cartesian component j
    double acc = 0.0, phase = 0.0;
    for (int i = 0; i < nd; ++i) {
      const double dkij = P.dk[i * SQB_MAX_DIMS + j];
      const double ni = (double)fft_int(pidx[i], P.d.shape[i]);
      const double fi = P.use_fraction ? P.fraction[i] : (double)fft_int(bidx[i], P.d.repeat[i]) / (double)P.d.repeat[i];
      acc = __dadd_rn(acc, __dmul_rn(dkij, ni));
      phase = __dadd_rn(phase, __dmul_rn(dkij, fi));
    }
    const double pt = __dadd_rn(acc, phase);
    const double shift = P.half_delta[j];
    const double wrapped = __dadd_rn(np_mod_pos(__dadd_rn(pt, shift), __dmul_rn(2.0, shift)), -shift);
    const double hk = __dmul_rn(P.hbar, wrapped);
    e = __dadd_rn(e, __ddiv_rn(__dmul_rn(hk, hk), __dmul_rn(2.0, P.mass)));
  }
  return e;
}

__device__ __forceinline__ void unravel(int64_t flat, const int* dims, int nd, int* out) {
  for (int a = nd - 1; a >= 0; --a) {
    out[a] = (int)(flat % dims[a]);
    flat /= dims[a];
  }
}

constexpr int kBuildThreads = 256;
constexpr int kRowsPerCta = 32;

// 32-bit unravel (n <= 2^20)
__device__ __forceinline__ void unravel32(int flat, const int* dims, int nd, int* out) {
  for (int a = nd - 1; a >= 0; --a) {
    const int q = flat / dims[a];
    out[a] = flat - q * dims[a];
    flat = q;
  }
}

__global__ void __launch_bounds__(kBuildThreads)
bloch_build_kernel(BuildParams P, const c128* __restrict__ vtable, int64_t block_begin, c128* __restrict__ H) {
  extern __shared__ c128 s_tab[];                 // [n] potential table / sqrt(N)
  __shared__ int s_ridx[kRowsPerCta][SQB_MAX_DIMS];  // multi index of each row of the tile
  __shared__ double s_kin[kRowsPerCta];              // kinetic energy of each row of the tile
  const int n = P.n;
  const int nd = P.d.nd;
  for (int i = threadIdx.x; i < n; i += kBuildThreads) {
    c128 v = vtable[i];
    s_tab[i] = make_double2(__ddiv_rn(v.x, P.inv_sqrt_n_den), __ddiv_rn(v.y, P.inv_sqrt_n_den));
  }
  const int64_t b_local = blockIdx.x;
  const int row0 = blockIdx.y * kRowsPerCta;
  const int rows = min(kRowsPerCta, n - row0);
  if (threadIdx.x < rows) {
    int bidx[SQB_MAX_DIMS], ridx[SQB_MAX_DIMS];
    unravel(block_begin + b_local, P.d.repeat, nd, bidx);
    unravel32(row0 + threadIdx.x, P.d.shape, nd, ridx);
    for (int a = 0; a < SQB_MAX_DIMS; ++a) s_ridx[threadIdx.x][a] = a < nd ? ridx[a] : 0;
    s_kin[threadIdx.x] = kinetic_value(P, ridx, bidx);
  }
  __syncthreads();

  c128* Hb = H + b_local * (int64_t)n * n + (int64_t)row0 * n;
  int stride[SQB_MAX_DIMS];
  {
    int st = 1;
    for (int a = nd - 1; a >= 0; --a) {
      stride[a] = st;
      st *= P.d.shape[a];
    }
  }
  for (int c = threadIdx.x; c < n; c += kBuildThreads) {
    int cidx[SQB_MAX_DIMS];
    unravel32(c, P.d.shape, nd, cidx);
    for (int r = 0; r < rows; ++r) {
      int flat = 0;
      for (int a = 0; a < nd; ++a) {
        int dlt = cidx[a] - s_ridx[r][a];
        if (dlt < 0) dlt += P.d.shape[a];
        flat += dlt * stride[a];
      }
      c128 v = s_tab[flat];
      if (c == row0 + r) v.x = __dadd_rn(v.x, s_kin[r]);
      Hb[(int64_t)r * n + c] = v;
    }
  }
}

__global__ void bloch_kinetic_kernel(BuildParams P, int64_t block_begin, int64_t block_count, c128* __restrict__ E) {
  const int64_t total = block_count * P.n;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int bidx[SQB_MAX_DIMS], pidx[SQB_MAX_DIMS];
    unravel(block_begin + i / P.n, P.d.repeat, P.d.nd, bidx);
    unravel(i % P.n, P.d.shape, P.d.nd, pidx);
    E[i] = make_double2(kinetic_value(P, pidx, bidx), 0.0);
  }
}

// K5: bloch flat index i = b*N + p  <->  supercell flat index s(i).
__global__ void bloch_permute_kernel(SqbDims d, int n, int64_t m, int64_t lead, int direction,
                                     const c128* __restrict__ in, c128* __restrict__ out) {
  const int64_t total = lead * m;
  for (int64_t g = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; g < total; g += (int64_t)gridDim.x * blockDim.x) {
    const int64_t l = g / m;
    const int64_t i = g - l * m;
    int bidx[SQB_MAX_DIMS], pidx[SQB_MAX_DIMS];
    unravel(i / n, d.repeat, d.nd, bidx);
    unravel(i % n, d.shape, d.nd, pidx);
    int64_t s = 0;
    for (int a = 0; a < d.nd; ++a) {
      const int big = d.shape[a] * d.repeat[a];
      int idx = fft_int(pidx[a], d.shape[a]) * d.repeat[a] + fft_int(bidx[a], d.repeat[a]);
      idx %= big;
      if (idx < 0) idx += big;
      s = s * big + idx;
    }
    if (direction == 0) {
      out[g] = in[l * m + s];
    } else {
      out[l * m + s] = in[g];
    }
  }
}

int fill_params(BuildParams& P, int nd, const int* shape, const int* repeat, const double* dk, double mass,
                double hbar, int64_t* nk) {
  int64_t n64;
  int rc = sqb_fill_dims(P.d, nd, shape, repeat, &n64, nk);
  if (rc) return rc;
  if (!dk || !(mass > 0.0) || n64 > (1 << 20)) return SQB_E_BADARG;
  P.n = (int)n64;
  P.mass = mass;
  P.hbar = hbar;
  P.inv_sqrt_n_den = sqrt((double)n64);
  P.use_fraction = 0;
  for (int i = 0; i < SQB_MAX_DIMS; ++i) P.fraction[i] = 0.0;
  for (int i = 0; i < SQB_MAX_DIMS * SQB_MAX_DIMS; ++i) P.dk[i] = 0.0;
  for (int i = 0; i < nd; ++i)
    for (int j = 0; j < nd; ++j) P.dk[i * SQB_MAX_DIMS + j] = dk[i * nd + j];
  for (int j = 0; j < SQB_MAX_DIMS; ++j) P.half_delta[j] = 1.0;
  for (int j = 0; j < nd; ++j) {
    // np.linalg.norm(delta_k, axis=1)[j], delta_k[j] = N_j * dk_j   (_hamiltonian.py:89-92)
    double s = 0.0;
    for (int c = 0; c < nd; ++c) {
      const double v = dk[j * nd + c] * (double)shape[j];
      s += v * v;
    }
    P.half_delta[j] = sqrt(s) / 2.0;
  }
  return SQB_OK;
}

// Time-reversal partner of (block multi-index, in-block multi-index): on an orthogonal lattice with a potential table
// that is Hermitian-symmetric (a real V(x)) the blocks of k and -k satisfy
//     H_{-k}[pi g, pi g'] = conj(H_k[g, g'])
// where the supercell momentum index s = q + R g (signed FFT-order values, mod M = R N per axis) maps to -s.  The
// eigenvalues of the two blocks are equal and the eigenvectors are V_{-k}[e, pi g] = conj(V_k[e, g]).
// (operator/_build/_hamiltonian.py:121-146: the kinetic diagonal is even in the wrapped k-point;
//  bloch/build.py:123-128: the potential part is the convolution matrix of the table.)
__device__ __forceinline__ void time_reversal_partner(const SqbDims& d, int64_t blk, int k, int64_t* blk_out, int* k_out) {
  int64_t b_out = 0, b_mul = 1;
  int g_out = 0, g_mul = 1;
  for (int a = d.nd - 1; a >= 0; --a) {
    const int r = d.repeat[a], na = d.shape[a];
    const int iq = (int)(blk % r), ig = k % na;
    blk /= r;
    k /= na;
    const int64_t m = (int64_t)r * na;
    int64_t s = ((int64_t)fft_int(iq, r) + (int64_t)r * fft_int(ig, na)) % m;
    if (s < 0) s += m;
    const int64_t sp = (m - s) % m;
    const int qi = (int)(sp % r);
    int64_t gi = ((sp - fft_int(qi, r)) / r) % na;
    if (gi < 0) gi += na;
    b_out += b_mul * qi;
    b_mul *= r;
    g_out += g_mul * (int)gi;
    g_mul *= na;
  }
  *blk_out = b_out;
  *k_out = g_out;
}

// one CTA per (target block, 8 eigenvector rows): the partner block and the in-block permutation are computed once per
// thread column; writes are contiguous, reads run through the permutation (mostly reversed runs)
__global__ void __launch_bounds__(256)
time_reversal_mirror_kernel(SqbDims d, int n, int64_t first, int64_t count, c128* __restrict__ v, double* __restrict__ w) {
  const int64_t bt = first + blockIdx.y;  // target block (global index)
  if (bt >= first + count) return;
  for (int k = threadIdx.x; k < n; k += blockDim.x) {
    int64_t bs;
    int ks;
    time_reversal_partner(d, bt, k, &bs, &ks);
    const c128* src = v + bs * (int64_t)n * n;
    c128* dst = v + bt * (int64_t)n * n;
    for (int e = blockIdx.x * 8; e < min(n, blockIdx.x * 8 + 8); ++e) {
      const c128 z = src[(int64_t)e * n + ks];
      dst[(int64_t)e * n + k] = make_double2(z.x, -z.y);
    }
    if (blockIdx.x == 0) w[bt * n + k] = w[bs * n + k];  // eigenvalue e = k here: same spectrum, same order
  }
}

}  // namespace

extern "C" int sqb_bloch_build(int nd, const int* shape_host, const int* repeat_host, const double* dk_host,
                               const sqb_c128* vtable, double mass, double hbar, int64_t block_begin,
                               int64_t block_count, sqb_c128* H_out, void* stream) {
  BuildParams P;
  int64_t nk;
  int rc = fill_params(P, nd, shape_host, repeat_host, dk_host, mass, hbar, &nk);
  if (rc) return rc;
  if (!vtable || !H_out || block_begin < 0 || block_count < 0 || block_begin + block_count > nk) return SQB_E_BADARG;
  if (block_count == 0) return SQB_OK;
  const size_t smem = (size_t)P.n * sizeof(c128);
  if (smem > 200 * 1024) return SQB_E_UNSUPPORTED;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(bloch_build_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
  }
  const int row_tiles = (P.n + kRowsPerCta - 1) / kRowsPerCta;
  if (row_tiles > 65535) return SQB_E_UNSUPPORTED;
  // block index on grid.x (2^31-1 limit), row tile on grid.y
  dim3 grid((unsigned)block_count, (unsigned)row_tiles);
  bloch_build_kernel<<<grid, kBuildThreads, smem, sqb_stream(stream)>>>(
      P, reinterpret_cast<const c128*>(vtable), block_begin, reinterpret_cast<c128*>(H_out));
  SQB_LAUNCH_CHECK();
  return SQB_OK;
}

extern "C" int sqb_bloch_kinetic(int nd, const int* shape_host, const int* repeat_host, const double* dk_host,
                                 double mass, double hbar, int64_t block_begin, int64_t block_count,
                                 sqb_c128* E_out, void* stream) {
  BuildParams P;
  int64_t nk;
  int rc = fill_params(P, nd, shape_host, repeat_host, dk_host, mass, hbar, &nk);
  if (rc) return rc;
  if (!E_out || block_begin < 0 || block_count < 0 || block_begin + block_count > nk) return SQB_E_BADARG;
  if (block_count == 0) return SQB_OK;
  const int64_t total = block_count * P.n;
  const int grid = (int)sqb_min64((total + 255) / 256, 148 * 8);
  bloch_kinetic_kernel<<<grid, 256, 0, sqb_stream(stream)>>>(P, block_begin, block_count,
                                                             reinterpret_cast<c128*>(E_out));
  SQB_LAUNCH_CHECK();
  return SQB_OK;
}

extern "C" int sqb_kinetic_energy(int nd, const int* shape_host, const double* dk_host, const double* fraction_host,
                                  double mass, double hbar, sqb_c128* E_out, void* stream) {
  BuildParams P;
  int64_t nk;
  const int ones[SQB_MAX_DIMS] = {1, 1, 1, 1};
  int rc = fill_params(P, nd, shape_host, ones, dk_host, mass, hbar, &nk);
  if (rc) return rc;
  if (!E_out) return SQB_E_BADARG;
  P.use_fraction = 1;
  for (int i = 0; i < nd; ++i) P.fraction[i] = fraction_host ? fraction_host[i] : 0.0;
  const int grid = (int)sqb_min64(((int64_t)P.n + 255) / 256, 148 * 8);
  bloch_kinetic_kernel<<<grid, 256, 0, sqb_stream(stream)>>>(P, 0, 1, reinterpret_cast<c128*>(E_out));
  SQB_LAUNCH_CHECK();
  return SQB_OK;
}

extern "C" int sqb_bloch_permute(int nd, const int* shape_host, const int* repeat_host, int direction,
                                 int64_t lead, const sqb_c128* in, sqb_c128* out, void* stream) {
  SqbDims d;
  int64_t n, nk;
  int rc = sqb_fill_dims(d, nd, shape_host, repeat_host, &n, &nk);
  if (rc) return rc;
  if (!in || !out || in == out || lead < 0 || (direction != 0 && direction != 1) || n > (1 << 20)) return SQB_E_BADARG;
  const int64_t m = n * nk;
  for (int a = 0; a < nd; ++a)
    if ((int64_t)d.shape[a] * d.repeat[a] > (1 << 30)) return SQB_E_UNSUPPORTED;
  if (lead == 0) return SQB_OK;
  const int64_t total = lead * m;
  const int grid = (int)sqb_min64((total + 255) / 256, 148 * 16);
  bloch_permute_kernel<<<grid, 256, 0, sqb_stream(stream)>>>(d, (int)n, m, lead, direction,
                                                             reinterpret_cast<const c128*>(in),
                                                             reinterpret_cast<c128*>(out));
  SQB_LAUNCH_CHECK();
  return SQB_OK;
}

extern "C" int sqb_bloch_time_reversal_mirror(int nd, const int* shape_host, const int* repeat_host, int64_t first,
                                              int64_t count, sqb_c128* transform, double* w, void* stream) {
  SqbDims d;
  int64_t n, nk;
  int rc = sqb_fill_dims(d, nd, shape_host, repeat_host, &n, &nk);
  if (rc) return rc;
  if (!transform || !w || first < 0 || count < 0 || first + count > nk || n > (1 << 20)) return SQB_E_BADARG;
  if (count == 0) return SQB_OK;
  for (int64_t b0 = 0; b0 < count; b0 += 65535) {
    const int64_t cnt = sqb_min64(65535, count - b0);
    time_reversal_mirror_kernel<<<dim3((unsigned)((n + 7) / 8), (unsigned)cnt), 256, 0, sqb_stream(stream)>>>(
        d, (int)n, first + b0, cnt, reinterpret_cast<c128*>(transform), w);
    SQB_LAUNCH_CHECK();
  }
  return SQB_OK;
}
