// K4: complex128 FFT, norm="ortho", shared-memory mixed-radix Stockham, plus the Bloch "cell FFT".
//
// One CTA transforms `lines_per_cta` 1-D lines that it first stages into shared memory (HBM traffic:
// one read + one write of the array per axis pass: 32 B per element per pass).  Lines longer than
// kMaxSmemLen are split L = L1*L2 (four-step: strided L1-point pass with the inter-stage twiddle fused
// into its store, then contiguous L2-point pass with a transposed store).
//
// Line addressing is generic: a line is identified by (outer index o, inner index i); both are decomposed
// into up to 4 mixed-radix digits with independent element strides on the input and on the output side.
// That lets the LAST pass of the Bloch cell FFT scatter straight into supercell position order
// (x_a = c_a*N_a + j_a) without a separate transpose pass.
//
// Cell FFT (sqb_cell_fft): with supercell momentum s_a = k_a*R_a + q_a and position x_a = c_a*N_a + j_a,
//   exp(2 pi i s_a x_a / L_a) = exp(2 pi i k_a j_a / N_a) * exp(2 pi i q_a j_a / L_a) * exp(2 pi i q_a c_a / R_a)
// so the supercell inverse FFT of Bloch-ordered data factorises into an N-point FFT inside every block, a
// twiddle, and an R-point FFT ACROSS blocks.  The first two are folded into the eigenvectors once
// (sqb_fold_transform), which leaves only the R-point passes per evolved state: they read/write
// N-element contiguous runs instead of gathering 16 B elements (replaces K5 + the big K4 on the way out).
//
// Convention (reference): position -> momentum on a ket index = np.fft.fft(norm="ortho"),
// momentum -> position = np.fft.ifft(norm="ortho")  (tests/test_operator.py:197-204).
#include <math.h>
#include <stdlib.h>

#include "sqb_common.cuh"

namespace {

constexpr int kFftThreads = 256;
constexpr int kFftThreadsWide = 512;  // in-place variant: one butterfly per thread and stage
constexpr int kMaxSmemLen = 4096;   // complex elements of one ping-pong buffer
constexpr int kMaxStages = 24;
constexpr int kMaxRadix = 31;
constexpr int kMaxLines = 64;       // lines per CTA (always a power of two)
constexpr int kDigits = 6;

struct Digits {  // flat index (C order over dims, last fastest) -> sum digit * stride
  int nd;
  int dims[kDigits];
  int64_t strides[kDigits];
};

__host__ __device__ inline int64_t digits_offset(const Digits& d, int64_t flat) {
  int64_t off = 0;
  if (flat < (1LL << 31)) {  // 32-bit divisions: ~5x cheaper than the 64-bit sequence on the device
    uint32_t f = (uint32_t)flat;
    for (int a = d.nd - 1; a >= 0; --a) {
      const uint32_t dim = (uint32_t)d.dims[a];
      const uint32_t q = f / dim;
      off += (int64_t)(f - q * dim) * d.strides[a];
      f = q;
    }
    return off;
  }
  for (int a = d.nd - 1; a >= 0; --a) {
    off += (flat % d.dims[a]) * d.strides[a];
    flat /= d.dims[a];
  }
  return off;
}

struct FftPass {
  int len;                 // transform length of this pass
  int n_stages;
  int radix[kMaxStages];
  int64_t n_outer, n_inner;  // lines = n_outer * n_inner, line id = o * n_inner + i
  Digits in_outer, in_inner, out_outer, out_inner;
  int64_t in_sl, out_sl;   // element stride of the position inside a line
  // optional two-level position stride on the input: pos -> (pos / in_split) * in_sl_hi + (pos % in_split) * in_sl
  // (a line that crosses rank-major chunks of an all-to-all receive buffer); in_split == 0: plain pos * in_sl
  int in_split;
  int64_t in_sl_hi;
  int out_split;
  int64_t out_sl_hi;
  int lines_per_cta;
  int inner_fastest;       // load order: 1 = a CTA's lines are adjacent in memory (coalesce across lines)
  int out_inner_fastest;   // same for the store
  int sign;                // -1 forward, +1 backward
  double scale;            // multiplied into every output
  int64_t tw_len;          // 0 = none; else multiply output (pos k, inner i) by exp(sign*2*pi*i*k*(i/tw_div)/tw_len)
  int64_t tw_div;          // divisor applied to the inner index before the twiddle (>= 1)
  int inplace_mid;         // 1 = the middle stages run in place (one shared buffer; needs work <= threads, see launch_pass)
  // K6 epilogue (kMom kernels only; mom_nd == 0: off).  Instead of being stored, every output element v at output
  // offset e adds |v|^2 {1, x_a, x_a^2, cos_a, sin_a} for every axis a of the position grid into per-CTA partial
  // sums; x_a = idx_a * spacing_a + offset_a(state), folded into [-length_a / 2, length_a / 2) when wrapped.
  int mom_nd;
  int mom_wrapped;
  int mom_grid[SQB_MAX_DIMS];
  double mom_spacing[SQB_MAX_DIMS], mom_length[SQB_MAX_DIMS];
  int64_t mom_m;                 // elements per state (all lines of a CTA belong to one state)
  int64_t mom_lead;              // number of states
  const double* mom_offsets;     // [nd][lead] or nullptr
  const c128* mom_cs;            // [sum_a L_a] (cos, sin)(2 pi idx / L_a), axis a at mom_cs_off[a]
  int mom_cs_off[SQB_MAX_DIMS];
  double* mom_partials;          // [n_ctas][1 + 4 nd]
  int mom_groups;                // consecutive line groups (of lines_per_cta lines) a CTA walks; all in one state
};

constexpr int kMomMax = 1 + 4 * SQB_MAX_DIMS;
struct MomLines {  // per CTA (shared): what is constant along a line of the LAST axis
  double x[SQB_MAX_DIMS][kMaxLines];   // leading axes a < nd-1: (wrapped) position of the line
  double c[SQB_MAX_DIMS][kMaxLines];   //                      : cos / sin of its scattering phase
  double s[SQB_MAX_DIMS][kMaxLines];
  int j_last[kMaxLines];               // index along the last axis of the line's first output (pos = 0)
};
struct MomAcc {  // per thread; every array index below is a compile-time constant (registers, not local memory)
  double w;                           // sum |v|^2
  double lead[SQB_MAX_DIMS - 1][4];   // leading axes: {w x, w x^2, w cos, w sin}
  double last[4];                     // last axis
  double off_last, inv_len_last, spacing_last, length_last, dx_last;
  const MomLines* lines;
};

// One output element of a line.  Only the LAST axis varies along a line: idx = j_last + pos * N_last, so
//   x = x0 + pos * dx          (x0 = j_last * spacing + offset, dx = N_last * spacing; folded when wrapped)
//   exp(2 pi i idx / L) = exp(2 pi i j_last / L) * conj(tab[pos])   (tab = the pass' own twiddle table, len = R)
// -> per element: |v|^2, x, two FMAs for x / x^2 and two for sum w conj-phase; the line's own phase is applied
// once per butterfly in mom_add_line.  No global-memory access in the epilogue.
__device__ __forceinline__ double mom_add_last(const FftPass& P, MomAcc& acc, double x0, int pos, c128 v, c128 tw,
                                               double& sc_x, double& sc_y) {
  const double w = fma(v.x, v.x, v.y * v.y);
  double x = fma((double)pos, acc.dx_last, x0);
  if (P.mom_wrapped) x -= acc.length_last * floor(fma(x, acc.inv_len_last, 0.5));
  acc.last[0] = fma(w, x, acc.last[0]);
  acc.last[1] = fma(w * x, x, acc.last[1]);
  sc_x = fma(w, tw.x, sc_x);
  sc_y = fma(w, tw.y, sc_y);
  return w;
}
// the weight `wsum` of one or more elements of a line: norm and the leading axes (constant along the line)
__device__ __forceinline__ void mom_add_line(const FftPass& P, MomAcc& acc, int line, double wsum, double sc_x,
                                             double sc_y) {
  acc.w += wsum;
  {
    // sum_pos w exp(+2 pi i pos / R) = conj(sum_pos w tab[pos]) = (sc_x, -sc_y); times the line's phase (cj, sj)
    const int al = P.mom_nd - 1;
    const double cj = acc.lines->c[al][line], sj = acc.lines->s[al][line];
    acc.last[2] += cj * sc_x + sj * sc_y;
    acc.last[3] += sj * sc_x - cj * sc_y;
  }
#pragma unroll
  for (int a = 0; a < SQB_MAX_DIMS - 1; ++a) {
    if (a < P.mom_nd - 1) {
      const double x = acc.lines->x[a][line];
      acc.lead[a][0] = fma(wsum, x, acc.lead[a][0]);
      acc.lead[a][1] = fma(wsum * x, x, acc.lead[a][1]);
      acc.lead[a][2] = fma(wsum, acc.lines->c[a][line], acc.lead[a][2]);
      acc.lead[a][3] = fma(wsum, acc.lines->s[a][line], acc.lead[a][3]);
    }
  }
}

__device__ __forceinline__ c128 tw_lookup(const c128* tab, int idx, int sign) {
  c128 w = tab[idx];
  if (sign > 0) w.y = -w.y;  // table holds exp(-2 pi i j / L)
  return w;
}

// Fully unrolled small DFT for odd primes (registers only): roots[j] = tab[j*root_stride] = exp(-2 pi i j/R).
template <int R>
__device__ __forceinline__ void dft_small(c128* x, int sign, const c128* tab, int root_stride) {
  c128 w[R];
#pragma unroll
  for (int j = 1; j < R; ++j) w[j] = tw_lookup(tab, j * root_stride, sign);
  c128 y[R];
  {
    c128 acc = x[0];
#pragma unroll
    for (int t = 1; t < R; ++t) acc = cadd(acc, x[t]);
    y[0] = acc;
  }
#pragma unroll
  for (int k = 1; k < R; ++k) {
    c128 acc = x[0];
#pragma unroll
    for (int t = 1; t < R; ++t) cfma(acc, x[t], w[(t * k) % R == 0 ? 1 : (t * k) % R]);
    y[k] = acc;
  }
#pragma unroll
  for (int k = 0; k < R; ++k) x[k] = y[k];
}

template <>
__device__ __forceinline__ void dft_small<2>(c128* x, int, const c128*, int) {
  c128 a = x[0], b = x[1];
  x[0] = cadd(a, b);
  x[1] = csub(a, b);
}

// multiply by exp(sign * i*pi/2) = sign*i
__device__ __forceinline__ c128 mul_i_sign(c128 a, int sign) {
  return sign > 0 ? make_double2(-a.y, a.x) : make_double2(a.y, -a.x);
}

template <>
__device__ __forceinline__ void dft_small<4>(c128* x, int sign, const c128*, int) {
  c128 a = cadd(x[0], x[2]), b = csub(x[0], x[2]);
  c128 c = cadd(x[1], x[3]), d = mul_i_sign(csub(x[1], x[3]), sign);
  x[0] = cadd(a, c);
  x[2] = csub(a, c);
  x[1] = cadd(b, d);
  x[3] = csub(b, d);
}

template <>
__device__ __forceinline__ void dft_small<8>(c128* x, int sign, const c128*, int) {
  const double h = 0.70710678118654752440;
  c128 a[4], b[4];
#pragma unroll
  for (int m = 0; m < 4; ++m) {
    a[m] = cadd(x[m], x[m + 4]);
    b[m] = csub(x[m], x[m + 4]);
  }
  // b[m] *= exp(sign * 2 pi i m / 8)
  {
    const c128 t1 = b[1], t3 = b[3];
    if (sign > 0) {
      b[1] = make_double2(h * (t1.x - t1.y), h * (t1.x + t1.y));    // * (1+i)/sqrt2
      b[3] = make_double2(-h * (t3.x + t3.y), h * (t3.x - t3.y));   // * (-1+i)/sqrt2
    } else {
      b[1] = make_double2(h * (t1.x + t1.y), h * (t1.y - t1.x));    // * (1-i)/sqrt2
      b[3] = make_double2(h * (t3.y - t3.x), -h * (t3.x + t3.y));   // * (-1-i)/sqrt2
    }
    b[2] = mul_i_sign(b[2], sign);
  }
  dft_small<4>(a, sign, nullptr, 0);
  dft_small<4>(b, sign, nullptr, 0);
#pragma unroll
  for (int m = 0; m < 4; ++m) {
    x[2 * m] = a[m];
    x[2 * m + 1] = b[m];
  }
}

// 16 = 4 x 4 Cooley-Tukey in registers (t = 4 t1 + t2, k = k1 + 4 k2): four 4-point DFTs over t1, the twiddles
// W16^(t2 k1), four 4-point DFTs over t2.  Lets 128- and 256-point lines run in TWO Stockham stages (one shared
// round trip) instead of three.
template <>
__device__ __forceinline__ void dft_small<16>(c128* x, int sign, const c128*, int) {
  const double c1 = 0.92387953251128674, s1 = 0.38268343236508977, h = 0.70710678118654752440;
  const double sg = sign > 0 ? 1.0 : -1.0;
  c128 y[4][4];  // y[t2][k1]
#pragma unroll
  for (int t2 = 0; t2 < 4; ++t2) {
    c128 v[4] = {x[t2], x[4 + t2], x[8 + t2], x[12 + t2]};
    dft_small<4>(v, sign, nullptr, 0);
#pragma unroll
    for (int k1 = 0; k1 < 4; ++k1) y[t2][k1] = v[k1];
  }
  const c128 w1 = make_double2(c1, sg * s1), w2 = make_double2(h, sg * h), w3 = make_double2(s1, sg * c1);
  const c128 w6 = make_double2(-h, sg * h), w9 = make_double2(-c1, -sg * s1);
  y[1][1] = cmul(y[1][1], w1);
  y[1][2] = cmul(y[1][2], w2);
  y[1][3] = cmul(y[1][3], w3);
  y[2][1] = cmul(y[2][1], w2);
  y[2][2] = mul_i_sign(y[2][2], sign);
  y[2][3] = cmul(y[2][3], w6);
  y[3][1] = cmul(y[3][1], w3);
  y[3][2] = cmul(y[3][2], w6);
  y[3][3] = cmul(y[3][3], w9);
#pragma unroll
  for (int k1 = 0; k1 < 4; ++k1) {
    c128 v[4] = {y[0][k1], y[1][k1], y[2][k1], y[3][k1]};
    dft_small<4>(v, sign, nullptr, 0);
#pragma unroll
    for (int k2 = 0; k2 < 4; ++k2) x[k1 + 4 * k2] = v[k2];
  }
}

// Generic small DFT of any radix r using the root table: roots[j] = exp(-2 pi i j / r).
__device__ __forceinline__ void dft_generic(c128* x, int r, const c128* tab, int tab_stride, int sign) {
  c128 y[kMaxRadix];
  for (int k = 0; k < r; ++k) {
    c128 acc = x[0];
    for (int t = 1; t < r; ++t) {
      c128 w = tw_lookup(tab, ((t * k) % r) * tab_stride, sign);
      cfma(acc, x[t], w);
    }
    y[k] = acc;
  }
  for (int k = 0; k < r; ++k) x[k] = y[k];
}

// table: exp(-2 pi i j / len), j in [0, len)
__global__ void fft_table_kernel(c128* tab, int len) {
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < len; j += gridDim.x * blockDim.x) {
    double s, c;
    sincospi(-2.0 * (double)j / (double)len, &s, &c);
    tab[j] = make_double2(c, s);
  }
}

// (w / d, w % d) with a shift fast path when d is a power of two (shift >= 0)
__device__ __forceinline__ void fast_divmod(int w, int d, int shift, int& q, int& r) {
  if (shift >= 0) {
    q = w >> shift;
    r = w & (d - 1);
  } else {
    q = w / d;
    r = w - q * d;
  }
}
__device__ __forceinline__ int pow2_shift(int d) { return (d & (d - 1)) == 0 ? 31 - __clz(d) : -1; }

__device__ __forceinline__ int64_t pos_offset(int pos, int split, int64_t sl_hi, int64_t sl) {
  if (split == 0) return (int64_t)pos * sl;
  const int hi = pos / split;
  return (int64_t)hi * sl_hi + (int64_t)(pos - hi * split) * sl;
}

struct StageIo {
  const c128* src;          // shared source (unused by the first stage)
  c128* dst;                // shared destination (unused by the last stage)
  const c128* gin;          // global input / output
  c128* gout;
  const int64_t* in_base;   // per-line element offsets (shared)
  const int64_t* out_base;
  const int64_t* inner_idx;
  bool first, last, line_fastest;
  bool line_major;          // shared layout: true = [pos][line] (line fastest), false = [line][pos]
};

// block reduction of the K6 accumulators -> mom_partials[blockIdx.x][1 + 4 nd]  (every thread of the CTA calls it)
template <int kThreads>
__device__ __forceinline__ void mom_flush(const FftPass& P, const MomAcc& acc) {
  __shared__ double s_red[kMomMax][kThreads / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int k_count = 1 + 4 * P.mom_nd;
  const int al = P.mom_nd - 1;
  {
    const double v = warp_sum(acc.w);
    if (lane == 0) s_red[0][warp] = v;
  }
#pragma unroll
  for (int a = 0; a < SQB_MAX_DIMS; ++a) {
    if (a <= al) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        // axis a: the last axis' sums live in acc.last, the leading ones in acc.lead[a]
        double mine = acc.last[j];
        if (a < SQB_MAX_DIMS - 1 && a < al) mine = acc.lead[a < SQB_MAX_DIMS - 1 ? a : 0][j];
        const double v = warp_sum(mine);
        if (lane == 0) s_red[1 + 4 * a + j][warp] = v;
      }
    }
  }
  __syncthreads();
  if (threadIdx.x < k_count) {
    double s = 0.0;
#pragma unroll
    for (int w = 0; w < kThreads / 32; ++w) s += s_red[threadIdx.x][w];
    P.mom_partials[(int64_t)blockIdx.x * k_count + threadIdx.x] = s;
  }
}

// One Stockham stage of radix R over `lines` lines held by the CTA.  The first stage reads its butterfly
// inputs straight from global memory and the last one writes straight to global memory, so an FFT of S
// stages makes S-1 round trips through shared memory instead of S+1.
template <int R, int kThreads, bool kMom = false>
__device__ __forceinline__ void stage_fixed(const FftPass& P, const StageIo& io, int ns, int lines, const c128* tab,
                                            MomAcc* acc = nullptr) {
  const int len = P.len, sign = P.sign, lpc = P.lines_per_cta;
  const int per_line = len / R;
  const int tw_step = len / (ns * R);
  const int pl_shift = pow2_shift(per_line), ns_shift = pow2_shift(ns);
  const int lshift = 31 - __clz(lpc);
  const int work = per_line * (io.line_fastest ? lpc : lines);
  // shared-memory index of (line, pos): with the [pos][line] layout a warp that walks 32 adjacent lines at one
  // position touches 32 consecutive elements (conflict free); [line][pos] would be a 32-way bank conflict there
  const int s_line = io.line_major ? 1 : len;
  const int s_pos = io.line_major ? lpc : 1;
  for (int w = threadIdx.x; w < work; w += kThreads) {
    int line, j, jq, k;
    if (io.line_fastest) {
      line = w & (lpc - 1);
      j = w >> lshift;
      if (line >= lines) continue;
    } else {
      fast_divmod(w, per_line, pl_shift, line, j);
    }
    fast_divmod(j, ns, ns_shift, jq, k);
    c128 x[R];
    if (io.first) {
      const int64_t base = io.in_base[line];
#pragma unroll
      for (int t = 0; t < R; ++t) x[t] = io.gin[base + pos_offset(j + t * per_line, P.in_split, P.in_sl_hi, P.in_sl)];
    } else {
#pragma unroll
      for (int t = 0; t < R; ++t) x[t] = io.src[line * s_line + (j + t * per_line) * s_pos];
    }
    if (ns > 1) {
#pragma unroll
      for (int t = 1; t < R; ++t) x[t] = cmul(x[t], tw_lookup(tab, t * k * tw_step, sign));
    }
    dft_small<R>(x, sign, tab, len / R);
    const int pos0 = (j - k) * R + k;
    if (io.last) {
      const int64_t base = io.out_base[line];
      double wsum = 0.0, x0 = 0.0, sc_x = 0.0, sc_y = 0.0;
      if constexpr (kMom) x0 = acc->lines->x[P.mom_nd - 1][line];
      (void)wsum;
      (void)base;
      (void)x0;
#pragma unroll
      for (int t = 0; t < R; ++t) {
        const int pos = pos0 + t * ns;
        c128 v = cscale(x[t], P.scale);
        if (P.tw_len) {
          double sn, cs;
          const int64_t prod = ((int64_t)pos * (io.inner_idx[line] / P.tw_div)) % P.tw_len;
          sincospi((double)sign * 2.0 * (double)prod / (double)P.tw_len, &sn, &cs);
          v = cmul(v, make_double2(cs, sn));
        }
        if constexpr (kMom) wsum += mom_add_last(P, *acc, x0, pos, v, tab[pos], sc_x, sc_y);
        else io.gout[base + pos_offset(pos, P.out_split, P.out_sl_hi, P.out_sl)] = v;
      }
      if constexpr (kMom) mom_add_line(P, *acc, line, wsum, sc_x, sc_y);
    } else {
#pragma unroll
      for (int t = 0; t < R; ++t) io.dst[line * s_line + (pos0 + t * ns) * s_pos] = x[t];
    }
  }
}

// Middle stage IN PLACE on one shared buffer ([pos][line] layout): every thread owns at most one butterfly, reads
// its R inputs, the CTA synchronises, then the outputs go back into the same buffer.  Halves the shared memory
// of a three-stage pass, which doubles the adjacent lines a CTA can take (longer coalesced runs).
template <int R, int kThreads>
__device__ __forceinline__ void stage_fixed_inplace(const FftPass& P, c128* buf, int ns, int lines, const c128* tab) {
  const int len = P.len, sign = P.sign, lpc = P.lines_per_cta;
  const int per_line = len / R;
  const int tw_step = len / (ns * R);
  const int ns_shift = pow2_shift(ns);
  const int lshift = 31 - __clz(lpc);
  const int w = threadIdx.x;
  const int line = w & (lpc - 1);
  const int j = w >> lshift;
  const bool active = w < per_line * lpc && line < lines;
  int jq, k;
  fast_divmod(j, ns, ns_shift, jq, k);
  c128 x[R];
  if (active) {
#pragma unroll
    for (int t = 0; t < R; ++t) x[t] = buf[line + (j + t * per_line) * lpc];
    if (ns > 1) {
#pragma unroll
      for (int t = 1; t < R; ++t) x[t] = cmul(x[t], tw_lookup(tab, t * k * tw_step, sign));
    }
    dft_small<R>(x, sign, tab, len / R);
  }
  __syncthreads();
  if (active) {
    const int pos0 = (j - k) * R + k;
#pragma unroll
    for (int t = 0; t < R; ++t) buf[line + (pos0 + t * ns) * lpc] = x[t];
  }
}

// Any radix (primes 11..31): shared memory on both sides (the caller brackets it with copy stages).
__device__ __forceinline__ void stage_generic(const c128* src, c128* dst, int len, int ns, int r, int lines,
                                              const c128* tab, int sign) {
  const int per_line = len / r;
  const int work = per_line * lines;
  const int tw_step = len / (ns * r);
  const int root_stride = len / r;
  for (int w = threadIdx.x; w < work; w += kFftThreads) {
    const int line = w / per_line;
    const int j = w - line * per_line;
    const int k = j % ns;
    c128 x[kMaxRadix];
    for (int t = 0; t < r; ++t) {
      c128 v = src[line * len + j + t * per_line];
      x[t] = t ? cmul(v, tw_lookup(tab, t * k * tw_step, sign)) : v;
    }
    dft_generic(x, r, tab, root_stride, sign);
    const int base = line * len + (j - k) * r + k;
    for (int t = 0; t < r; ++t) dst[base + t * ns] = x[t];
  }
}

__device__ __forceinline__ bool radix_is_fixed(int r) { return r == 2 || r == 3 || r == 4 || r == 5 || r == 7 || r == 8 || r == 16; }

// kGeneric = false: every radix is one of {2,3,4,5,7,8} (the common case; no local-memory radix arrays are
// compiled in, which keeps the register count down for 4 CTAs / SM)
// kVariant 0: radices {2,4,8} only (power-of-two lengths, the hot cell-FFT passes); 1: {2,3,4,5,7,8};
// 2: any radix <= 31 (local-memory arrays); 3: {2,4,8,16} (power-of-two lengths of 128 points and more: the radix-16
// butterfly needs ~128 registers, two CTAs / SM).  Fewer compiled-in radices = fewer registers = more CTAs / SM.
// One group of `lines` lines (first line `line0`) through all stages of the pass.
template <int kVariant, int kThreads, bool kMom>
__device__ __forceinline__ void fft_pass_group(const FftPass& P, const c128* __restrict__ in, c128* __restrict__ out,
                                               const c128* tab, c128* buf0, c128* buf1, int64_t* s_in_base,
                                               int64_t* s_out_base, int64_t* s_inner_idx, MomLines& s_mom, MomAcc& acc,
                                               int64_t line0, int lines) {
  const int len = P.len;
  const int lpc = P.lines_per_cta;
  if (threadIdx.x < lines) {
    const int64_t line = line0 + threadIdx.x;
    int64_t o, i;
    if (line < (1LL << 31) && P.n_inner < (1LL << 31)) {
      const uint32_t q = (uint32_t)line / (uint32_t)P.n_inner;
      o = q;
      i = (uint32_t)line - q * (uint32_t)P.n_inner;
    } else {
      o = line / P.n_inner;
      i = line - o * P.n_inner;
    }
    s_in_base[threadIdx.x] = digits_offset(P.in_outer, o) + digits_offset(P.in_inner, i);
    s_out_base[threadIdx.x] = digits_offset(P.out_outer, o) + digits_offset(P.out_inner, i);
    s_inner_idx[threadIdx.x] = i;
  }
  __syncthreads();
  if constexpr (kMom) {
    // launch_pass guarantees one state per CTA: state = first line / lines per state (32-bit when it fits)
    const int64_t lps = (P.n_outer * P.n_inner) / P.mom_lead;
    const int64_t state = (line0 < (1LL << 31) && lps < (1LL << 31)) ? (int64_t)((uint32_t)line0 / (uint32_t)lps) : line0 / lps;
    const int al = P.mom_nd - 1;
    acc.off_last = P.mom_offsets ? P.mom_offsets[(int64_t)al * P.mom_lead + state] : 0.0;
    if (threadIdx.x < lines) {
      // decode the line's first output offset into the position-grid indices (C order, last axis fastest)
      uint32_t m = (uint32_t)(s_out_base[threadIdx.x] - state * P.mom_m);
      for (int a = P.mom_nd - 1; a >= 0; --a) {
        const uint32_t L = (uint32_t)P.mom_grid[a];
        const uint32_t q = m / L;
        const int idx = (int)(m - q * L);
        m = q;
        if (a == al) {
          s_mom.j_last[threadIdx.x] = idx;
          const c128 cs = P.mom_cs[P.mom_cs_off[a] + idx];
          s_mom.x[a][threadIdx.x] = fma((double)idx, P.mom_spacing[a], acc.off_last);  // x0 of the line
          s_mom.c[a][threadIdx.x] = cs.x;
          s_mom.s[a][threadIdx.x] = cs.y;
        } else {
          const double off = P.mom_offsets ? P.mom_offsets[(int64_t)a * P.mom_lead + state] : 0.0;
          double x = fma((double)idx, P.mom_spacing[a], off);
          if (P.mom_wrapped) x -= P.mom_length[a] * floor(x / P.mom_length[a] + 0.5);
          const c128 cs = P.mom_cs[P.mom_cs_off[a] + idx];
          s_mom.x[a][threadIdx.x] = x;
          s_mom.c[a][threadIdx.x] = cs.x;
          s_mom.s[a][threadIdx.x] = cs.y;
        }
      }
    }
    __syncthreads();
  }

  const int lshift = 31 - __clz(lpc);
  // [pos][line] shared layout (and line-fastest work mapping in every stage) whenever the CTA's lines are
  // adjacent in memory; the generic-radix / copy paths below keep the plain [line][pos] layout
  bool all_fixed = P.n_stages > 0;
  for (int s = 0; s < P.n_stages; ++s) all_fixed = all_fixed && radix_is_fixed(P.radix[s]);
  const bool line_major = all_fixed && P.inner_fastest && P.out_inner_fastest;
  // A generic-radix stage at either end (or a length-1 transform) cannot touch global memory itself:
  // stage the data through shared memory with plain copies in that case.
  const bool copy_in = P.n_stages == 0 || !radix_is_fixed(P.radix[0]);
  const bool copy_out = P.n_stages == 0 || !radix_is_fixed(P.radix[P.n_stages - 1]);
  c128* src = buf0;
  c128* dst = buf1;
  if (copy_in) {
    if (P.inner_fastest) {
      const int l = threadIdx.x & (lpc - 1);
      if (l < lines)
        for (int pos = threadIdx.x >> lshift; pos < len; pos += kThreads >> lshift)
          buf0[l * len + pos] = in[s_in_base[l] + pos_offset(pos, P.in_split, P.in_sl_hi, P.in_sl)];
    } else {
      for (int e = threadIdx.x; e < lines * len; e += kThreads) {
        const int l = e / len, pos = e - l * len;
        buf0[l * len + pos] = in[s_in_base[l] + pos_offset(pos, P.in_split, P.in_sl_hi, P.in_sl)];
      }
    }
    __syncthreads();
  }

  int ns = 1;
  for (int s = 0; s < P.n_stages; ++s) {
    const int r = P.radix[s];
    StageIo io;
    io.first = (s == 0) && !copy_in;
    io.last = (s == P.n_stages - 1) && !copy_out;
    io.src = src;
    io.dst = io.first ? buf0 : dst;
    io.gin = in;
    io.gout = out;
    io.in_base = s_in_base;
    io.out_base = s_out_base;
    io.inner_idx = s_inner_idx;
    io.line_major = line_major;
    io.line_fastest = line_major || (io.first && P.inner_fastest) || (io.last && P.out_inner_fastest);
    const bool mid_inplace = P.inplace_mid && !io.first && !io.last;
    if (mid_inplace) {
      switch (r) {
        case 8: stage_fixed_inplace<8, kThreads>(P, src, ns, lines, tab); break;
        case 4: stage_fixed_inplace<4, kThreads>(P, src, ns, lines, tab); break;
        default: stage_fixed_inplace<2, kThreads>(P, src, ns, lines, tab); break;
      }
      __syncthreads();
      ns *= r;
      continue;
    }
    switch (r) {
      case 16: if constexpr (kVariant == 3) stage_fixed<16, kThreads, kMom>(P, io, ns, lines, tab, &acc); break;
      case 8: stage_fixed<8, kThreads, kMom>(P, io, ns, lines, tab, &acc); break;
      case 4: stage_fixed<4, kThreads, kMom>(P, io, ns, lines, tab, &acc); break;
      case 2: stage_fixed<2, kThreads, kMom>(P, io, ns, lines, tab, &acc); break;
      case 3: if constexpr (kVariant == 1 || kVariant == 2) stage_fixed<3, kThreads, kMom>(P, io, ns, lines, tab, &acc); break;
      case 5: if constexpr (kVariant == 1 || kVariant == 2) stage_fixed<5, kThreads, kMom>(P, io, ns, lines, tab, &acc); break;
      case 7: if constexpr (kVariant == 1 || kVariant == 2) stage_fixed<7, kThreads, kMom>(P, io, ns, lines, tab, &acc); break;
      default:
        if constexpr (kVariant == 2) stage_generic(src, dst, len, ns, r, lines, tab, P.sign);
        break;
    }
    if (io.last) return;
    __syncthreads();
    if (io.first) {
      src = buf0;  // first stage wrote buf0
      dst = buf1;
    } else {
      c128* t = src; src = dst; dst = t;
    }
    ns *= r;
  }

  // ---- copy out (generic last stage / length 1)
  auto emit = [&](int l, int pos) {
    c128 v = cscale(src[l * len + pos], P.scale);
    if (P.tw_len) {
      double sn, cs;
      const int64_t prod = ((int64_t)pos * (s_inner_idx[l] / P.tw_div)) % P.tw_len;
      sincospi((double)P.sign * 2.0 * (double)prod / (double)P.tw_len, &sn, &cs);
      v = cmul(v, make_double2(cs, sn));
    }
    if constexpr (kMom) {
      double sc_x = 0.0, sc_y = 0.0;
      const double w = mom_add_last(P, acc, s_mom.x[P.mom_nd - 1][l], pos, v, tab[pos], sc_x, sc_y);
      mom_add_line(P, acc, l, w, sc_x, sc_y);
    }
    else out[s_out_base[l] + pos_offset(pos, P.out_split, P.out_sl_hi, P.out_sl)] = v;
  };
  if (P.out_inner_fastest) {
    const int l = threadIdx.x & (lpc - 1);
    if (l < lines)
      for (int pos = threadIdx.x >> lshift; pos < len; pos += kThreads >> lshift) emit(l, pos);
  } else {
    for (int e = threadIdx.x; e < lines * len; e += kThreads) {
      const int l = e / len;
      emit(l, e - l * len);
    }
  }
}


template <int kVariant, int kThreads, bool kMom = false>
__global__ void __launch_bounds__(kThreads, kVariant == 3 ? 2 : (kMom ? 3 : (kThreads == kFftThreadsWide ? 2 : (kVariant == 0 ? 4 : (kVariant == 1 ? 3 : 2)))))
fft_pass_kernel(FftPass P, const c128* __restrict__ in, c128* __restrict__ out, const c128* __restrict__ tab_g) {
  extern __shared__ c128 smem[];
  __shared__ int64_t s_in_base[kMaxLines], s_out_base[kMaxLines], s_inner_idx[kMaxLines];
  const int len = P.len;
  const int lpc = P.lines_per_cta;
  c128* tab = smem;                       // [len]
  c128* buf0 = smem + len;                // [lpc][len]
  c128* buf1 = buf0 + (size_t)lpc * len;  // [lpc][len]  (only when the pass needs it, see launch_pass)

  for (int j = threadIdx.x; j < len; j += kThreads) tab[j] = tab_g[j];
  const int64_t n_lines = P.n_outer * P.n_inner;
  MomAcc acc;
  __shared__ MomLines s_mom;
  if constexpr (kMom) {
    acc.w = 0.0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      acc.last[j] = 0.0;
#pragma unroll
      for (int a = 0; a < SQB_MAX_DIMS - 1; ++a) acc.lead[a][j] = 0.0;
    }
    acc.lines = &s_mom;
    const int al = P.mom_nd - 1;
    acc.length_last = P.mom_length[al];
    acc.spacing_last = P.mom_spacing[al];
    acc.inv_len_last = 1.0 / acc.length_last;
    acc.dx_last = (double)P.out_sl * acc.spacing_last;
  }
  // kMom: a CTA walks `mom_groups` consecutive groups of lines of ONE state, keeps the observable sums in registers
  // and reduces them once (the block reduction would otherwise cost more than the 8 outputs a thread produces)
  const int groups = kMom ? P.mom_groups : 1;
  for (int g = 0; g < groups; ++g) {
    const int64_t line0 = ((int64_t)blockIdx.x * groups + g) * lpc;
    const int lines = (int)sqb_min64(lpc, n_lines - line0);
    if (lines <= 0) break;
    if (g > 0) __syncthreads();  // the previous group's shared buffers are free
    fft_pass_group<kVariant, kThreads, kMom>(P, in, out, tab, buf0, buf1, s_in_base, s_out_base, s_inner_idx, s_mom,
                                             acc, line0, lines);
  }
  if constexpr (kMom) mom_flush<kThreads>(P, acc);
}

// per-state sums of the per-CTA partials (fixed order: deterministic); out [lead][nd][5] = {w, wx, wx^2, wcos, wsin}
__global__ void mom_reduce_kernel(const double* __restrict__ partials, int ctas_per_state, int nd,
                                  double* __restrict__ out) {
  const int64_t state = blockIdx.x;
  const int k_count = 1 + 4 * nd;
  const int k = threadIdx.x;
  if (k >= k_count) return;
  const double* p = partials + state * (int64_t)ctas_per_state * k_count + k;
  double s = 0.0;
  for (int c = 0; c < ctas_per_state; ++c) s += p[(int64_t)c * k_count];
  if (k == 0) {
    for (int a = 0; a < nd; ++a) out[(state * nd + a) * 5] = s;
  } else {
    const int a = (k - 1) / 4, j = (k - 1) % 4;
    out[(state * nd + a) * 5 + 1 + j] = s;
  }
}

// cos / sin tables of the scattering phase, axis a at offset off[a]
__global__ void mom_table_kernel(c128* tab, int len) {
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < len; j += gridDim.x * blockDim.x) {
    double sn, cs;
    sincospi(2.0 * (double)j / (double)len, &sn, &cs);
    tab[j] = make_double2(cs, sn);
  }
}

// Bloch twiddle folded into the (already in-cell transformed) eigenvectors:
//   W[b, e, j] *= prod_a exp(sign * 2 pi i q_a(b) j_a / L_a),  q_a = FFT-order signed block index
// The twiddle depends on (block, j) only: a CTA computes the n factors of its block once into shared memory
// (one sincospi per j instead of per element) and streams kFoldRows eigenvector rows through them.
constexpr int kFoldRows = 32;
// Rows of one block are `rows` vectors of n entries, row r of block bl at W[bl * block_stride + r * row_stride]:
//   eigenvector blocks [batch, n, n]:           rows = n,    row_stride = n,       block_stride = n * n
//   state lists in block order [lead, n_k, n]:  rows = lead, row_stride = n_k * n, block_stride = n
__global__ void __launch_bounds__(256)
fold_twiddle_kernel(SqbDims d, int n, int64_t block_begin, int64_t batch, int sign, int64_t rows,
                    int64_t row_stride, int64_t block_stride, const c128* __restrict__ in, c128* __restrict__ W) {
  extern __shared__ __align__(16) unsigned char fold_smem[];
  c128* tw = reinterpret_cast<c128*>(fold_smem);  // [n]
  const int64_t bl = blockIdx.y;                   // block inside the batch
  if (bl >= batch) return;
  for (int j = threadIdx.x; j < n; j += blockDim.x) {
    int64_t b = block_begin + bl;
    int jj = j;
    double phase = 0.0;  // in turns
    for (int a = d.nd - 1; a >= 0; --a) {
      const int ja = jj % d.shape[a];
      jj /= d.shape[a];
      const int ba = (int)(b % d.repeat[a]);
      b /= d.repeat[a];
      const int qa = ba < (d.repeat[a] + 1) / 2 ? ba : ba - d.repeat[a];
      const int64_t big = (int64_t)d.shape[a] * d.repeat[a];
      const int64_t prod = ((int64_t)qa * ja) % big;
      phase += (double)prod / (double)big;
    }
    double sn, cs;
    sincospi((double)sign * 2.0 * phase, &sn, &cs);
    tw[j] = make_double2(cs, sn);
  }
  __syncthreads();
  const int64_t r0 = (int64_t)blockIdx.x * kFoldRows;
  const int64_t r1 = sqb_min64(rows, r0 + kFoldRows);
  const c128* Ib = in + bl * block_stride;
  c128* Wb = W + bl * block_stride;
  for (int64_t r = r0; r < r1; ++r) {
    const c128* src = Ib + r * row_stride;
    c128* row = Wb + r * row_stride;
    for (int j = threadIdx.x; j < n; j += blockDim.x) row[j] = cmul(src[j], tw[j]);
  }
}

// Fused fold (sqb_fold_transform) for small cells whose axes are single fixed radices (16 x 16, 8 x 8, 7 x 7 x 7, ...):
// one pass over the eigenvectors instead of one FFT pass per axis plus the twiddle pass.  A CTA stages `rows`
// eigenvector rows of one block in shared memory (the last axis padded by one element when its length is even, so
// the register butterflies of that axis are bank-conflict free), transforms every axis in place - one thread per
// line, the whole line in registers - and the pass of axis 0 multiplies by the ortho scale and the block's Bloch
// twiddle and stores straight to global memory (runs of n / shape[0] contiguous elements).
__device__ __forceinline__ void fold_dft(c128* x, int len, int sign, const c128* roots) {
  switch (len) {
    case 16: dft_small<16>(x, sign, roots, 1); break;
    case 8: dft_small<8>(x, sign, roots, 1); break;
    case 7: dft_small<7>(x, sign, roots, 1); break;
    case 5: dft_small<5>(x, sign, roots, 1); break;
    case 4: dft_small<4>(x, sign, roots, 1); break;
    case 3: dft_small<3>(x, sign, roots, 1); break;
    case 2: dft_small<2>(x, sign, roots, 1); break;
    default: break;  // length 1
  }
}
__global__ void __launch_bounds__(256, 2)
fold_fused_kernel(SqbDims d, int n, int rows_per_cta, int64_t block_begin, const c128* __restrict__ in,
                  c128* __restrict__ out) {
  extern __shared__ __align__(16) unsigned char fold_smem[];
  const int n_last = d.shape[d.nd - 1];
  const int pad = (n_last & 1) ? 0 : 1;
  const int row_stride = (n / n_last) * (n_last + pad);
  c128* tw = reinterpret_cast<c128*>(fold_smem);  // [n] scale * Bloch twiddle
  c128* roots = tw + n;                            // [nd][16] exp(-2 pi i j / shape[a])
  c128* buf = roots + 16 * SQB_MAX_DIMS;           // [rows][row_stride]
  const int64_t bl = blockIdx.y;
  const int r0 = blockIdx.x * rows_per_cta;
  const int rows = min(rows_per_cta, n - r0);
  const double scale = rsqrt((double)n);
  for (int j = threadIdx.x; j < n; j += blockDim.x) {
    int64_t b = block_begin + bl;
    int jj = j;
    double phase = 0.0;  // in turns
    for (int a = d.nd - 1; a >= 0; --a) {
      const int ja = jj % d.shape[a];
      jj /= d.shape[a];
      const int ba = (int)(b % d.repeat[a]);
      b /= d.repeat[a];
      const int qa = ba < (d.repeat[a] + 1) / 2 ? ba : ba - d.repeat[a];
      const int64_t big = (int64_t)d.shape[a] * d.repeat[a];
      phase += (double)(((int64_t)qa * ja) % big) / (double)big;
    }
    double sn, cs;
    sincospi(2.0 * phase, &sn, &cs);
    tw[j] = make_double2(cs * scale, sn * scale);
  }
  if (threadIdx.x < 16 * d.nd) {
    const int a = threadIdx.x >> 4, j = threadIdx.x & 15;
    double sn, cs;
    sincospi(-2.0 * (double)j / (double)d.shape[a], &sn, &cs);
    roots[threadIdx.x] = make_double2(cs, sn);
  }
  const c128* src = in + (bl * n + r0) * (int64_t)n;
  c128* dst = out + (bl * n + r0) * (int64_t)n;
  const int n_shift = pow2_shift(n), nl_shift = pow2_shift(n_last);
  const int total = rows * n;
  for (int base = 0; base < total; base += 8 * 256) {  // 8 independent 16-byte loads per thread in flight
    c128 v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int idx = base + u * 256 + threadIdx.x;
      v[u] = idx < total ? src[idx] : make_double2(0.0, 0.0);
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int idx = base + u * 256 + threadIdx.x;
      if (idx < total) {
        int r, e, hi, lo;
        fast_divmod(idx, n, n_shift, r, e);
        fast_divmod(e, n_last, nl_shift, hi, lo);
        buf[r * row_stride + hi * (n_last + pad) + lo] = v[u];
      }
    }
  }
  __syncthreads();
  int s_a = 1;  // element stride of axis a inside a row (unpadded index)
  for (int a = d.nd - 1; a >= 0; --a) {
    const int len = d.shape[a];
    const int lines = n / len;
    const int ln_shift = pow2_shift(lines), sa_shift = pow2_shift(s_a);
    // padded offset of element e0 + t s_a: axis nd-1 walks inside a padded run, the others jump whole runs
    const int pstride = a == d.nd - 1 ? 1 : (s_a / n_last) * (n_last + pad);
    for (int w = threadIdx.x; w < rows * lines; w += blockDim.x) {
      int r, line, outer, inner, hi, lo;
      fast_divmod(w, lines, ln_shift, r, line);
      fast_divmod(line, s_a, sa_shift, outer, inner);
      const int e0 = outer * len * s_a + inner;
      fast_divmod(e0, n_last, nl_shift, hi, lo);
      c128* p0 = buf + r * row_stride + hi * (n_last + pad) + lo;
      c128 x[16];
#pragma unroll
      for (int t = 0; t < 16; ++t)
        if (t < len) x[t] = p0[t * pstride];
      fold_dft(x, len, +1, roots + 16 * a);
      if (a == 0) {
        c128* g0 = dst + (int64_t)r * n + e0;
#pragma unroll
        for (int t = 0; t < 16; ++t)
          if (t < len) g0[t * s_a] = cmul(x[t], tw[e0 + t * s_a]);
      } else {
#pragma unroll
        for (int t = 0; t < 16; ++t)
          if (t < len) p0[t * pstride] = x[t];
      }
    }
    s_a *= len;
    if (a > 0) __syncthreads();
  }
}

// ------------------------------------------------------------------------------------ 2-D cell FFT in one pass (clusters)
// Both across-block axes of a 64 x 64 block grid in ONE kernel: HBM traffic is one read + one write of the state
// list instead of two of each.  A thread-block CLUSTER of 8 CTAs owns one (state, group of 8 adjacent cell
// columns): 64 x 64 x 8 elements = 512 KB, 64 KB of shared memory per CTA.
//   phase 1  CTA r loads the block rows b_0 in [8r, 8r + 8) (128-byte runs) and transforms axis 1 (64 points, two
//            radix-8 stages; the second in place: a butterfly reads and writes the same eight positions);
//   exchange the first radix-8 stage of axis 0 reads its eight inputs b_0 = jj + 8 t straight out of the eight CTAs'
//            shared memories (DSMEM: input t lives in CTA t) into registers - the whole axis-0 working set of a
//            CTA (4096 elements) is 16 registers' worth per thread - bracketed by two cluster barriers (slabs
//            complete / slabs no longer read), after which the slab is overwritten with the stage's outputs;
//   phase 2  second radix-8 stage of axis 0 for the block columns c_1 in [8r, 8r + 8), scaled, scattered into the
//            supercell position layout x_a = c_a N_a + j_a in 128-byte runs.
#ifndef SQB_C2_CTAS
#define SQB_C2_CTAS 2
#endif
constexpr int kC2Cluster = 8, kC2Threads = 256;
constexpr size_t kC2Smem = (size_t)8 * 64 * 8 * sizeof(c128);

__device__ __forceinline__ void c2_cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ c128 c2_ld_remote(const c128* local, uint32_t rank) {
  uint32_t la = (uint32_t)__cvta_generic_to_shared(local), ra;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(la), "r"(rank));
  c128 v;
  asm volatile("ld.shared::cluster.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(ra) : "memory");
  return v;
}

__global__ void __cluster_dims__(kC2Cluster, 1, 1) __launch_bounds__(kC2Threads, SQB_C2_CTAS)
cell_fft2_cluster_kernel(int n, int n1, int64_t item0, const c128* __restrict__ in, c128* __restrict__ out,
                         const c128* __restrict__ tab_g, double scale) {
  extern __shared__ __align__(16) c128 c2_S[];  // [8 rows][64 positions][8 columns]
  __shared__ c128 tab[64];
  const int tid = threadIdx.x;
  uint32_t rank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  const int64_t item = item0 + (blockIdx.x >> 3);
  const int groups = n >> 3;  // groups of 8 adjacent cell columns
  const int64_t t = item / groups;
  const int jg = (int)(item - t * groups);
  if (tid < 64) tab[tid] = tab_g[tid];
  const int jl = tid & 7, a = (tid >> 3) & 7, hi = tid >> 6;
  const int64_t m = (int64_t)4096 * n;
  const c128* src = in + t * m + (8 * jg + jl);
  // ---- phase 1, stage 1: rows b_0l = hi, hi + 4; butterfly jj = a: positions b_1 = jj + 8 tt from global memory
  c128 x[2][8];
#pragma unroll
  for (int pass = 0; pass < 2; ++pass) {
    const int b0 = 8 * (int)rank + hi + 4 * pass;
#pragma unroll
    for (int tt = 0; tt < 8; ++tt) x[pass][tt] = src[(int64_t)(b0 * 64 + a + 8 * tt) * n];
  }
#pragma unroll
  for (int pass = 0; pass < 2; ++pass) {
    dft_small<8>(x[pass], +1, nullptr, 0);
    const int b0l = hi + 4 * pass;
#pragma unroll
    for (int tt = 0; tt < 8; ++tt) c2_S[(b0l * 64 + a * 8 + tt) * 8 + jl] = x[pass][tt];
  }
  __syncthreads();
  // ---- phase 1, stage 2 (in place): butterfly jj2 = a reads and writes positions jj2 + 8 tt
#pragma unroll
  for (int pass = 0; pass < 2; ++pass) {
    const int b0l = hi + 4 * pass;
    c128 v[8];
#pragma unroll
    for (int tt = 0; tt < 8; ++tt) {
      v[tt] = c2_S[(b0l * 64 + a + 8 * tt) * 8 + jl];
      if (tt) v[tt] = cmul(v[tt], tw_lookup(tab, tt * a, +1));
    }
    dft_small<8>(v, +1, nullptr, 0);
#pragma unroll
    for (int tt = 0; tt < 8; ++tt) c2_S[(b0l * 64 + a + 8 * tt) * 8 + jl] = v[tt];
  }
  c2_cluster_sync();  // every CTA's slab holds the axis-1 transform of its block rows
  // ---- axis 0, stage 1: column c_1 = 8 rank + a, butterfly jj = hi, hi + 4: input tt = row b_0 = jj + 8 tt = row jj
  //      of CTA tt
#pragma unroll
  for (int pass = 0; pass < 2; ++pass) {
    const int jj = hi + 4 * pass;
    const c128* local = c2_S + (jj * 64 + 8 * (int)rank + a) * 8 + jl;
#pragma unroll
    for (int tt = 0; tt < 8; ++tt) x[pass][tt] = c2_ld_remote(local, (uint32_t)tt);
  }
  c2_cluster_sync();  // nobody reads a slab any more: it becomes the axis-0 work buffer [c_1l][position][column]
#pragma unroll
  for (int pass = 0; pass < 2; ++pass) {
    dft_small<8>(x[pass], +1, nullptr, 0);
    const int jj = hi + 4 * pass;
#pragma unroll
    for (int tt = 0; tt < 8; ++tt) c2_S[(a * 64 + jj * 8 + tt) * 8 + jl] = x[pass][tt];
  }
  __syncthreads();
  // ---- axis 0, stage 2: butterfly jj2 = hi, hi + 4; outputs c_0 = jj2 + 8 tt scattered into the position layout
  const int j = 8 * jg + jl;
  const int j0 = j / n1, j1 = j - j0 * n1;
  const int n0 = n / n1;
  const int64_t l1 = (int64_t)64 * n1;
  c128* dst = out + t * m + (int64_t)j0 * l1 + (int64_t)(8 * (int)rank + a) * n1 + j1;
#pragma unroll
  for (int pass = 0; pass < 2; ++pass) {
    const int jj2 = hi + 4 * pass;
    c128 v[8];
#pragma unroll
    for (int tt = 0; tt < 8; ++tt) {
      v[tt] = c2_S[(a * 64 + jj2 + 8 * tt) * 8 + jl];
      if (tt) v[tt] = cmul(v[tt], tw_lookup(tab, tt * jj2, +1));
    }
    dft_small<8>(v, +1, nullptr, 0);
#pragma unroll
    for (int tt = 0; tt < 8; ++tt) dst[(int64_t)(jj2 + 8 * tt) * n0 * l1] = cscale(v[tt], scale);
  }
}

// Push variant: the second axis-1 stage sends every output straight into the shared memory of the CTA that owns
// its block column (remote stores are fire and forget; the pull variant stalls on 16 remote loads per thread), so a
// single cluster barrier separates the two axes.  Shared memory: a 32 KB scratch for axis 1 (two halves of four
// block rows) + the 64 KB axis-0 buffer [c_1l][b_0][column] the peers write into.  (Measured and dropped: the same
// kernel made persistent - no gain, cluster launches are not the cost - and a 512-thread one-CTA-per-SM version
// that prefetches the next item's rows into registers: 52.6 instead of 43.3 ms per 4096 cfg3 states.)
constexpr size_t kC2PushSmem = (size_t)(4 * 64 * 8 + 8 * 64 * 8) * sizeof(c128);
__device__ __forceinline__ void c2_st_remote(c128* local, uint32_t rank, c128 v) {
  uint32_t la = (uint32_t)__cvta_generic_to_shared(local), ra;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(la), "r"(rank));
  asm volatile("st.shared::cluster.v2.f64 [%0], {%1, %2};" ::"r"(ra), "d"(v.x), "d"(v.y) : "memory");
}

__global__ void __cluster_dims__(kC2Cluster, 1, 1) __launch_bounds__(kC2Threads, 2)
cell_fft2_cluster_push_kernel(int n, int n1, int64_t item0, const c128* __restrict__ in, c128* __restrict__ out,
                              const c128* __restrict__ tab_g, double scale) {
  extern __shared__ __align__(16) c128 c2_S[];
  c128* Sx = c2_S;                 // [4 rows][64 positions][8 columns]   axis-1 scratch (one half at a time)
  c128* P2 = c2_S + 4 * 64 * 8;    // [8 c_1l][64 b_0][8 columns]         axis-0 buffer, written by the peers
  __shared__ c128 tab[64];
  const int tid = threadIdx.x;
  uint32_t rank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  const int64_t item = item0 + (blockIdx.x >> 3);
  const int groups = n >> 3;
  const int64_t t = item / groups;
  const int jg = (int)(item - t * groups);
  if (tid < 64) tab[tid] = tab_g[tid];
  const int jl = tid & 7, a = (tid >> 3) & 7, hi = tid >> 6;
  const int64_t m = (int64_t)4096 * n;
  const c128* src = in + t * m + (8 * jg + jl);
  c128 x[2][8];
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int b0 = 8 * (int)rank + 4 * h + hi;
#pragma unroll
    for (int tt = 0; tt < 8; ++tt) x[h][tt] = src[(int64_t)(b0 * 64 + a + 8 * tt) * n];
  }
  // every CTA of the cluster must be running before its shared memory is written remotely
  asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory");
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    dft_small<8>(x[h], +1, nullptr, 0);
    if (h) __syncthreads();  // the first half's scratch reads are over
#pragma unroll
    for (int tt = 0; tt < 8; ++tt) Sx[(hi * 64 + a * 8 + tt) * 8 + jl] = x[h][tt];
    __syncthreads();
    c128 v[8];
#pragma unroll
    for (int tt = 0; tt < 8; ++tt) {
      v[tt] = Sx[(hi * 64 + a + 8 * tt) * 8 + jl];
      if (tt) v[tt] = cmul(v[tt], tw_lookup(tab, tt * a, +1));
    }
    dft_small<8>(v, +1, nullptr, 0);
    if (h == 0) asm volatile("barrier.cluster.wait.aligned;" ::: "memory");
    // output tt is block column c_1 = a + 8 tt: column a of CTA tt, row b_0 = 8 rank + 4 h + hi
    c128* slot = P2 + (a * 64 + 8 * (int)rank + 4 * h + hi) * 8 + jl;
#pragma unroll
    for (int tt = 0; tt < 8; ++tt) c2_st_remote(slot, (uint32_t)tt, v[tt]);
  }
  c2_cluster_sync();  // all eight CTAs have delivered: P2 holds the axis-1 transform of EVERY block row for my columns
#pragma unroll
  for (int pass = 0; pass < 2; ++pass) {
    const int jj = hi + 4 * pass;
#pragma unroll
    for (int tt = 0; tt < 8; ++tt) x[pass][tt] = P2[(a * 64 + jj + 8 * tt) * 8 + jl];
    dft_small<8>(x[pass], +1, nullptr, 0);
  }
  __syncthreads();
#pragma unroll
  for (int pass = 0; pass < 2; ++pass) {
    const int jj = hi + 4 * pass;
#pragma unroll
    for (int tt = 0; tt < 8; ++tt) P2[(a * 64 + jj * 8 + tt) * 8 + jl] = x[pass][tt];
  }
  __syncthreads();
  const int j = 8 * jg + jl;
  const int j0 = j / n1, j1 = j - j0 * n1;
  const int n0 = n / n1;
  const int64_t l1 = (int64_t)64 * n1;
  c128* dst = out + t * m + (int64_t)j0 * l1 + (int64_t)(8 * (int)rank + a) * n1 + j1;
#pragma unroll
  for (int pass = 0; pass < 2; ++pass) {
    const int jj2 = hi + 4 * pass;
    c128 v[8];
#pragma unroll
    for (int tt = 0; tt < 8; ++tt) {
      v[tt] = P2[(a * 64 + jj2 + 8 * tt) * 8 + jl];
      if (tt) v[tt] = cmul(v[tt], tw_lookup(tab, tt * jj2, +1));
    }
    dft_small<8>(v, +1, nullptr, 0);
#pragma unroll
    for (int tt = 0; tt < 8; ++tt) dst[(int64_t)(jj2 + 8 * tt) * n0 * l1] = cscale(v[tt], scale);
  }
}

int cell_fft_cluster_mode() {
  static int cached = -1;  // 0: off (two passes), 1: pull variant, 2: push variant
  if (cached < 0) {
    const char* v = getenv("SQB_CELL_FFT_CLUSTER");
    cached = v ? atoi(v) : 2;
  }
  return cached;
}

// ------------------------------------------------------------------------------------------ host side
bool fft_radix16_on() {
  static int cached = -1;
  if (cached < 0) {
    const char* v = getenv("SQB_FFT_RADIX16");
    cached = (v && v[0] == '0') ? 0 : 1;
  }
  return cached == 1;
}

// allow16: power-of-two lengths may use radix-16 stages when that saves a stage (2^7, 2^8: two stages instead of
// three; 2^10 .. 2^12: three instead of four); the fewest 16s that reach the minimum, 16s first.
bool factorize(int len, int* radix, int* n_stages, bool allow16 = false) {
  int n = 0, rem = len;
  if (allow16 && len >= 128 && (len & (len - 1)) == 0 && fft_radix16_on()) {
    int k = 0;
    while ((1 << k) < len) ++k;
    int best16 = 0, best = (k + 2) / 3;
    for (int n16 = 1; 4 * n16 <= k; ++n16) {
      const int st = n16 + (k - 4 * n16 + 2) / 3;
      if (st < best) {
        best = st;
        best16 = n16;
      }
    }
    for (int q = 0; q < best16; ++q) {
      radix[n++] = 16;
      rem /= 16;
    }
  }
  while (rem % 8 == 0) { radix[n++] = 8; rem /= 8; if (n >= kMaxStages) return false; }
  while (rem % 4 == 0) { radix[n++] = 4; rem /= 4; if (n >= kMaxStages) return false; }
  while (rem % 2 == 0) { radix[n++] = 2; rem /= 2; if (n >= kMaxStages) return false; }
  for (int p = 3; p <= kMaxRadix && rem > 1; p += 2) {
    while (rem % p == 0) { radix[n++] = p; rem /= p; if (n >= kMaxStages) return false; }
  }
  if (rem != 1) return false;
  *n_stages = n;
  return true;
}

// split len = l1 * l2 with both <= kMaxSmemLen (len itself is smooth); returns false if impossible
bool split_len(int len, int* l1, int* l2) {
  int best = 0;
  for (int a = 1; (int64_t)a * a <= len; ++a) {
    if (len % a) continue;
    int b = len / a;
    if (b <= kMaxSmemLen && a <= kMaxSmemLen) best = a;  // largest a <= sqrt(len) -> most balanced
  }
  if (!best) return false;
  *l1 = best;
  *l2 = len / best;
  return true;
}

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

Digits digits1(int64_t dim, int64_t stride) {
  Digits d{};
  d.nd = 1;
  d.dims[0] = (int)dim;  // callers keep every digit below 2^31
  d.strides[0] = stride;
  return d;
}

bool smooth(int len) {
  int radix[kMaxStages], ns;
  return factorize(len, radix, &ns);
}

bool fft_inplace_on() {
  static int cached = -1;
  if (cached < 0) {
    const char* v = getenv("SQB_FFT_INPLACE");
    cached = (v && v[0] == '0') ? 0 : 1;
  }
  return cached == 1;
}

int launch_pass(FftPass& P, const c128* in, c128* out, c128* tab, cudaStream_t st) {
  if (!factorize(P.len, P.radix, &P.n_stages, P.mom_nd == 0)) return SQB_E_UNSUPPORTED;
  if (P.len > kMaxSmemLen) return SQB_E_UNSUPPORTED;
  // buffers: first fixed-radix stage reads global, last writes global -> S-1 shared round trips
  auto fixed = [](int r) { return r == 2 || r == 3 || r == 4 || r == 5 || r == 7 || r == 8 || r == 16; };
  const bool copy_in = P.n_stages == 0 || !fixed(P.radix[0]);
  const bool copy_out = P.n_stages == 0 || !fixed(P.radix[P.n_stages - 1]);
  const int smem_writes = P.n_stages - 1 + (copy_in ? 1 : 0) + (copy_out ? 1 : 0);  // stages writing shared
  const int n_buf = smem_writes >= 2 ? 2 : 1;
  int n_buf_used = n_buf;
  // lines per CTA (power of two): enough for 128..256 B coalesced runs across lines and >= 1024 elements of
  // work per CTA, small enough for several CTAs per SM (<= ~72 KB of shared memory when possible)
  int lpc = 1;
  while (lpc * 2 <= kMaxLines && lpc * 2 * P.len <= 2048) lpc *= 2;
  if (P.inner_fastest || P.out_inner_fastest) {
    // strided lines: at least 4 adjacent lines (64 B runs), 8 when the buffers stay small
    int want = 8;
    while (want > 1 && (size_t)want * P.len * n_buf * sizeof(c128) > 72 * 1024) want /= 2;
    if (want < 4 && (size_t)4 * P.len * n_buf * sizeof(c128) <= 160 * 1024) want = 4;
    if (lpc < want) lpc = want;
  }
  const int64_t n_lines = P.n_outer * P.n_inner;
  while (lpc > 1 && lpc / 2 >= n_lines) lpc /= 2;
  int variant = 0;
  for (int q = 0; q < P.n_stages; ++q) {
    const int r = P.radix[q];
    if (!fixed(r)) variant = 2;
    else if ((r == 3 || r == 5 || r == 7) && variant < 1) variant = 1;
  }
  if (P.radix[0] == 16) {
    variant = 3;  // pure power of two (factorize): {16, 8, 4, 2} only
    // two CTAs / SM (128 registers): up to 32 lines per CTA (64 KB) keep ~128 KB in flight per SM; measured on the
    // position stage of the weak-scaling k-grids: 128 x 64 blocks 71.2 -> 63.5 ms, 256 x 64 69.7 -> 64.5 ms per rank step
    static int lines16 = -1;
    if (lines16 < 0) {
      const char* v = getenv("SQB_FFT_LINES16");
      lines16 = v ? atoi(v) : 32;
    }
    int want16 = lines16;
    while (want16 > 1 && (size_t)want16 * P.len * n_buf * sizeof(c128) > 100 * 1024) want16 /= 2;
    if ((P.inner_fastest || P.out_inner_fastest) && lpc < want16) {
      lpc = want16;
      while (lpc > 1 && lpc / 2 >= n_lines) lpc /= 2;
    }
  }
  // Long strided power-of-two lines with three stages (256, 512 points: the across-block cell FFT of an enlarged
  // k-grid): the two-buffer version can only take 8 / 4 adjacent lines (128 / 64 B runs).  Running the middle stage
  // in place on ONE buffer, with 512 threads so that every thread owns exactly one butterfly per stage, doubles
  // the lines per CTA in the same shared memory.  SQB_FFT_INPLACE=0 switches it off.
  P.inplace_mid = 0;
  int threads = kFftThreads;
  if (P.mom_nd == 0 && variant == 0 && P.n_stages == 3 && P.len >= 256 && P.inner_fastest && P.out_inner_fastest && fft_inplace_on()) {
    const int per_line_mid = P.len / P.radix[1];
    int wide = kFftThreadsWide / per_line_mid;  // lines such that the middle stage has one butterfly per thread
    while (wide > 1 && wide / 2 >= n_lines) wide /= 2;
    if (wide > lpc && wide <= kMaxLines && (size_t)(wide + 1) * P.len * sizeof(c128) <= 100 * 1024) {
      lpc = wide;
      n_buf_used = 1;
      P.inplace_mid = 1;
      threads = kFftThreadsWide;
    }
  }
  P.lines_per_cta = lpc;
  auto* kern = P.inplace_mid ? fft_pass_kernel<0, kFftThreadsWide>
                             : (variant == 0 ? fft_pass_kernel<0, kFftThreads>
                                             : (variant == 1 ? fft_pass_kernel<1, kFftThreads>
                                                             : (variant == 2 ? fft_pass_kernel<2, kFftThreads> : fft_pass_kernel<3, kFftThreads>)));
  if (P.mom_nd > 0) {
    // K6 epilogue: every CTA must stay inside one state (lines per state divisible by the lines per CTA)
    const int64_t lines_per_state = n_lines / P.mom_lead;
    while (lpc > 1 && lines_per_state % lpc) lpc /= 2;
    P.lines_per_cta = lpc;
    const int64_t groups_per_state = lines_per_state / lpc;
    int groups = 1;
    while (groups * 2 <= 64 && groups_per_state % (groups * 2) == 0) groups *= 2;
    P.mom_groups = groups;
    kern = variant == 0 ? fft_pass_kernel<0, kFftThreads, true>
                        : (variant == 1 ? fft_pass_kernel<1, kFftThreads, true> : fft_pass_kernel<2, kFftThreads, true>);
  }
  const size_t smem = ((size_t)n_buf_used * lpc * P.len + P.len) * sizeof(c128);
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  if (e != cudaSuccess) return (int)e;
  fft_table_kernel<<<(P.len + 255) / 256, 256, 0, st>>>(tab, P.len);
  const int64_t per_cta = (int64_t)lpc * (P.mom_nd > 0 ? P.mom_groups : 1);
  const int64_t grid = (n_lines + per_cta - 1) / per_cta;
  if (grid > 2147483647LL) return SQB_E_UNSUPPORTED;
  kern<<<(unsigned)grid, threads, smem, st>>>(P, in, out, tab);
  SQB_LAUNCH_CHECK();
  return SQB_OK;
}

// ------------------------------------------------------------------------------------------ Bluestein (any length)
// Axis lengths with a prime factor above kMaxRadix (the shared-memory passes stop at 31) run as a chirp-z
// convolution on the power-of-two passes: with w[n] = exp(sign i pi n^2 / L),
//   y[k] = w[k] * sum_n (x[n] w[n]) * conj(w[k - n])
// i.e. three M-point FFTs (M = 2^p >= 2 L - 1), one of them of the filter only.  The reference takes any length
// through numpy's FFT; this keeps the mirror from refusing e.g. 37- or 101-point axes.
__global__ void bluestein_chirp_kernel(c128* __restrict__ w, int len, int sign) {
  for (int n = blockIdx.x * blockDim.x + threadIdx.x; n < len; n += gridDim.x * blockDim.x) {
    const int64_t q = ((int64_t)n * n) % (2 * (int64_t)len);  // exact reduction of the phase n^2 pi / L
    double sn, cs;
    sincospi((double)sign * (double)q / (double)len, &sn, &cs);
    w[n] = make_double2(cs, sn);
  }
}
__global__ void bluestein_filter_kernel(c128* __restrict__ f, const c128* __restrict__ w, int len, int m) {
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < m; j += gridDim.x * blockDim.x) {
    c128 v = make_double2(0.0, 0.0);
    if (j < len) v = cconj(w[j]);
    else if (m - j < len) v = cconj(w[m - j]);
    f[j] = v;
  }
}
// mode 0: a[o][j][i] = j < len ? in[o][j][i] * w[j] : 0      (pad to m)
// mode 1: a[o][j][i] *= f[j]
// mode 2: out[o][k][i] = a[o][k][i] * w[k] * scale            (k < len)
__global__ void bluestein_pointwise_kernel(int mode, int64_t outer, int len, int m, int64_t inner,
                                           const c128* __restrict__ in, c128* __restrict__ a,
                                           const c128* __restrict__ tabv, c128* __restrict__ out, double scale) {
  const int64_t rows = mode == 2 ? len : m;
  const int64_t total = outer * rows * inner;
  for (int64_t g = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; g < total; g += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = g % inner;
    const int64_t j = (g / inner) % rows;
    const int64_t o = g / (inner * rows);
    if (mode == 0) {
      a[g] = j < len ? cmul(in[(o * len + j) * inner + i], tabv[j]) : make_double2(0.0, 0.0);
    } else if (mode == 1) {
      a[g] = cmul(a[g], tabv[j]);
    } else {
      out[g] = cscale(cmul(a[(o * m + j) * inner + i], tabv[j]), scale);
    }
  }
}

int bluestein_m(int len) {
  int m = 1;
  while (m < 2 * len - 1) m *= 2;
  return m;
}
// elements of scratch a Bluestein axis needs behind the chirp table: filter + padded array (+ the four-step scratch
// of the M-point passes when M does not fit shared memory)
size_t bluestein_scratch_elems(int len, int64_t lines) {
  const size_t m = (size_t)bluestein_m(len);
  size_t e = align_up((size_t)len, 16) + align_up(m, 16) + align_up((size_t)lines * m, 16);
  if (m > (size_t)kMaxSmemLen) e += align_up((size_t)lines * m, 16);
  return e;
}

int fft_one_axis(int64_t outer, int len, int64_t inner, int sign, double scale, const c128* in, c128* out, c128* tab,
                 c128* scratch, cudaStream_t st);

int bluestein_axis(int64_t outer, int len, int64_t inner, int sign, double scale, const c128* in, c128* out, c128* tab,
                   c128* scratch, cudaStream_t st) {
  if (!scratch) return SQB_E_WORKSPACE;
  const int m = bluestein_m(len);
  const int64_t lines = outer * inner;
  c128* w = scratch;
  c128* f = w + align_up((size_t)len, 16);
  c128* a = f + align_up((size_t)m, 16);
  c128* scratch2 = m > kMaxSmemLen ? a + align_up((size_t)lines * m, 16) : nullptr;
  const int64_t total = lines * m;
  const int grid = (int)sqb_min64((total + 255) / 256, 148 * 16);
  bluestein_chirp_kernel<<<(len + 255) / 256, 256, 0, st>>>(w, len, sign);
  bluestein_filter_kernel<<<(m + 255) / 256, 256, 0, st>>>(f, w, len, m);
  SQB_LAUNCH_CHECK();
  int rc = fft_one_axis(1, m, 1, -1, 1.0, f, f, tab, scratch2, st);
  if (rc) return rc;
  bluestein_pointwise_kernel<<<grid, 256, 0, st>>>(0, outer, len, m, inner, in, a, w, nullptr, 1.0);
  SQB_LAUNCH_CHECK();
  rc = fft_one_axis(outer, m, inner, -1, 1.0, a, a, tab, scratch2, st);
  if (rc) return rc;
  bluestein_pointwise_kernel<<<grid, 256, 0, st>>>(1, outer, len, m, inner, nullptr, a, f, nullptr, 1.0);
  SQB_LAUNCH_CHECK();
  rc = fft_one_axis(outer, m, inner, +1, 1.0 / (double)m, a, a, tab, scratch2, st);
  if (rc) return rc;
  bluestein_pointwise_kernel<<<grid, 256, 0, st>>>(2, outer, len, m, inner, nullptr, a, w, out, scale);
  SQB_LAUNCH_CHECK();
  return SQB_OK;
}

// One axis of a [outer, len, inner] array, in -> out (in == out allowed when the line fits shared memory).
int fft_one_axis(int64_t outer, int len, int64_t inner, int sign, double scale, const c128* in, c128* out,
                 c128* tab, c128* scratch, cudaStream_t st) {
  if (outer >= (1LL << 31) || inner >= (1LL << 31)) return SQB_E_UNSUPPORTED;
  if (!smooth(len)) return bluestein_axis(outer, len, inner, sign, scale, in, out, tab, scratch, st);
  if (len <= kMaxSmemLen) {
    FftPass P{};
    P.len = len;
    P.n_outer = outer;
    P.n_inner = inner;
    P.in_outer = P.out_outer = digits1(outer, (int64_t)len * inner);
    P.in_inner = P.out_inner = digits1(inner, 1);
    P.in_sl = P.out_sl = inner;
    P.inner_fastest = P.out_inner_fastest = inner > 1;
    P.sign = sign;
    P.scale = scale;
    P.tw_len = 0;
    P.tw_div = 1;
    return launch_pass(P, in, out, tab, st);
  }
  // four-step: view the line as [l1, l2] (n = n1*l2 + n2), every element carrying `inner` trailing entries.
  // pass A: l1-point transforms over n1 for every (outer, n2, i), twiddle W_len^(k1*n2), in -> scratch
  // pass B: l2-point transforms over n2 for every (outer, k1, i), output index k1 + l1*k2, scratch -> out
  int l1, l2;
  if (!split_len(len, &l1, &l2) || !smooth(l1) || !smooth(l2)) return SQB_E_UNSUPPORTED;
  if (!scratch || (int64_t)l2 * inner >= (1LL << 31) || (int64_t)l1 * inner >= (1LL << 31)) return SQB_E_UNSUPPORTED;
  {
    FftPass P{};
    P.len = l1;
    P.n_outer = outer;
    P.n_inner = (int64_t)l2 * inner;  // combined (n2, i), contiguous
    P.in_outer = P.out_outer = digits1(outer, (int64_t)len * inner);
    P.in_inner = P.out_inner = digits1(P.n_inner, 1);
    P.in_sl = P.out_sl = (int64_t)l2 * inner;
    P.inner_fastest = P.out_inner_fastest = 1;
    P.sign = sign;
    P.scale = 1.0;
    P.tw_len = len;
    P.tw_div = inner;
    int rc = launch_pass(P, in, scratch, tab, st);
    if (rc) return rc;
  }
  {
    FftPass P{};
    P.len = l2;
    P.n_outer = outer;
    P.n_inner = (int64_t)l1 * inner;  // (k1, i)
    P.in_outer = P.out_outer = digits1(outer, (int64_t)len * inner);
    Digits din{}, dout{};
    din.nd = dout.nd = 2;
    din.dims[0] = dout.dims[0] = l1;
    din.dims[1] = dout.dims[1] = (int)inner;
    din.strides[0] = (int64_t)l2 * inner;
    din.strides[1] = 1;
    dout.strides[0] = inner;
    dout.strides[1] = 1;
    P.in_inner = din;
    P.out_inner = dout;
    P.in_sl = inner;
    P.out_sl = (int64_t)l1 * inner;
    P.inner_fastest = inner > 1;
    P.out_inner_fastest = 1;
    P.sign = sign;
    P.scale = scale;
    P.tw_len = 0;
    P.tw_div = 1;
    return launch_pass(P, scratch, out, tab, st);
  }
}

}  // namespace

// table area: the longest table any pass of this transform builds (a Bluestein axis runs M-point passes)
static int fft_table_len(int nd, const int* dims_host) {
  int max_len = 1;
  for (int a = 0; a < nd; ++a) {
    const int l = smooth(dims_host[a]) ? dims_host[a] : bluestein_m(dims_host[a]);
    if (l > max_len) max_len = l;
  }
  return max_len;
}

extern "C" size_t sqb_fft_workspace_bytes(int nd, const int* dims_host, int64_t batch) {
  if (nd < 1 || nd > SQB_MAX_DIMS || !dims_host || batch < 0) return 0;
  int64_t total = batch;
  bool any_long = false;
  for (int a = 0; a < nd; ++a) {
    if (dims_host[a] < 1) return 0;
    total *= dims_host[a];
    if (dims_host[a] > kMaxSmemLen) any_long = true;
  }
  size_t bytes = align_up((size_t)fft_table_len(nd, dims_host) * sizeof(c128), 256);
  size_t scratch = any_long ? (size_t)total : 0;  // four-step split of long smooth axes
  for (int a = 0; a < nd; ++a)
    if (!smooth(dims_host[a])) {
      const size_t need = bluestein_scratch_elems(dims_host[a], total / dims_host[a]);
      if (need > scratch) scratch = need;
    }
  if (scratch) bytes += align_up(scratch * sizeof(c128), 256);
  return bytes + 256;
}

extern "C" int sqb_fft_c2c(int nd, const int* dims_host, int64_t batch, int sign, const sqb_c128* in,
                           sqb_c128* out, void* workspace, size_t workspace_bytes, void* stream) {
  if (nd < 1 || nd > SQB_MAX_DIMS || !dims_host || !in || !out || batch < 0 || (sign != 1 && sign != -1))
    return SQB_E_BADARG;
  if (!workspace || workspace_bytes < sqb_fft_workspace_bytes(nd, dims_host, batch)) return SQB_E_WORKSPACE;
  if (batch == 0) return SQB_OK;
  int64_t total = 1;
  int max_len = 1;
  for (int a = 0; a < nd; ++a) {
    if (dims_host[a] < 1) return SQB_E_BADARG;
    total *= dims_host[a];
    if (dims_host[a] > max_len) max_len = dims_host[a];
  }
  c128* tab = reinterpret_cast<c128*>(workspace);
  c128* scratch = reinterpret_cast<c128*>(reinterpret_cast<char*>(workspace) +
                                          align_up((size_t)fft_table_len(nd, dims_host) * sizeof(c128), 256));
  const double scale = 1.0 / sqrt((double)total);
  cudaStream_t st = sqb_stream(stream);
  const c128* src = reinterpret_cast<const c128*>(in);
  c128* dst = reinterpret_cast<c128*>(out);
  // last axis first (contiguous), then the strided ones in place on `out`
  int64_t inner = 1;
  for (int a = nd - 1; a >= 0; --a) {
    const int len = dims_host[a];
    const int64_t outer = batch * (total / inner / len);
    const double sc = (a == 0) ? scale : 1.0;
    int rc = fft_one_axis(outer, len, inner, sign, sc, src, dst, tab, scratch, st);
    if (rc) return rc;
    src = dst;
    inner *= len;
  }
  return SQB_OK;
}

extern "C" int sqb_fft_axis(int64_t outer, int len, int64_t inner, int sign, const sqb_c128* in, sqb_c128* out,
                            void* workspace, size_t workspace_bytes, void* stream) {
  if (outer < 0 || len < 1 || inner < 1 || !in || !out || (sign != 1 && sign != -1)) return SQB_E_BADARG;
  int dims[1] = {len};
  if (!workspace || workspace_bytes < sqb_fft_workspace_bytes(1, dims, outer * inner)) return SQB_E_WORKSPACE;
  if (outer == 0) return SQB_OK;
  c128* tab = reinterpret_cast<c128*>(workspace);
  c128* scratch = reinterpret_cast<c128*>(reinterpret_cast<char*>(workspace) +
                                          align_up((size_t)fft_table_len(1, dims) * sizeof(c128), 256));
  return fft_one_axis(outer, len, inner, sign, 1.0 / sqrt((double)len), reinterpret_cast<const c128*>(in),
                      reinterpret_cast<c128*>(out), tab, scratch, sqb_stream(stream));
}

extern "C" size_t sqb_cell_fft_workspace_bytes(int nd, const int* repeat_host) {
  if (nd < 1 || nd > SQB_MAX_DIMS || !repeat_host) return 0;
  int max_len = 1;
  for (int a = 0; a < nd; ++a)
    if (repeat_host[a] > max_len) max_len = repeat_host[a];
  return align_up((size_t)max_len * sizeof(c128), 256) + 256;
}

struct MomRequest {  // K6 fused into the last cell-FFT pass (direction 1 only); see sqb_cell_fft_moments
  int stage;                 // 0: run the in-place axes first; 1: in_scratch already holds them
  const double* spacing;     // [nd] host
  const double* length;      // [nd] host
  const double* offsets;     // [nd][lead] device or nullptr
  int wrapped;
  double* out;               // [lead][nd][5] device
  c128* cs_tab;              // device scratch: cos / sin tables
  double* partials;          // device scratch
  size_t partial_doubles;
};

static int cell_fft_impl(int nd, const int* shape_host, const int* repeat_host, int direction, int64_t lead,
                         int ranks, sqb_c128* in_scratch, sqb_c128* out, void* workspace, size_t workspace_bytes,
                         void* stream, const MomRequest* mom = nullptr) {
  SqbDims d;
  int64_t n, nk;
  int rc = sqb_fill_dims(d, nd, shape_host, repeat_host, &n, &nk);
  if (rc) return rc;
  if (!in_scratch || (!out && !mom) || lead < 0 || (direction != 0 && direction != 1)) return SQB_E_BADARG;
  if (!workspace || workspace_bytes < sqb_cell_fft_workspace_bytes(nd, repeat_host)) return SQB_E_WORKSPACE;
  if (!mom && nd > 1 && in_scratch == out) return SQB_E_BADARG;  // the scatter / gather pass is out of place
  for (int a = 0; a < nd; ++a)
    if (d.repeat[a] > kMaxSmemLen || !smooth(d.repeat[a])) return SQB_E_UNSUPPORTED;
  if (ranks < 1 || d.repeat[0] % ranks || (ranks > 1 && direction != 1)) return SQB_E_BADARG;
  if (lead == 0) return SQB_OK;
  if (lead >= (1LL << 31)) return SQB_E_UNSUPPORTED;
  c128* tab = reinterpret_cast<c128*>(workspace);
  cudaStream_t st = sqb_stream(stream);
  const int64_t m = n * nk;
  const double scale = 1.0 / sqrt((double)nk);
  // rank-major input (ranks > 1): [ranks][lead][n_k / ranks][N], the receive buffer of the k-shard -> time-shard
  // all-to-all; block axis 0 is split as b_0 = r * P0 + b_0l with P0 = R_0 / ranks.
  const int P0 = d.repeat[0] / ranks;
  const int64_t chunk = (nk / ranks) * n;       // elements of one (rank, lead) slab
  const int64_t inner0 = (nk / d.repeat[0]) * n;  // stride of b_0l inside a slab
  c128* src = reinterpret_cast<c128*>(in_scratch);
  c128* dst = reinterpret_cast<c128*>(out);

  // supercell strides S_a = prod_{a' > a} L_a'
  int64_t S[SQB_MAX_DIMS];
  {
    int64_t s = 1;
    for (int a = nd - 1; a >= 0; --a) {
      S[a] = s;
      s *= (int64_t)d.shape[a] * d.repeat[a];
    }
  }
  // "cell layout" [lead][b_0..b_{nd-1}][j]  <->  "position layout" [lead][x_0..x_{nd-1}], x_a = c_a N_a + j_a.
  // direction 1 (cell -> position, sign +1): axes 0..nd-2 in place on the cell layout (this CLOBBERS
  //   in_scratch), the last axis scatters into `out`.
  // direction 0 (position -> cell, sign -1): the last axis gathers from the position layout into `out`,
  //   then the other axes run in place on `out`.
  auto scatter_pass = [&](const c128* pin, c128* pout, bool cell_is_input) -> int {
    const int a = nd - 1;
    FftPass P{};
    P.len = d.repeat[a];
    P.n_outer = lead * (nk / d.repeat[a]);
    P.n_inner = n;
    Digits cell_outer = digits1(P.n_outer, (int64_t)d.repeat[a] * n);
    const Digits cell_inner = digits1(n, 1);
    if (ranks > 1 && nd > 1) {
      // outer index o = (lead, c_0 = (r, c_0l), c_1 .. c_{nd-2}) in C order; slab strides for (r, lead)
      cell_outer.nd = nd + 1;
      cell_outer.dims[0] = (int)lead;  cell_outer.strides[0] = chunk;
      cell_outer.dims[1] = ranks;      cell_outer.strides[1] = lead * chunk;
      cell_outer.dims[2] = P0;         cell_outer.strides[2] = inner0;
      int64_t st_q = inner0;
      for (int q = 1; q < nd - 1; ++q) {
        st_q /= d.repeat[q];
        cell_outer.dims[q + 2] = d.repeat[q];
        cell_outer.strides[q + 2] = st_q;
      }
    }
    if (ranks > 1 && nd == 1) cell_outer = digits1(lead, chunk);
    Digits pos_outer{}, pos_inner{};
    pos_outer.nd = nd;  // digits (lead, c_0 .. c_{nd-2})
    pos_outer.dims[0] = (int)lead;
    pos_outer.strides[0] = m;
    for (int q = 0; q < nd - 1; ++q) {
      pos_outer.dims[q + 1] = d.repeat[q];
      pos_outer.strides[q + 1] = (int64_t)d.shape[q] * S[q];
    }
    pos_inner.nd = nd;  // digits (j_0 .. j_{nd-1})
    for (int q = 0; q < nd; ++q) {
      pos_inner.dims[q] = d.shape[q];
      pos_inner.strides[q] = S[q];
    }
    const int64_t cell_sl = n, pos_sl = d.shape[a];
    if (cell_is_input) {
      P.in_outer = cell_outer; P.in_inner = cell_inner; P.in_sl = cell_sl;
      P.out_outer = pos_outer; P.out_inner = pos_inner; P.out_sl = pos_sl;
      P.sign = +1;
      if (ranks > 1 && nd == 1) {  // the scattered axis is the rank-split axis itself
        P.in_split = P0;
        P.in_sl_hi = lead * chunk;
      }
    } else {
      P.in_outer = pos_outer; P.in_inner = pos_inner; P.in_sl = pos_sl;
      P.out_outer = cell_outer; P.out_inner = cell_inner; P.out_sl = cell_sl;
      P.sign = -1;
    }
    P.inner_fastest = P.out_inner_fastest = 1;
    P.scale = scale;
    P.tw_len = 0;
    P.tw_div = 1;
    if (mom && cell_is_input) {
      if (m >= (1LL << 31)) return SQB_E_UNSUPPORTED;
      P.mom_nd = nd;
      P.mom_wrapped = mom->wrapped;
      P.mom_m = m;
      P.mom_lead = lead;
      P.mom_offsets = mom->offsets;
      P.mom_cs = mom->cs_tab;
      P.mom_partials = mom->partials;
      int off = 0;
      for (int q = 0; q < nd; ++q) {
        P.mom_grid[q] = d.shape[q] * d.repeat[q];
        P.mom_spacing[q] = mom->spacing[q];
        P.mom_length[q] = mom->length[q];
        P.mom_cs_off[q] = off;
        mom_table_kernel<<<(P.mom_grid[q] + 255) / 256, 256, 0, st>>>(mom->cs_tab + off, P.mom_grid[q]);
        off += P.mom_grid[q];
      }
      int rc2 = launch_pass(P, pin, pout, tab, st);
      if (rc2) return rc2;
      const int64_t per_cta = (int64_t)P.lines_per_cta * P.mom_groups;
      const int64_t n_ctas = (P.n_outer * P.n_inner + per_cta - 1) / per_cta;
      if ((size_t)n_ctas * (1 + 4 * nd) > mom->partial_doubles || n_ctas % lead) return SQB_E_WORKSPACE;
      mom_reduce_kernel<<<(unsigned)lead, 32, 0, st>>>(mom->partials, (int)(n_ctas / lead), nd, mom->out);
      SQB_LAUNCH_CHECK();
      return SQB_OK;
    }
    return launch_pass(P, pin, pout, tab, st);
  };
  auto inplace_axis = [&](int a, c128* p, int sign) -> int {
    int64_t inner = n;
    for (int q = a + 1; q < nd; ++q) inner *= d.repeat[q];
    int64_t outer = lead;
    for (int q = 0; q < a; ++q) outer *= d.repeat[q];
    if (ranks > 1 && a == 0) {
      // lines along b_0 cross the rank slabs: two-level position stride, one line per (lead, inner index)
      FftPass P{};
      P.len = d.repeat[0];
      P.n_outer = lead;
      P.n_inner = inner;
      P.in_outer = P.out_outer = digits1(lead, chunk);
      P.in_inner = P.out_inner = digits1(inner, 1);
      P.in_sl = P.out_sl = inner;
      P.in_split = P.out_split = P0;
      P.in_sl_hi = P.out_sl_hi = lead * chunk;
      P.inner_fastest = P.out_inner_fastest = 1;
      P.sign = sign;
      P.scale = 1.0;
      P.tw_len = 0;
      P.tw_div = 1;
      return launch_pass(P, p, p, tab, st);
    }
    // axes >= 1 never cross a slab: the order of the outer index is irrelevant for an in-place pass
    return fft_one_axis(outer, d.repeat[a], inner, sign, 1.0, p, p, tab, nullptr, st);
  };

  if (direction == 1 && nd == 2 && ranks == 1 && !mom && d.repeat[0] == 64 && d.repeat[1] == 64 && d.shape[1] % 8 == 0 &&
      in_scratch != out && cell_fft_cluster_mode() > 0) {
    // one pass over the state list: both block axes inside a cluster of 8 CTAs (cell_fft2_cluster_*_kernel).  A
    // device that refuses the cluster launch (e.g. a partitioned GPU without 8 co-schedulable SMs) falls through to
    // the pass per axis below: nothing has been written yet when the first launch fails.
    const bool push = cell_fft_cluster_mode() == 2;
    cudaError_t e = push ? cudaFuncSetAttribute(cell_fft2_cluster_push_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kC2PushSmem)
                         : cudaFuncSetAttribute(cell_fft2_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kC2Smem);
    bool launched = false;
    if (e == cudaSuccess) {
      fft_table_kernel<<<1, 256, 0, st>>>(tab, 64);
      const int64_t items = lead * (n / 8);
      const int64_t per_launch = (int64_t)1 << 26;  // clusters per launch (grid.x = 8 x clusters < 2^31)
      for (int64_t i0 = 0; i0 < items; i0 += per_launch) {
        const int64_t cnt = sqb_min64(per_launch, items - i0);
        if (push)
          cell_fft2_cluster_push_kernel<<<(unsigned)(cnt * kC2Cluster), kC2Threads, kC2PushSmem, st>>>((int)n, d.shape[1], i0,
                                                                                                    src, dst, tab, scale);
        else
          cell_fft2_cluster_kernel<<<(unsigned)(cnt * kC2Cluster), kC2Threads, kC2Smem, st>>>((int)n, d.shape[1], i0, src, dst,
                                                                                            tab, scale);
        e = cudaGetLastError();
        if (e != cudaSuccess) {
          if (i0 == 0) break;  // refused outright: use the passes below
          return (int)e;
        }
        launched = true;
      }
    } else {
      (void)cudaGetLastError();
    }
    if (launched) return SQB_OK;
  }
  if (direction == 1) {
    for (int a = 0; a < nd - 1 && !(mom && mom->stage == 1); ++a) {
      rc = inplace_axis(a, src, +1);
      if (rc) return rc;
    }
    return scatter_pass(src, dst, true);
  }
  rc = scatter_pass(src, dst, false);
  if (rc) return rc;
  for (int a = 0; a < nd - 1; ++a) {
    rc = inplace_axis(a, dst, -1);
    if (rc) return rc;
  }
  return SQB_OK;
}

extern "C" int sqb_cell_fft(int nd, const int* shape_host, const int* repeat_host, int direction, int64_t lead,
                            sqb_c128* in_scratch, sqb_c128* out, void* workspace, size_t workspace_bytes,
                            void* stream) {
  return cell_fft_impl(nd, shape_host, repeat_host, direction, lead, 1, in_scratch, out, workspace, workspace_bytes,
                       stream);
}

// K6 fused into the last pass of the cell FFT: the position-basis states are never written.
extern "C" size_t sqb_cell_fft_moments_workspace_bytes(int nd, const int* shape_host, const int* repeat_host,
                                                       int64_t lead) {
  if (nd < 1 || nd > SQB_MAX_DIMS || !shape_host || !repeat_host || lead < 0) return 0;
  size_t grid_sum = 0;
  int64_t n = 1, nk = 1;
  for (int a = 0; a < nd; ++a) {
    grid_sum += (size_t)shape_host[a] * repeat_host[a];
    n *= shape_host[a];
    nk *= repeat_host[a];
  }
  const int64_t lines = lead * (nk / repeat_host[nd - 1]) * n;  // one partial row per CTA, at worst one line per CTA
  return align_up(sqb_cell_fft_workspace_bytes(nd, repeat_host), 256) + align_up(grid_sum * sizeof(c128), 256) +
         align_up((size_t)lines * (1 + 4 * nd) * sizeof(double), 256) + 256;
}

extern "C" int sqb_cell_fft_moments(int nd, const int* shape_host, const int* repeat_host, int64_t lead, int stage,
                                    sqb_c128* in_scratch, const double* spacing_host, const double* length_host,
                                    const double* offsets, int wrapped, double* out, void* workspace,
                                    size_t workspace_bytes, void* stream) {
  if (nd < 1 || nd > SQB_MAX_DIMS || !shape_host || !repeat_host || !spacing_host || !length_host || !out ||
      (stage != 0 && stage != 1))
    return SQB_E_BADARG;
  if (!workspace || workspace_bytes < sqb_cell_fft_moments_workspace_bytes(nd, shape_host, repeat_host, lead))
    return SQB_E_WORKSPACE;
  if (lead == 0) return SQB_OK;
  char* base = reinterpret_cast<char*>(align_up(reinterpret_cast<size_t>(workspace), 256));
  const size_t fft_ws = align_up(sqb_cell_fft_workspace_bytes(nd, repeat_host), 256);
  size_t grid_sum = 0;
  for (int a = 0; a < nd; ++a) grid_sum += (size_t)shape_host[a] * repeat_host[a];
  MomRequest mom;
  mom.stage = stage;
  mom.spacing = spacing_host;
  mom.length = length_host;
  mom.offsets = offsets;
  mom.wrapped = wrapped;
  mom.out = out;
  mom.cs_tab = reinterpret_cast<c128*>(base + fft_ws);
  mom.partials = reinterpret_cast<double*>(base + fft_ws + align_up(grid_sum * sizeof(c128), 256));
  mom.partial_doubles = (workspace_bytes - (size_t)(reinterpret_cast<char*>(mom.partials) - reinterpret_cast<char*>(workspace))) / sizeof(double);
  return cell_fft_impl(nd, shape_host, repeat_host, 1, lead, 1, in_scratch, nullptr, base, fft_ws, stream, &mom);
}

extern "C" int sqb_cell_fft_sharded(int nd, const int* shape_host, const int* repeat_host, int64_t lead, int ranks,
                                    sqb_c128* in_scratch, sqb_c128* out, void* workspace, size_t workspace_bytes,
                                    void* stream) {
  return cell_fft_impl(nd, shape_host, repeat_host, 1, lead, ranks, in_scratch, out, workspace, workspace_bytes,
                       stream);
}

extern "C" int sqb_fold_transform(int nd, const int* shape_host, const int* repeat_host, int64_t block_begin,
                                  int64_t batch, const sqb_c128* transform, sqb_c128* folded, void* workspace,
                                  size_t workspace_bytes, void* stream) {
  SqbDims d;
  int64_t n, nk;
  int rc = sqb_fill_dims(d, nd, shape_host, repeat_host, &n, &nk);
  if (rc) return rc;
  if (!transform || !folded || batch < 0 || block_begin < 0 || block_begin + batch > nk) return SQB_E_BADARG;
  if (batch == 0) return SQB_OK;
  // fused single pass when every cell axis is one fixed radix (SQB_FOLD_FUSED=0: the per-axis passes below)
  {
    static int fused_on = -1;
    if (fused_on < 0) {
      const char* v = getenv("SQB_FOLD_FUSED");
      fused_on = (v && v[0] == '0') ? 0 : 1;
    }
    bool ok = fused_on && n <= 2048 && transform != folded;
    for (int a = 0; a < nd; ++a) {
      const int l = d.shape[a];
      ok = ok && (l == 1 || l == 2 || l == 3 || l == 4 || l == 5 || l == 7 || l == 8 || l == 16);
    }
    if (ok) {
      const int n_last = d.shape[nd - 1];
      const int row_stride = (int)(n / n_last) * (n_last + ((n_last & 1) ? 0 : 1));
      int rows = (int)((64 * 1024) / ((size_t)row_stride * sizeof(c128)));
      if (rows < 1) rows = 1;
      if (rows > n) rows = (int)n;
      const size_t smem = ((size_t)n + 16 * SQB_MAX_DIMS + (size_t)rows * row_stride) * sizeof(c128);
      cudaError_t e = cudaFuncSetAttribute(fold_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return (int)e;
      for (int64_t b0 = 0; b0 < batch; b0 += 65535) {
        const int64_t cnt = sqb_min64(65535, batch - b0);
        const dim3 grid((unsigned)((n + rows - 1) / rows), (unsigned)cnt);
        fold_fused_kernel<<<grid, 256, smem, sqb_stream(stream)>>>(
            d, (int)n, rows, block_begin + b0, reinterpret_cast<const c128*>(transform) + b0 * n * n,
            reinterpret_cast<c128*>(folded) + b0 * n * n);
        SQB_LAUNCH_CHECK();
      }
      return SQB_OK;
    }
  }
  // (1) in-cell inverse FFT of every eigenvector: [batch*n, shape...] ortho, momentum -> position
  rc = sqb_fft_c2c(nd, shape_host, batch * n, +1, transform, folded, workspace, workspace_bytes, stream);
  if (rc) return rc;
  // (2) Bloch twiddle exp(+2 pi i q_a j_a / L_a)
  // grid.y is limited to 65535 blocks per launch
  for (int64_t b0 = 0; b0 < batch; b0 += 65535) {
    const int64_t cnt = sqb_min64(65535, batch - b0);
    const dim3 grid((unsigned)((n + kFoldRows - 1) / kFoldRows), (unsigned)cnt);
    c128* fb = reinterpret_cast<c128*>(folded) + b0 * n * n;
    fold_twiddle_kernel<<<grid, 256, (size_t)n * sizeof(c128), sqb_stream(stream)>>>(
        d, (int)n, block_begin + b0, cnt, +1, n, n, n * n, fb, fb);
    SQB_LAUNCH_CHECK();
  }
  return SQB_OK;
}

// The same factorisation for VECTORS (state lists) in Bloch order [lead, n_k, N]:
//   direction 1: Bloch momentum order -> cell layout  = in-cell FFT (sign) of every block, then the twiddle
//   direction 0: cell layout -> Bloch momentum order  = conjugate twiddle, then the in-cell FFT (-sign)
// `sign` is the direction of the ket transform momentum -> position (+1); bra (dual) indices pass -1.
extern "C" int sqb_bloch_cell_vectors(int nd, const int* shape_host, const int* repeat_host, int direction, int sign,
                                      int64_t lead, const sqb_c128* in, sqb_c128* out, void* workspace,
                                      size_t workspace_bytes, void* stream) {
  SqbDims d;
  int64_t n, nk;
  int rc = sqb_fill_dims(d, nd, shape_host, repeat_host, &n, &nk);
  if (rc) return rc;
  if (!in || !out || lead < 0 || (direction != 0 && direction != 1) || (sign != 1 && sign != -1)) return SQB_E_BADARG;
  if (lead == 0) return SQB_OK;
  if (nk > 65535 * (int64_t)65535) return SQB_E_UNSUPPORTED;
  const c128* src = reinterpret_cast<const c128*>(in);
  c128* dst = reinterpret_cast<c128*>(out);
  auto twiddle = [&](const c128* pin, c128* pout, int sg) -> int {
    for (int64_t b0 = 0; b0 < nk; b0 += 65535) {
      const int64_t cnt = sqb_min64(65535, nk - b0);
      const dim3 grid((unsigned)((lead + kFoldRows - 1) / kFoldRows), (unsigned)cnt);
      fold_twiddle_kernel<<<grid, 256, (size_t)n * sizeof(c128), sqb_stream(stream)>>>(
          d, (int)n, b0, cnt, sg, lead, nk * n, n, pin + b0 * n, pout + b0 * n);
      SQB_LAUNCH_CHECK();
    }
    return SQB_OK;
  };
  if (direction == 1) {
    rc = sqb_fft_c2c(nd, shape_host, lead * nk, sign, in, out, workspace, workspace_bytes, stream);
    if (rc) return rc;
    return twiddle(dst, dst, sign);
  }
  rc = twiddle(src, dst, -sign);
  if (rc) return rc;
  return sqb_fft_c2c(nd, shape_host, lead * nk, -sign, out, out, workspace, workspace_bytes, stream);
}
