// K4: complex128 FFT, norm="ortho", shared-memory mixed-radix Stockham.
//
// One CTA transforms `lines_per_cta` 1-D lines that it first stages into shared memory (HBM traffic:
// one read + one write of the array per axis pass: 32 B per element per pass).  Lines longer than
// kMaxSmemLen are split L = L1*L2 (four-step: strided L1-point pass with the inter-stage twiddle fused
// into its store, then contiguous L2-point pass with a transposed store).
//
// Convention (reference): position -> momentum on a ket index = np.fft.fft(norm="ortho"),
// momentum -> position = np.fft.ifft(norm="ortho")  (tests/test_operator.py:197-204).
#include <math.h>

#include "sqb_common.cuh"

namespace {

constexpr int kFftThreads = 256;
constexpr int kMaxSmemLen = 4096;       // complex elements of one ping-pong buffer
constexpr int kMaxStages = 24;
constexpr int kMaxRadix = 31;

struct FftPass {
  int len;                 // transform length of this pass
  int n_stages;
  int radix[kMaxStages];
  int64_t n_outer, n_inner;  // lines = n_outer * n_inner
  int64_t in_so, in_sl, in_si;     // element strides of (outer, position-in-line, inner) on input
  int64_t out_so, out_sl, out_si;  // ... on output
  int lines_per_cta;
  int inner_fastest;       // load order: 1 = a CTA's lines are adjacent in memory (coalesce across lines)
  int out_inner_fastest;   // same for the store
  int sign;                // -1 forward, +1 backward
  double scale;            // multiplied into every output
  int64_t tw_len;          // 0 = none; else multiply output (k, inner=i) by exp(sign*2*pi*i*k*i/tw_len)
};

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
#pragma unroll
  for (int k = 0; k < R; ++k) {
    c128 acc = x[0];
#pragma unroll
    for (int t = 1; t < R; ++t) cfma(acc, x[t], w[(t * k) % R == 0 ? 1 : (t * k) % R]);
    y[k] = acc;
  }
  // k = 0 row must use w^0 = 1 (the ternary above only dodges an out-of-range index at compile time)
  {
    c128 acc = x[0];
#pragma unroll
    for (int t = 1; t < R; ++t) acc = cadd(acc, x[t]);
    y[0] = acc;
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

// multiply by -i*sign... helper: rotate by exp(sign * i*pi/2) = sign*i
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

// Generic small DFT of prime (or any) radix r using the root table: roots[j] = exp(-2 pi i j / r).
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

template <int R>
__device__ __forceinline__ void stage_fixed(const c128* src, c128* dst, int len, int ns, int lines,
                                            const c128* tab, int sign) {
  const int per_line = len / R;
  const int work = per_line * lines;
  const int tw_step = len / (ns * R);
  for (int w = threadIdx.x; w < work; w += kFftThreads) {
    const int line = w / per_line;
    const int j = w - line * per_line;
    const int k = j % ns;
    c128 x[R];
#pragma unroll
    for (int t = 0; t < R; ++t) x[t] = src[line * len + j + t * per_line];
#pragma unroll
    for (int t = 1; t < R; ++t) x[t] = cmul(x[t], tw_lookup(tab, t * k * tw_step, sign));
    dft_small<R>(x, sign, tab, len / R);
    const int base = line * len + (j - k) * R + k;
#pragma unroll
    for (int t = 0; t < R; ++t) dst[base + t * ns] = x[t];
  }
}

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

__global__ void __launch_bounds__(kFftThreads)
fft_pass_kernel(FftPass P, const c128* __restrict__ in, c128* __restrict__ out, const c128* __restrict__ tab_g) {
  extern __shared__ c128 smem[];
  const int len = P.len;
  const int lpc = P.lines_per_cta;
  c128* buf0 = smem;
  c128* buf1 = smem + (size_t)lpc * len;
  c128* tab = buf1 + (size_t)lpc * len;  // [len]

  const int64_t n_lines = P.n_outer * P.n_inner;
  const int64_t line0 = (int64_t)blockIdx.x * lpc;
  const int lines = (int)sqb_min64(lpc, n_lines - line0);
  if (lines <= 0) return;

  for (int j = threadIdx.x; j < len; j += kFftThreads) tab[j] = tab_g[j];

  // ---- load
  const int total = lines * len;
  if (P.inner_fastest) {
    for (int e = threadIdx.x; e < total; e += kFftThreads) {
      const int pos = e / lines, l = e - pos * lines;
      const int64_t line = line0 + l;
      const int64_t o = line / P.n_inner, i = line - o * P.n_inner;
      buf0[l * len + pos] = in[o * P.in_so + pos * P.in_sl + i * P.in_si];
    }
  } else {
    for (int e = threadIdx.x; e < total; e += kFftThreads) {
      const int l = e / len, pos = e - l * len;
      const int64_t line = line0 + l;
      const int64_t o = line / P.n_inner, i = line - o * P.n_inner;
      buf0[l * len + pos] = in[o * P.in_so + pos * P.in_sl + i * P.in_si];
    }
  }
  __syncthreads();

  // ---- stages
  c128* src = buf0;
  c128* dst = buf1;
  int ns = 1;
  for (int s = 0; s < P.n_stages; ++s) {
    const int r = P.radix[s];
    switch (r) {
      case 2: stage_fixed<2>(src, dst, len, ns, lines, tab, P.sign); break;
      case 4: stage_fixed<4>(src, dst, len, ns, lines, tab, P.sign); break;
      case 3: stage_fixed<3>(src, dst, len, ns, lines, tab, P.sign); break;
      case 5: stage_fixed<5>(src, dst, len, ns, lines, tab, P.sign); break;
      case 7: stage_fixed<7>(src, dst, len, ns, lines, tab, P.sign); break;
      default: stage_generic(src, dst, len, ns, r, lines, tab, P.sign); break;
    }
    __syncthreads();
    c128* t = src; src = dst; dst = t;
    ns *= r;
  }

  // ---- store (optionally with the four-step twiddle)
  if (P.out_inner_fastest) {
    for (int e = threadIdx.x; e < total; e += kFftThreads) {
      const int pos = e / lines, l = e - pos * lines;
      const int64_t line = line0 + l;
      const int64_t o = line / P.n_inner, i = line - o * P.n_inner;
      c128 v = cscale(src[l * len + pos], P.scale);
      if (P.tw_len) {
        double sn, cs;
        const int64_t prod = ((int64_t)pos * i) % P.tw_len;
        sincospi((double)P.sign * 2.0 * (double)prod / (double)P.tw_len, &sn, &cs);
        v = cmul(v, make_double2(cs, sn));
      }
      out[o * P.out_so + pos * P.out_sl + i * P.out_si] = v;
    }
  } else {
    for (int e = threadIdx.x; e < total; e += kFftThreads) {
      const int l = e / len, pos = e - l * len;
      const int64_t line = line0 + l;
      const int64_t o = line / P.n_inner, i = line - o * P.n_inner;
      c128 v = cscale(src[l * len + pos], P.scale);
      if (P.tw_len) {
        double sn, cs;
        const int64_t prod = ((int64_t)pos * i) % P.tw_len;
        sincospi((double)P.sign * 2.0 * (double)prod / (double)P.tw_len, &sn, &cs);
        v = cmul(v, make_double2(cs, sn));
      }
      out[o * P.out_so + pos * P.out_sl + i * P.out_si] = v;
    }
  }
}

// ------------------------------------------------------------------------------------------ host side
bool factorize(int len, int* radix, int* n_stages) {
  int n = 0, rem = len;
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

// workspace layout: [twiddle table: max_len c128][scratch: total elements c128 (only for long lines)]
struct AxisPlan {
  bool long_line;
  int l1, l2;
};

bool plan_axis(int len, AxisPlan* ap) {
  int radix[kMaxStages], ns;
  if (len <= kMaxSmemLen) {
    ap->long_line = false;
    ap->l1 = len;
    ap->l2 = 1;
    return factorize(len, radix, &ns);
  }
  ap->long_line = true;
  if (!split_len(len, &ap->l1, &ap->l2)) return false;
  return factorize(ap->l1, radix, &ns) && factorize(ap->l2, radix, &ns);
}

int launch_pass(FftPass& P, const c128* in, c128* out, c128* tab, cudaStream_t st) {
  if (!factorize(P.len, P.radix, &P.n_stages)) return SQB_E_UNSUPPORTED;
  int lpc = kMaxSmemLen / P.len;
  if (lpc < 1) lpc = 1;
  if (lpc > 16) lpc = 16;
  const int64_t n_lines = P.n_outer * P.n_inner;
  if (P.inner_fastest) {
    // a CTA's lines must share `outer`: make lines_per_cta divide n_inner
    while (lpc > 1 && (P.n_inner % lpc)) --lpc;
  }
  if (lpc > n_lines) lpc = (int)n_lines;
  P.lines_per_cta = lpc;
  const size_t smem = ((size_t)2 * lpc * P.len + P.len) * sizeof(c128);
  cudaError_t e = cudaFuncSetAttribute(fft_pass_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  if (e != cudaSuccess) return (int)e;
  fft_table_kernel<<<(P.len + 255) / 256, 256, 0, st>>>(tab, P.len);
  const int64_t grid = (n_lines + lpc - 1) / lpc;
  if (grid > 2147483647LL) return SQB_E_UNSUPPORTED;
  fft_pass_kernel<<<(unsigned)grid, kFftThreads, smem, st>>>(P, in, out, tab);
  SQB_LAUNCH_CHECK();
  return SQB_OK;
}

// One axis of a [outer, len, inner] array, in -> out (in == out allowed when !long_line).
int fft_one_axis(int64_t outer, int len, int64_t inner, int sign, double scale, const c128* in, c128* out,
                 c128* tab, c128* scratch, cudaStream_t st) {
  AxisPlan ap;
  if (!plan_axis(len, &ap)) return SQB_E_UNSUPPORTED;
  if (!ap.long_line) {
    FftPass P{};
    P.len = len;
    P.n_outer = outer;
    P.n_inner = inner;
    P.in_so = P.out_so = (int64_t)len * inner;
    P.in_sl = P.out_sl = inner;
    P.in_si = P.out_si = 1;
    P.inner_fastest = P.out_inner_fastest = inner > 1;
    P.sign = sign;
    P.scale = scale;
    P.tw_len = 0;
    return launch_pass(P, in, out, tab, st);
  }
  // four-step: view the line as [l1, l2] (n = n1*l2 + n2).
  // pass A: l1-point transforms over n1 for every (outer, n2, inner), twiddle W_len^(k1*n2), in -> scratch
  // pass B: l2-point transforms over n2 for every (outer, k1, inner), output index k1 + l1*k2, scratch -> out
  const int l1 = ap.l1, l2 = ap.l2;
  if (inner != 1) {
    // treat (n2, inner) as the combined inner index for pass A; twiddle needs n2 alone -> only inner == 1 here
    return SQB_E_UNSUPPORTED;
  }
  {
    FftPass P{};
    P.len = l1;
    P.n_outer = outer;
    P.n_inner = l2;
    P.in_so = P.out_so = len;
    P.in_sl = P.out_sl = l2;
    P.in_si = P.out_si = 1;
    P.inner_fastest = P.out_inner_fastest = 1;
    P.sign = sign;
    P.scale = 1.0;
    P.tw_len = len;
    int rc = launch_pass(P, in, scratch, tab, st);
    if (rc) return rc;
  }
  {
    FftPass P{};
    P.len = l2;
    P.n_outer = outer;
    P.n_inner = l1;  // "inner" index = k1, but lines are contiguous in n2: strides below
    P.in_so = len;
    P.in_sl = 1;
    P.in_si = l2;
    P.out_so = len;
    P.out_sl = l1;
    P.out_si = 1;
    P.inner_fastest = 0;
    P.out_inner_fastest = 1;
    P.sign = sign;
    P.scale = scale;
    P.tw_len = 0;
    return launch_pass(P, scratch, out, tab, st);
  }
}

}  // namespace

extern "C" size_t sqb_fft_workspace_bytes(int nd, const int* dims_host, int64_t batch) {
  if (nd < 1 || nd > SQB_MAX_DIMS || !dims_host || batch < 0) return 0;
  int64_t total = batch;
  int max_len = 1;
  bool any_long = false;
  for (int a = 0; a < nd; ++a) {
    total *= dims_host[a];
    if (dims_host[a] > max_len) max_len = dims_host[a];
    if (dims_host[a] > kMaxSmemLen) any_long = true;
  }
  size_t bytes = align_up((size_t)max_len * sizeof(c128), 256);
  if (any_long) bytes += align_up((size_t)total * sizeof(c128), 256);
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
                                          align_up((size_t)max_len * sizeof(c128), 256));
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
                                          align_up((size_t)len * sizeof(c128), 256));
  return fft_one_axis(outer, len, inner, sign, 1.0 / sqrt((double)len), reinterpret_cast<const c128*>(in),
                      reinterpret_cast<c128*>(out), tab, scratch, sqb_stream(stream));
}
