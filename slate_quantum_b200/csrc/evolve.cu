// K3: eigenbasis projection, phase evolution and back-transformation.
//
//   sqb_project       c[b,e] = sum_k conj(V_b[e,k]) psi[b,k]                  (HBM bound: reads V once)
//   sqb_evolve_eigen  a[t,i] = c[i] exp(-1j * w[i] * t / hbar)                (materialises the reference's
//                                                                              return value, schrodinger.py:51-53)
//   sqb_evolve_bloch  out[t,b,k] = sum_e (c[b,e] exp(-1j w[b,e] t/hbar)) V_b[e,k]
//                     batched complex GEMM [T x n] x [n x n] per k-block on the FP64 tensor cores
//                     (DMMA.8x8x4 through mma.sync.m8n8k4.f64; 4 real MMAs per complex MAC tile).
//                     The A operand (phases) is GENERATED in shared memory by the CTA - the [T, M]
//                     eigenbasis array is never written or read - and the B operand (eigenvectors) is
//                     staged by 1-D bulk TMA copies (cp.async.bulk + mbarrier), double buffered.
//
// Reference: slate_quantum/dynamics/schrodinger.py:31-56; ExplicitUnitaryBasis conversions reached from
// state/_state.py:37-43 (projection) and :153-165 (back-transform).
#include <math.h>
#include <stdlib.h>

#include "sqb_common.cuh"

namespace {

// ------------------------------------------------------------------------------------------------
// projection: one warp per (b, e) row
__global__ void __launch_bounds__(256)
project_kernel(int n, int64_t rows, const c128* __restrict__ V, const c128* __restrict__ psi, c128* __restrict__ c) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int64_t b = row / n;
  const c128* v = V + row * n;
  const c128* p = psi + b * n;
  c128 acc = make_double2(0.0, 0.0);
  for (int k = lane; k < n; k += 32) cfmac(acc, v[k], p[k]);
  acc.x = warp_sum(acc.x);
  acc.y = warp_sum(acc.y);
  if (lane == 0) c[row] = acc;
}

// generic explicit-basis conversions for [lead, batch*n] vector lists (not the fused hot path):
//   mode 0 (explicit -> inner): out[l,b,k] = sum_e in[l,b,e] V[b,e,k]
//   mode 1 (inner -> explicit): out[l,b,e] = sum_k conj(V[b,e,k]) in[l,b,k]
__global__ void __launch_bounds__(256)
apply_transform_kernel(int n, int64_t batch, int64_t lead, int mode, const c128* __restrict__ V,
                       const c128* __restrict__ in, c128* __restrict__ out) {
  if (mode == 1) {
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);  // (l, b, e)
    if (row >= lead * batch * n) return;
    const int64_t be = row % (batch * n), l = row / (batch * n);
    const int64_t b = be / n;
    const c128* v = V + be * n;
    const c128* p = in + (l * batch + b) * n;
    c128 acc = make_double2(0.0, 0.0);
    for (int k = lane; k < n; k += 32) cfmac(acc, v[k], p[k]);
    acc.x = warp_sum(acc.x);
    acc.y = warp_sum(acc.y);
    if (lane == 0) out[row] = acc;
  } else {
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // (l, b, k)
    if (g >= lead * batch * n) return;
    const int k = (int)(g % n);
    const int64_t lb = g / n;
    const int64_t b = lb % batch;
    const c128* v = V + b * (int64_t)n * n + k;
    const c128* a = in + lb * n;
    c128 acc = make_double2(0.0, 0.0);
    for (int e = 0; e < n; ++e) cfma(acc, a[e], v[(int64_t)e * n]);
    out[g] = acc;
  }
}

// literal phase of dynamics/schrodinger.py:51-53: exp(-1j * E * t / hbar) with E real-valued complex128
__device__ __forceinline__ c128 phase_factor(double w, double t, double hbar) {
  const double theta = -((w * t) / hbar);
  double s, c;
  sincos(theta, &s, &c);
  return make_double2(c, s);
}

__global__ void __launch_bounds__(256)
evolve_eigen_kernel(int64_t m, const double* __restrict__ w, const c128* __restrict__ c, const double* __restrict__ times,
                    int64_t n_times, double hbar, c128* __restrict__ out) {
  const int64_t total = m * n_times;
  for (int64_t g = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; g < total; g += (int64_t)gridDim.x * blockDim.x) {
    const int64_t t = g / m, i = g - t * m;
    out[g] = cmul(c[i], phase_factor(w[i], times[t], hbar));
  }
}

constexpr int TM = 64;    // times per CTA tile (both GEMM variants)

// Phase tables of a uniform time grid t_j = times[0] + j dt (per eigenvalue i, m = batch * n of them):
//   step8[i]       = exp(-1j w_i 8 dt / hbar)              factor between time rows 8 apart
//   pow8[i][g]     = exp(-1j w_i g dt / hbar), g = 0..7    offset of row g inside a group of 8
//   tilebase[q][i] = exp(-1j w_i times[64 q] / hbar)       first row of CTA time tile q
// so that a producer thread of the GEMM builds its 8 A entries with complex multiplications only.
__global__ void __launch_bounds__(256)
phase_tables_kernel(int64_t m, const double* __restrict__ w, const double* __restrict__ times, int64_t t_count,
                    double dt, double hbar, c128* __restrict__ step8, c128* __restrict__ pow8,
                    c128* __restrict__ tilebase, int* __restrict__ plane_valid_flag) {
  if (blockIdx.x == 0 && threadIdx.x == 0) *plane_valid_flag = 1;  // ordered after bsum_kernel on the stream
  const int64_t n_tiles = (t_count + TM - 1) / TM;
  const int64_t total = m * (9 + n_tiles);
  for (int64_t g = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; g < total; g += (int64_t)gridDim.x * blockDim.x) {
    const int64_t k = g / m, i = g - k * m;  // i fastest: coalesced
    const double wi = w[i];
    if (k == 0) {
      step8[i] = phase_factor(wi, 8.0 * dt, hbar);
    } else if (k < 9) {
      pow8[i * 8 + (k - 1)] = phase_factor(wi, (double)(k - 1) * dt, hbar);
    } else {
      tilebase[(k - 9) * m + i] = phase_factor(wi, times[(k - 9) * TM], hbar);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// fused phase x eigenvector GEMM, warp specialised:
//   warps 0..7   consumers: 32 x 32 complex warp tiles (2 x 4 over the 64 x 128 CTA tile), DMMA.8x8x4 only
//   warps 8..11  producers: generate the A chunk (c_e * exp(-i w_e t / hbar), one sincos per element) into
//                shared memory and issue the 1-D bulk (TMA) copies of the B chunk (eigenvector rows)
// kGemmStages-deep ring, full/empty mbarriers; the tensor pipe never waits for sincos or for L2.
constexpr int TN = 128;   // components per CTA tile
constexpr int KC = 16;    // eigen-index chunk
constexpr int LDA = KC + 4;   // complex elements; == 4 (mod 8) -> conflict-free 16 B fragment loads
constexpr int LDB = TN + 2;   // == 2 (mod 8)
constexpr int kGemmStages = 3;
constexpr int kConsumerThreads = 256;
constexpr int kProducerThreads = 128;
constexpr int kGemmThreads = kConsumerThreads + kProducerThreads;
constexpr int kStageElems = TM * LDA + KC * LDB;
constexpr size_t kGemmSmem = (size_t)kGemmStages * kStageElems * sizeof(c128) + 128;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.expect_tx.relaxed.cta.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}

__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(d0), "+d"(d1)
               : "d"(a), "d"(b));
}

__global__ void __launch_bounds__(kGemmThreads, 1)
evolve_bloch_kernel(int n, const c128* __restrict__ V, const double* __restrict__ w, const c128* __restrict__ coef,
                    const double* __restrict__ times, int64_t t_count, double hbar, c128* __restrict__ out,
                    int64_t out_row_stride, const c128* __restrict__ step8) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  c128* stage0 = reinterpret_cast<c128*>(smem_raw);  // [stages][ A: TM x LDA | B: KC x LDB ]
  uint64_t* full = reinterpret_cast<uint64_t*>(stage0 + kGemmStages * kStageElems);  // [stages]
  uint64_t* empty = full + kGemmStages;                                               // [stages]

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int p0 = blockIdx.x * TN;
  const int64_t t0 = (int64_t)blockIdx.y * TM;
  const int64_t b = blockIdx.z;
  const c128* Vb = V + b * (int64_t)n * n;
  const int ncols = min(TN, n - p0);
  const int nchunks = (n + KC - 1) / KC;

  if (tid == 0) {
    for (int s = 0; s < kGemmStages; ++s) {
      mbar_init(&full[s], kProducerThreads);
      mbar_init(&empty[s], kConsumerThreads / 32);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  if (warp >= kConsumerThreads / 32) {
    // ============================== producers ==============================
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
    const int ptid = tid - kConsumerThreads;
    const int ge = ptid % KC;   // eigen index inside the chunk
    const int gj = ptid / KC;   // 0..7 ; rows gj + 8 q
    const double* wb = w + b * n;
    const c128* cb = coef + b * n;
    double tj[TM / 8];
#pragma unroll
    for (int q = 0; q < TM / 8; ++q) {
      const int64_t t = t0 + gj + 8 * q;
      tj[q] = t < t_count ? times[t] : 0.0;
    }
    // uniform time grids: rows gj, gj+8, gj+16, .. differ by the constant factor step8[e] = exp(-i w_e 8 dt / hbar)
    // (precomputed per eigenvalue), so a thread needs ONE sincos per chunk instead of TM/8.
    const c128* sb = step8 ? step8 + b * n : nullptr;
    double we_next = 0.0;
    c128 ce_next = make_double2(0.0, 0.0), se_next = make_double2(1.0, 0.0);
    if (ge < n) {
      we_next = wb[ge];
      ce_next = cb[ge];
      if (sb) se_next = sb[ge];
    }
    for (int kc = 0; kc < nchunks; ++kc) {
      const int s = kc % kGemmStages;
      const uint32_t ph = (uint32_t)((kc / kGemmStages) & 1);
      c128* As = stage0 + s * kStageElems;
      c128* Bs = As + TM * LDA;
      const double we = we_next;
      const c128 ce = ce_next, se = se_next;
      const int e = kc * KC + ge;
      {  // prefetch the next chunk's eigenvalue / coefficient
        const int en = e + KC;
        we_next = 0.0;
        ce_next = make_double2(0.0, 0.0);
        if (en < n) {
          we_next = wb[en];
          ce_next = cb[en];
          if (sb) se_next = sb[en];
        }
      }
      mbar_wait(&empty[s], ph ^ 1u);
      const int e0 = kc * KC;
      const int nrows = min(KC, n - e0);
      if (ptid == 0) {
        mbar_expect_tx(&full[s], (uint32_t)(nrows * ncols * sizeof(c128)));
        for (int r = 0; r < nrows; ++r)
          bulk_g2s(Bs + r * LDB, Vb + (int64_t)(e0 + r) * n + p0, (uint32_t)(ncols * sizeof(c128)), &full[s]);
      }
      if (nrows < KC) {
        for (int idx = ptid; idx < (KC - nrows) * LDB; idx += kProducerThreads)
          Bs[nrows * LDB + idx] = make_double2(0.0, 0.0);
      }
      if (sb) {
        c128 v = make_double2(0.0, 0.0);
        if (e < n) v = cmul(ce, phase_factor(we, tj[0], hbar));
#pragma unroll
        for (int q = 0; q < TM / 8; ++q) {
          const int j = gj + 8 * q;
          As[j * LDA + ge] = (t0 + j < t_count) ? v : make_double2(0.0, 0.0);
          v = cmul(v, se);
        }
      } else {
#pragma unroll
        for (int q = 0; q < TM / 8; ++q) {
          const int j = gj + 8 * q;
          c128 v = make_double2(0.0, 0.0);
          if (e < n && t0 + j < t_count) v = cmul(ce, phase_factor(we, tj[q], hbar));
          As[j * LDA + ge] = v;
        }
      }
      mbar_arrive(&full[s]);
    }
    return;
  }

  // ============================== consumers ==============================
  asm volatile("setmaxnreg.inc.sync.aligned.u32 232;");
  const int wm = (warp & 1) * 32;
  const int wn = (warp >> 1) * 32;
  double cr[4][4][2], ci[4][4][2];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      cr[a][c][0] = cr[a][c][1] = 0.0;
      ci[a][c][0] = ci[a][c][1] = 0.0;
    }
  const int frow = lane >> 2, fcol = lane & 3;
  for (int kc = 0; kc < nchunks; ++kc) {
    const int s = kc % kGemmStages;
    const uint32_t ph = (uint32_t)((kc / kGemmStages) & 1);
    const c128* Asb = stage0 + s * kStageElems;
    const c128* Bsb = Asb + TM * LDA;
    mbar_wait(&full[s], ph);
    // fragments are double buffered in registers; the two products of a complex MAC that hit the same
    // accumulator are issued a full sweep (32 DMMAs) apart so no DMMA waits on its predecessor
    c128 af[2][4], bf[2][4];
#pragma unroll
    for (int a = 0; a < 4; ++a) af[0][a] = Asb[(wm + a * 8 + frow) * LDA + fcol];
#pragma unroll
    for (int c = 0; c < 4; ++c) bf[0][c] = Bsb[fcol * LDB + wn + c * 8 + frow];
#pragma unroll
    for (int k4 = 0; k4 < KC; k4 += 4) {
      const int cur = (k4 >> 2) & 1, nxt = cur ^ 1;
      if (k4 + 4 < KC) {
#pragma unroll
        for (int a = 0; a < 4; ++a) af[nxt][a] = Asb[(wm + a * 8 + frow) * LDA + k4 + 4 + fcol];
#pragma unroll
        for (int c = 0; c < 4; ++c) bf[nxt][c] = Bsb[(k4 + 4 + fcol) * LDB + wn + c * 8 + frow];
      }
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          dmma(cr[a][c][0], cr[a][c][1], af[cur][a].x, bf[cur][c].x);
          dmma(ci[a][c][0], ci[a][c][1], af[cur][a].x, bf[cur][c].y);
        }
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        const double nai = -af[cur][a].y;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          dmma(cr[a][c][0], cr[a][c][1], nai, bf[cur][c].y);
          dmma(ci[a][c][0], ci[a][c][1], af[cur][a].y, bf[cur][c].x);
        }
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[s]);
  }

  // epilogue: lane holds C[row = frow][cols 2*fcol, 2*fcol+1] of each 8x8 tile
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const int64_t t = t0 + wm + a * 8 + frow;
    if (t >= t_count) continue;
    c128* orow = out + t * out_row_stride + b * (int64_t)n;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int p = p0 + wn + c * 8 + fcol * 2;
      if (p < n) orow[p] = make_double2(cr[a][c][0], ci[a][c][0]);
      if (p + 1 < n) orow[p + 1] = make_double2(cr[a][c][1], ci[a][c][1]);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// 3M variant of the fused GEMM (uniform time grids): a complex product needs THREE real products
//   P1 = Ar Br, P2 = Ai Bi, P3 = (Ar + Ai)(Br + Bi);   Cr = P1 - P2,  Ci = P3 - P1 - P2
// i.e. 3 DMMA.8x8x4 per complex 8x8x4 tile instead of 4 (25 % less tensor work; normwise stable, error
// ~1e-15 relative to |a||b|, far inside the 1e-9 tolerance).  Three accumulator planes cost 1.5x the
// registers, so the warp tile is 32 x 16 and the CTA tile 64 (t) x 64 (k); the producers also store the
// plane Ar + Ai, and the plane Br + Bi of the eigenvectors is precomputed once per call (bsum_kernel, row
// stride padded to an even number of doubles so that its rows can be bulk-copied) - a DADD in the consumer
// warps would queue behind their own DMMAs on the shared FP64 pipe.
__global__ void __launch_bounds__(256)
bsum_kernel(int n, int n_pad, int64_t rows, const c128* __restrict__ V, double* __restrict__ bsum,
            const int* __restrict__ valid_flag, int force) {
  // reuse requested by the caller: skip unless the NUFFT route ran on a new transform since the plane was built
  if (!force && *valid_flag == 1) return;
  const int64_t total = rows * n_pad;
  for (int64_t g = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; g < total; g += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = g / n_pad;
    const int k = (int)(g - r * n_pad);
    double v = 0.0;
    if (k < n) {
      const c128 z = V[r * n + k];
      v = z.x + z.y;
    }
    bsum[g] = v;
  }
}

constexpr int TN3 = 64;
constexpr int LDB3 = TN3 + 2;   // == 2 (mod 8)
constexpr int LDAS = KC + 4;    // doubles; == 4 (mod 16) -> conflict-free 8 B fragment loads
constexpr int LDBS = TN3 + 4;    // doubles; == 4 (mod 16)
constexpr int kStages3 = 4;
constexpr int kStage3Bytes = TM * LDA * 16 + TM * LDAS * 8 + KC * LDB3 * 16 + KC * LDBS * 8;
constexpr size_t kGemm3Smem = (size_t)kStages3 * kStage3Bytes + 128;

__global__ void __launch_bounds__(kGemmThreads, 1)
evolve_bloch_3m_kernel(int n, const c128* __restrict__ V, const c128* __restrict__ coef, int64_t t_count,
                       c128* __restrict__ out, int64_t out_row_stride, const c128* __restrict__ step8,
                       const c128* __restrict__ pow8, const c128* __restrict__ tilebase, int64_t m_total,
                       const double* __restrict__ bsum, int n_pad, int tiles_per_cta) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + (size_t)kStages3 * kStage3Bytes);  // [stages]
  uint64_t* empty = full + kStages3;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int p0 = blockIdx.x * TN3;
  // a CTA walks tiles_per_cta consecutive time tiles of its (block, column tile): the ring keeps flowing across
  // tiles, so the producers' start-up latency and the consumers' epilogue overlap with the neighbouring tile
  // (with one tile per CTA and one CTA per SM they were exposed: ~12 ms per step)
  const int64_t tile_first = (int64_t)blockIdx.y * tiles_per_cta;
  const int64_t n_t_tiles = (t_count + TM - 1) / TM;
  const int my_tiles = (int)sqb_min64(tiles_per_cta, n_t_tiles - tile_first);
  const int64_t b = blockIdx.z;
  const c128* Vb = V + b * (int64_t)n * n;
  const int ncols = min(TN3, n - p0);
  const int nchunks = (n + KC - 1) / KC;

  if (tid == 0) {
    for (int s = 0; s < kStages3; ++s) {
      mbar_init(&full[s], 32);
      mbar_init(&empty[s], kConsumerThreads / 32);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  // Producers are the lowest warp ids and each producer warp OWNS one ring stage: warp w generates the whole A
  // chunk (and issues the TMA copies) of chunks w, w+4, ...  ncu showed that the FP64 pipe is handed over only
  // when its current owner stalls: a producer waited ~1700 cycles for the consumers' DMMA burst to end before
  // its own short burst could start, and with all four producer warps contributing to EVERY chunk the
  // consumers waited for the slowest of four such hand-overs per chunk (21 % of their time on the full
  // barrier).  With one warp per stage a producer needs one hand-over per FOUR chunk times.
  static_assert(kProducerThreads / 32 == kStages3, "one producer warp per ring stage");
  if (warp < kProducerThreads / 32) {
    // ============================== producers ==============================
    const int s = warp;
    const int ge = lane & 15;          // eigen index inside the chunk
    const int gh = (lane >> 4) * 4;    // rows gh + g + 8 q, g < 4, q < 8
    // A[t0 + j + 8q, e] = c_e * tilebase[tile][e] * pow8[e][j] * step8[e]^q : multiplications only
    const c128* cb = coef + b * n;
    const c128* sb = step8 + b * n;
    const c128* pb = pow8 + b * n * 8 + gh;
    const c128* tb0 = tilebase + tile_first * m_total + b * n;
    unsigned char* st = smem_raw + (size_t)s * kStage3Bytes;
    c128* Ac = reinterpret_cast<c128*>(st);
    double* As = reinterpret_cast<double*>(st + TM * LDA * 16);
    c128* Bs = reinterpret_cast<c128*>(st + TM * LDA * 16 + TM * LDAS * 8);
    double* Bq = reinterpret_cast<double*>(st + TM * LDA * 16 + TM * LDAS * 8 + KC * LDB3 * 16);
    const int total_chunks = my_tiles * nchunks;
    for (int kg = s; kg < total_chunks; kg += kStages3) {
      const uint32_t ph = (uint32_t)((kg / kStages3) & 1);
      const int tile = kg / nchunks;
      const int kc = kg - tile * nchunks;
      const int64_t t0 = (tile_first + tile) * TM;
      const c128* tb = tb0 + (int64_t)tile * m_total;
      const int e = kc * KC + ge;
      c128 base = make_double2(0.0, 0.0), se = make_double2(1.0, 0.0);
      c128 pw[4];
#pragma unroll
      for (int g = 0; g < 4; ++g) pw[g] = make_double2(0.0, 0.0);
      if (e < n) {
        const c128 c_raw = cb[e], t_raw = tb[e];
        se = sb[e];
#pragma unroll
        for (int g = 0; g < 4; ++g) pw[g] = pb[(int64_t)e * 8 + g];
        base = cmul(c_raw, t_raw);
      }
      mbar_wait(&empty[s], ph ^ 1u);
      const int e0 = kc * KC;
      const int nrows = min(KC, n - e0);
      {
        // one bulk copy pair per eigenvector row, issued by 16 lanes at once (a single lane looping over the
        // 32 copies of a chunk cost ~8 ms per step); the expected byte count is registered first
        const int ncols2 = (ncols + 1) & ~1;  // bulk copies move multiples of 16 B; the pad column is zero
        if (lane == 0)
          mbar_expect_tx(&full[s], (uint32_t)(nrows * (ncols * sizeof(c128) + ncols2 * sizeof(double))));
        __syncwarp();
        if (lane < nrows) {
          const double* qb = bsum + (b * n + e0 + lane) * (int64_t)n_pad + p0;
          bulk_g2s(Bs + lane * LDB3, Vb + (int64_t)(e0 + lane) * n + p0, (uint32_t)(ncols * sizeof(c128)), &full[s]);
          bulk_g2s(Bq + lane * LDBS, qb, (uint32_t)(ncols2 * sizeof(double)), &full[s]);
        }
      }
      if (nrows < KC) {
        for (int idx = lane; idx < (KC - nrows) * LDB3; idx += 32) Bs[nrows * LDB3 + idx] = make_double2(0.0, 0.0);
        for (int idx = lane; idx < (KC - nrows) * LDBS; idx += 32) Bq[nrows * LDBS + idx] = 0.0;
      }
      // powers of the step factor by squaring: the 8 rows of a column are then products of depth <= 3 instead
      // of an 8-deep recurrence (the FP64 pipe is held for the whole latency-bound burst, see above)
      const c128 se2 = cmul(se, se);
      const c128 se4 = cmul(se2, se2);
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        c128 v[TM / 8];
        v[0] = cmul(base, pw[g]);  // zero for e >= n
        v[1] = cmul(v[0], se);
        v[2] = cmul(v[0], se2);
        v[3] = cmul(v[1], se2);
#pragma unroll
        for (int q = 0; q < 4; ++q) v[4 + q] = cmul(v[q], se4);
#pragma unroll
        for (int q = 0; q < TM / 8; ++q) {
          const int j = gh + g + 8 * q;
          const c128 o = (t0 + j < t_count) ? v[q] : make_double2(0.0, 0.0);
          Ac[j * LDA + ge] = o;
          As[j * LDAS + ge] = o.x + o.y;
        }
      }
      mbar_arrive(&full[s]);
    }
    return;
  }

  // ============================== consumers ==============================
  const int cwarp = warp - kProducerThreads / 32;
  const int wm = (cwarp & 1) * 32;
  const int wn = (cwarp >> 1) * 16;
  const int frow = lane >> 2, fcol = lane & 3;
  int kg = 0;
  for (int tile = 0; tile < my_tiles; ++tile) {
  const int64_t t0 = (tile_first + tile) * TM;
  double p1[4][2][2], p2[4][2][2], p3[4][2][2];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      p1[a][c][0] = p1[a][c][1] = 0.0;
      p2[a][c][0] = p2[a][c][1] = 0.0;
      p3[a][c][0] = p3[a][c][1] = 0.0;
    }
  for (int kc = 0; kc < nchunks; ++kc, ++kg) {
    const int s = kg % kStages3;
    const uint32_t ph = (uint32_t)((kg / kStages3) & 1);
    const unsigned char* st = smem_raw + (size_t)s * kStage3Bytes;
    const c128* Ac = reinterpret_cast<const c128*>(st);
    const double* As = reinterpret_cast<const double*>(st + TM * LDA * 16);
    const c128* Bs = reinterpret_cast<const c128*>(st + TM * LDA * 16 + TM * LDAS * 8);
    const double* Bq = reinterpret_cast<const double*>(st + TM * LDA * 16 + TM * LDAS * 8 + KC * LDB3 * 16);
    mbar_wait(&full[s], ph);
    c128 af[2][4], bf[2][2];
    double as[2][4], bs[2][2];
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      af[0][a] = Ac[(wm + a * 8 + frow) * LDA + fcol];
      as[0][a] = As[(wm + a * 8 + frow) * LDAS + fcol];
    }
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      bf[0][c] = Bs[fcol * LDB3 + wn + c * 8 + frow];
      bs[0][c] = Bq[fcol * LDBS + wn + c * 8 + frow];
    }
#pragma unroll
    for (int k4 = 0; k4 < KC; k4 += 4) {
      const int cur = (k4 >> 2) & 1, nxt = cur ^ 1;
      if (k4 + 4 < KC) {
#pragma unroll
        for (int a = 0; a < 4; ++a) {
          af[nxt][a] = Ac[(wm + a * 8 + frow) * LDA + k4 + 4 + fcol];
          as[nxt][a] = As[(wm + a * 8 + frow) * LDAS + k4 + 4 + fcol];
        }
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          bf[nxt][c] = Bs[(k4 + 4 + fcol) * LDB3 + wn + c * 8 + frow];
          bs[nxt][c] = Bq[(k4 + 4 + fcol) * LDBS + wn + c * 8 + frow];
        }
      }
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          dmma(p1[a][c][0], p1[a][c][1], af[cur][a].x, bf[cur][c].x);
          dmma(p2[a][c][0], p2[a][c][1], af[cur][a].y, bf[cur][c].y);
          dmma(p3[a][c][0], p3[a][c][1], as[cur][a], bs[cur][c]);
        }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[s]);
  }

#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const int64_t t = t0 + wm + a * 8 + frow;
    if (t >= t_count) continue;
    c128* orow = out + t * out_row_stride + b * (int64_t)n;
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      const int p = p0 + wn + c * 8 + fcol * 2;
      if (p < n) orow[p] = make_double2(p1[a][c][0] - p2[a][c][0], p3[a][c][0] - p1[a][c][0] - p2[a][c][0]);
      if (p + 1 < n) orow[p + 1] = make_double2(p1[a][c][1] - p2[a][c][1], p3[a][c][1] - p1[a][c][1] - p2[a][c][1]);
    }
  }
  }  // tiles
}

}  // namespace

#include "evolve_nufft.cuh"

extern "C" int sqb_project(int n, int64_t batch, const sqb_c128* transform, const sqb_c128* psi, sqb_c128* c,
                           void* stream) {
  if (n < 1 || batch < 0 || !transform || !psi || !c) return SQB_E_BADARG;
  if (batch == 0) return SQB_OK;
  const int64_t rows = batch * n;
  const int64_t grid = (rows + 7) / 8;
  if (grid > 2147483647LL) return SQB_E_UNSUPPORTED;
  project_kernel<<<(unsigned)grid, 256, 0, sqb_stream(stream)>>>(n, rows, reinterpret_cast<const c128*>(transform),
                                                                reinterpret_cast<const c128*>(psi),
                                                                reinterpret_cast<c128*>(c));
  SQB_LAUNCH_CHECK();
  return SQB_OK;
}

extern "C" int sqb_apply_transform(int n, int64_t batch, int64_t lead, int mode, const sqb_c128* transform,
                                   const sqb_c128* in, sqb_c128* out, void* stream) {
  if (n < 1 || batch < 0 || lead < 0 || (mode != 0 && mode != 1) || !transform || !in || !out || in == out)
    return SQB_E_BADARG;
  const int64_t total = lead * batch * n;
  if (total == 0) return SQB_OK;
  const int64_t grid = mode == 1 ? (total + 7) / 8 : (total + 255) / 256;
  if (grid > 2147483647LL) return SQB_E_UNSUPPORTED;
  apply_transform_kernel<<<(unsigned)grid, 256, 0, sqb_stream(stream)>>>(
      n, batch, lead, mode, reinterpret_cast<const c128*>(transform), reinterpret_cast<const c128*>(in),
      reinterpret_cast<c128*>(out));
  SQB_LAUNCH_CHECK();
  return SQB_OK;
}

extern "C" int sqb_evolve_eigen(int64_t m, const double* w, const sqb_c128* c, const double* times, int64_t n_times,
                                double hbar, sqb_c128* out, void* stream) {
  if (m < 0 || n_times < 0 || !w || !c || !times || !out || hbar == 0.0) return SQB_E_BADARG;
  if (m == 0 || n_times == 0) return SQB_OK;
  const int64_t total = m * n_times;
  const int grid = (int)sqb_min64((total + 255) / 256, 148 * 32);
  evolve_eigen_kernel<<<grid, 256, 0, sqb_stream(stream)>>>(m, w, reinterpret_cast<const c128*>(c), times, n_times,
                                                            hbar, reinterpret_cast<c128*>(out));
  SQB_LAUNCH_CHECK();
  return SQB_OK;
}

namespace {
int launch_evolve_bloch(int n, int64_t batch, const sqb_c128* transform, const double* w, const sqb_c128* c,
                        const double* times, int64_t t_count, double hbar, sqb_c128* out, int64_t out_row_stride,
                        const c128* step8, const c128* pow8, const c128* tilebase, const double* bsum, int n_pad,
                        void* stream) {
  if (n < 1 || batch < 0 || t_count < 0 || !transform || !w || !c || !times || !out || hbar == 0.0 ||
      out_row_stride < batch * n)
    return SQB_E_BADARG;
  if (batch == 0 || t_count == 0) return SQB_OK;
  const bool use_3m = step8 != nullptr;  // uniform grid: 3M complex product, 64 x 64 CTA tiles
  cudaError_t e = use_3m ? cudaFuncSetAttribute(evolve_bloch_3m_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                (int)kGemm3Smem)
                         : cudaFuncSetAttribute(evolve_bloch_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                (int)kGemmSmem);
  if (e != cudaSuccess) return (int)e;
  int64_t t_tiles = (t_count + TM - 1) / TM;
  static const int tiles_env = getenv("SQB_EVOLVE_TILES_PER_CTA") ? atoi(getenv("SQB_EVOLVE_TILES_PER_CTA")) : 0;
  const int tiles_per_cta = tiles_env > 0 ? tiles_env : 16;  // measured at cfg3: 4 -> 224.0, 8 -> 221.6, 16 -> 220.4, 32 -> 220.5 ms
  if (use_3m) t_tiles = (t_tiles + tiles_per_cta - 1) / tiles_per_cta;  // CTAs walk several time tiles
  if (t_tiles > 65535) return SQB_E_UNSUPPORTED;
  const int tn = use_3m ? TN3 : TN;
  const unsigned p_tiles = (unsigned)((n + tn - 1) / tn);
  for (int64_t b0 = 0; b0 < batch; b0 += 65535) {
    const int64_t cnt = sqb_min64(65535, batch - b0);
    dim3 grid(p_tiles, (unsigned)t_tiles, (unsigned)cnt);
    const c128* vb = reinterpret_cast<const c128*>(transform) + b0 * (int64_t)n * n;
    const c128* cbp = reinterpret_cast<const c128*>(c) + b0 * n;
    c128* ob = reinterpret_cast<c128*>(out) + b0 * n;
    if (use_3m)
      evolve_bloch_3m_kernel<<<grid, kGemmThreads, kGemm3Smem, sqb_stream(stream)>>>(
          n, vb, cbp, t_count, ob, out_row_stride, step8 + b0 * n, pow8 + b0 * n * 8, tilebase + b0 * n, batch * n,
          bsum + b0 * n * (int64_t)n_pad, n_pad, tiles_per_cta);
    else
      evolve_bloch_kernel<<<grid, kGemmThreads, kGemmSmem, sqb_stream(stream)>>>(
          n, vb, w + b0 * n, cbp, times, t_count, hbar, ob, out_row_stride, nullptr);
    SQB_LAUNCH_CHECK();
  }
  return SQB_OK;
}
}  // namespace

extern "C" int sqb_evolve_bloch(int n, int64_t batch, const sqb_c128* transform, const double* w, const sqb_c128* c,
                                const double* times, int64_t t_count, double hbar, sqb_c128* out,
                                int64_t out_row_stride, void* stream) {
  return launch_evolve_bloch(n, batch, transform, w, c, times, t_count, hbar, out, out_row_stride, nullptr, nullptr,
                             nullptr, nullptr, 0, stream);
}

namespace {
// uniform grids: 1 = NUFFT route (evolve_nufft.cuh), 0 = 3M GEMM.  SQB_EVOLVE = gemm | nufft overrides the
// heuristic (the NUFFT kernel takes <= 512 sources; a tile of 512 rows costs the same whatever t_count, so short
// row counts stay on the GEMM)
int evolve_uniform_route(int n, int64_t batch, int64_t t_count) {
  const char* env = getenv("SQB_EVOLVE");  // read per call: bench.py times both routes in one process
  if (n > kNuMaxN || n > kNuThreads || batch < 1 || t_count < 1) return 0;  // one thread per source holds (w, c)
  if (env && env[0] == 'g') return 0;
  if (env && env[0] == 'n') return 1;
  return (t_count >= 192 && n >= 32) ? 1 : 0;
}
inline size_t round_up_256(size_t x) { return (x + 255) & ~(size_t)255; }
}  // namespace

extern "C" int sqb_evolve_uniform_route(int n, int64_t batch, int64_t t_count) {
  return evolve_uniform_route(n, batch, t_count);
}

extern "C" size_t sqb_evolve_uniform_workspace_bytes(int n, int64_t batch, int64_t t_count) {
  if (n < 1 || batch < 0 || t_count < 0) return 0;
  // [flag][NUFFT plan] serve the NUFFT route; the GEMM route adds [Re+Im plane][phase tables] behind them (2 GB at
  // cfg3), so a workspace sized for a NUFFT call is small and a GEMM call on it reports SQB_E_WORKSPACE
  const size_t head = 256 + round_up_256(nufft_plan_bytes(n, batch));
  if (evolve_uniform_route(n, batch, t_count)) return head + 256;
  const int64_t tiles = (t_count + TM - 1) / TM;
  const int64_t n_pad = n + (n & 1);
  return head + round_up_256((size_t)(batch * n) * n_pad * sizeof(double)) +
         (size_t)(batch * n) * (size_t)(9 + tiles) * sizeof(c128) + 512;
}

extern "C" int sqb_evolve_bloch_uniform(int n, int64_t batch, const sqb_c128* transform, const double* w,
                                        const sqb_c128* c, const double* times, int64_t t_count, double dt, double hbar,
                                        sqb_c128* out, int64_t out_row_stride, void* workspace,
                                        size_t workspace_bytes, int reuse_transform_plane, void* stream) {
  if (n < 1 || batch < 0 || !w || !times || hbar == 0.0) return SQB_E_BADARG;
  if (batch == 0 || t_count == 0) return SQB_OK;
  if (!workspace || workspace_bytes < sqb_evolve_uniform_workspace_bytes(n, batch, t_count)) return SQB_E_WORKSPACE;
  const int64_t m = batch * n;
  const int64_t tiles = (t_count + TM - 1) / TM;
  // workspace layout: [flag: "the Re+Im plane matches the transform"][NUFFT plan][bsum: m * n_pad doubles][step8: m]
  // [pow8: 8 m][tilebase: tiles * m]; everything before the phase tables has a position that does not depend on
  // t_count, so the plane can be reused across calls on the same transform.  A workspace that is too small for the
  // GEMM part never held a valid plane: the flag is only trusted when the plane fits.
  const int n_pad = n + (n & 1);
  unsigned char* base = reinterpret_cast<unsigned char*>(workspace);
  int* plane_flag = reinterpret_cast<int*>(base);
  base += 256;
  void* plan_base = base;
  base += round_up_256(nufft_plan_bytes(n, batch));
  if (evolve_uniform_route(n, batch, t_count)) {
    if (!transform || !c || !out || out_row_stride < batch * n) return SQB_E_BADARG;
    return launch_evolve_nufft(n, batch, transform, w, c, times, t_count, dt, hbar, out, out_row_stride, plan_base,
                               plane_flag, reuse_transform_plane ? 0 : 1, sqb_stream(stream));
  }
  double* bsum = reinterpret_cast<double*>(base);  // [batch*n][n_pad] plane Re + Im of `transform`
  base += round_up_256((size_t)m * n_pad * sizeof(double));
  c128* step8 = reinterpret_cast<c128*>(base);
  c128* pow8 = step8 + m;
  c128* tilebase = pow8 + 8 * m;
  {
    const int grid2 = (int)sqb_min64((m * n_pad + 255) / 256, 148 * 32);
    bsum_kernel<<<grid2, 256, 0, sqb_stream(stream)>>>(n, n_pad, m, reinterpret_cast<const c128*>(transform), bsum,
                                                       plane_flag, reuse_transform_plane ? 0 : 1);
    SQB_LAUNCH_CHECK();
  }
  const int64_t total = m * (9 + tiles);
  const int grid = (int)sqb_min64((total + 255) / 256, 148 * 16);
  phase_tables_kernel<<<grid, 256, 0, sqb_stream(stream)>>>(m, w, times, t_count, dt, hbar, step8, pow8, tilebase,
                                                            plane_flag);
  SQB_LAUNCH_CHECK();
  return launch_evolve_bloch(n, batch, transform, w, c, times, t_count, hbar, out, out_row_stride, step8, pow8, tilebase,
                             bsum, n_pad, stream);
}
