// K2, stage 3 for 64 < n <= 512: blocked (compact WY) Householder back-transformation on the FP64 tensor cores.
// Included by eigh.cu (one translation unit).
//
// The reflectors H_i = I - tau_i v_i v_i^H of the tridiagonalisation are grouped in blocks of NB (32 for n <= 256,
// 16 for n <= 512 where a 32-reflector block would not fit in shared memory):
//   H_{i0} ... H_{i0+NB-1} = I - V T V^H      (LAPACK zlarft, forward / columnwise; T upper triangular)
// and a block is applied to 16 eigenvectors at a time as three small complex GEMMs in the row layout
// Z[e][k] (e = eigenvector, k = component):
//   W  = Z conj(V)        [16 x NB]   (K = support of the block, split over the warps)
//   W2 = W T^T            [16 x NB]
//   Z -= W2 V^T           [16 x n]
// The reflector-by-reflector kernel (backtransform_kernel) spends one shuffle reduction and two shared-memory
// sweeps per (reflector, eigenvector pair); here V is read from shared memory once per GEMM for 16 eigenvectors
// and all multiply-adds are DMMA.8x8x4 (8x fewer issue slots than DFMA for the same flops - on B200 the two
// paths have the same FP64 peak, so the gain is in issue slots and dependent chains, not in flops).
//
// Z lives in registers for the whole kernel in the MMA accumulator layout: warp w owns the 8-column tiles
// c = w, w+8, ... of all 16 eigenvectors.  The same registers are the A operand of W = Z conj(V): the k-slot q of
// the MMA is mapped to component 8c + 2q (first k-step) / 8c + 2q + 1 (second k-step), which is exactly what lane
// (row, q) holds, and the B fragments are loaded with the matching row order - no shuffles, no staging of Z.
// The numpy model of the blocking is in tests/eigh_model.py::back_transform_wy.
#pragma once

namespace {

constexpr int kWyWarps = 8;

// NB = reflectors per block, LC = 8-column tiles per warp (the kernel covers n <= 64 LC), ETT = 8-row tiles of
// eigenvectors per CTA
template <int NB, int LC, int ETT>
struct WyCfg {
  static constexpr int kEt = 8 * ETT;
  static constexpr int kMaxN = 64 * LC;
  static constexpr int kLdv = kMaxN + 1;  // c128 row stride of Vs[j][k] (== 1 mod 8: conflict-free fragment loads)
  static constexpr int kLdw = NB + 1;     // row stride of W / W2 / T (== 1 mod 8)
  static constexpr size_t kSmemV = (size_t)NB * kLdv * sizeof(c128);
  static constexpr size_t kSmemW = (size_t)kWyWarps * kEt * kLdw * sizeof(c128);  // per-warp partial W
  static constexpr size_t kSmemT = (size_t)NB * kLdw * sizeof(c128);
  static constexpr size_t kSmem = kSmemV + kSmemW + kSmemT + 64;
};

// c128 elements reserved per matrix for the T factors: ceil((n-1)/NB) padded [NB][NB+1] blocks, NB <= 32
__host__ __device__ inline size_t wy_t_elems(int n) { return (size_t)((n - 1 + 31) / 32 + 1) * 32 * 33; }

// ------------------------------------------------------------------------------------------------
// T factors: one CTA per (block, matrix).  G = V^H V (upper triangle) in 4 x 4 register tiles per warp, then the
// zlarft recurrence T[j][j] = tau_j, T[0:j, j] = -tau_j T[0:j, 0:j] G[0:j, j] by one warp.  Output [NB][NB+1] row
// major, zero outside the upper triangle / beyond the block's reflector count.
template <int NB>
__global__ void __launch_bounds__(256)
wy_tfactor_kernel(int n, const c128* __restrict__ y_all, const c128* __restrict__ tau_all, c128* __restrict__ t_all) {
  __shared__ c128 g[NB][NB + 1];
  __shared__ c128 t[NB][NB + 1];
  const int blk = blockIdx.x;
  const int64_t b = blockIdx.y;
  const int i0 = blk * NB;
  const int kb = min(NB, n - 1 - i0);
  const c128* y = y_all + b * (int64_t)n * n;
  const c128* tau = tau_all + b * n;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int idx = threadIdx.x; idx < NB * (NB + 1); idx += blockDim.x) {
    (&g[0][0])[idx] = make_double2(0.0, 0.0);
    (&t[0][0])[idx] = make_double2(0.0, 0.0);
  }
  __syncthreads();
  // G[a][c] = sum_r conj(V[r][a]) V[r][c] for a < c < kb, V[r][j] = y[i0+j][r] for r >= i0+j+1 (else 0).
  // A warp computes a 4 x 4 tile of entries at a time: 8 row loads feed 16 products per 32 rows (entry by entry
  // the kernel re-read the block NB-1 times through L1/L2)
  constexpr int kT = NB / 4;
  for (int tile = warp; tile < kT * (kT + 1) / 2; tile += 8) {
    // upper-triangular kT x kT grid of 4 x 4 tiles, row-major enumeration
    int ta = 0, rem = tile;
    while (rem >= kT - ta) {
      rem -= kT - ta;
      ++ta;
    }
    const int tc = ta + rem;
    const int a0 = 4 * ta, c0 = 4 * tc;
    if (c0 >= kb) continue;
    c128 acc[4][4];
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int v = 0; v < 4; ++v) acc[u][v] = make_double2(0.0, 0.0);
    for (int r = i0 + c0 + 1 + lane; r < n; r += 32) {
      c128 xa[4], xc[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        // rows beyond the block's reflector count and entries above a reflector's support (never written by the
        // tridiagonalisation: could be NaN bit patterns) are masked to zero
        xa[u] = (a0 + u < kb && r >= i0 + a0 + u + 1) ? y[(int64_t)(i0 + a0 + u) * n + r] : make_double2(0.0, 0.0);
        xc[u] = (c0 + u < kb && r >= i0 + c0 + u + 1) ? y[(int64_t)(i0 + c0 + u) * n + r] : make_double2(0.0, 0.0);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int v = 0; v < 4; ++v) cfmac(acc[u][v], xa[u], xc[v]);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int v = 0; v < 4; ++v) {
        c128 z = acc[u][v];
        z.x = warp_sum(z.x);
        z.y = warp_sum(z.y);
        if (lane == 0 && a0 + u < c0 + v && c0 + v < kb) g[a0 + u][c0 + v] = z;
      }
  }
  __syncthreads();
  if (warp == 0) {
    for (int j = 0; j < kb; ++j) {
      const c128 tj = tau[i0 + j];
      if (lane < j) {
        c128 acc = make_double2(0.0, 0.0);
        for (int l = lane; l < j; ++l) cfma(acc, t[lane][l], g[l][j]);
        t[lane][j] = make_double2(-(tj.x * acc.x - tj.y * acc.y), -(tj.x * acc.y + tj.y * acc.x));
      } else if (lane == j) {
        t[j][j] = tj;
      }
      __syncwarp();
    }
  }
  __syncthreads();
  // padded [NB][NB+1] rows: the back-transformation copies a block's T with ONE bulk copy into the same layout
  const int nblk_all = (n - 1 + NB - 1) / NB;
  c128* out = t_all + (b * nblk_all + blk) * (int64_t)(NB * (NB + 1));
  for (int idx = threadIdx.x; idx < NB * (NB + 1); idx += blockDim.x) out[idx] = (&t[0][0])[idx];
}

// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void wy_dmma(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(d0), "+d"(d1)
               : "d"(a), "d"(b));
}
__device__ __forceinline__ uint32_t wy_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void wy_mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(wy_smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void wy_mbar_arrive_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(wy_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void wy_mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(wy_smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void wy_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(wy_smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(wy_smem_u32(bar))
      : "memory");
}

template <int NB, int LC, int ETT>
__global__ void __launch_bounds__(32 * kWyWarps, ETT == 1 ? 2 : 1)
backtransform_wy_kernel(int n, c128* __restrict__ A_all, const double* __restrict__ qt_all,
                        const c128* __restrict__ y_all, const c128* __restrict__ t_all) {
  using Cfg = WyCfg<NB, LC, ETT>;
  constexpr int kLdv = Cfg::kLdv, kLdw = Cfg::kLdw, kWyEt = Cfg::kEt;
  constexpr int kJt = NB / 8;  // 8-column tiles of W
  extern __shared__ __align__(16) unsigned char wy_smem[];
  c128* Vs = reinterpret_cast<c128*>(wy_smem);                              // [j][kLdv]  V[k][j] at Vs[j*ld + k]
  c128* Wp = reinterpret_cast<c128*>(wy_smem + Cfg::kSmemV);                // [warp][e][kLdw] partial W
  c128* Ts = reinterpret_cast<c128*>(wy_smem + Cfg::kSmemV + Cfg::kSmemW);  // [j][kLdw]
  c128* Wf = Wp;                                                            // [e][kLdw] summed W (after a sync)
  c128* W2 = Wp + kWyEt * kLdw;                                             // [e][kLdw] W T^T
  uint64_t* bar = reinterpret_cast<uint64_t*>(wy_smem + Cfg::kSmemV + Cfg::kSmemW + Cfg::kSmemT);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int frow = lane >> 2, q = lane & 3;
  const int64_t b = blockIdx.y;
  const int e0 = blockIdx.x * kWyEt;
  const double* qt = qt_all + b * (int64_t)n * n;
  const c128* y = y_all + b * (int64_t)n * n;
  const int nblk = (n - 1 + NB - 1) / NB;
  const c128* tb = t_all + b * nblk * (int64_t)(NB * kLdw);
  const int ntiles = (n + 7) / 8;
  if (tid == 0) {
    wy_mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  // Z in the accumulator layout: zr/zi[a][lc][h] = Z[e0 + 8a + frow][8c + 2q + h], c = warp + 8 lc
  double zr[ETT][LC][2], zi[ETT][LC][2];
#pragma unroll
  for (int a = 0; a < ETT; ++a)
#pragma unroll
    for (int lc = 0; lc < LC; ++lc)
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int e = e0 + 8 * a + frow, k = 8 * (warp + 8 * lc) + 2 * q + h;
        zr[a][lc][h] = (e < n && k < n) ? qt[(int64_t)e * n + k] : 0.0;
        zi[a][lc][h] = 0.0;
      }

  for (int blk = nblk - 1; blk >= 0; --blk) {
    const int i0 = blk * NB;
    const int kb = min(NB, n - 1 - i0);
    const int c_first = (i0 + 1) / 8;  // tiles below hold no support row of this block
    // ---- stage V and T: reflector j of the block is ONE bulk (TMA) copy of its support rows k >= i0+j+1 into
    // Vs[j][k]; the few entries between the first active tile and the support (and the rows of a short last
    // block) are zeroed by the threads meanwhile
    if (warp == 0) {
      // reflector j = lane: one bulk copy each, all issued at once; lane 0 registers the byte count first and
      // also brings in the block's T.  Earlier generic-proxy stores (zero fills, W buffers) are ordered before
      // the async-proxy writes by the fence.
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      if (lane == 0) {
        uint32_t bytes = (uint32_t)(NB * kLdw * sizeof(c128));
        for (int j = 0; j < kb; ++j) bytes += (uint32_t)((n - (i0 + j + 1)) * sizeof(c128));
        wy_mbar_arrive_tx(bar, bytes);
        wy_bulk_g2s(Ts, tb + (int64_t)blk * (NB * kLdw), (uint32_t)(NB * kLdw * sizeof(c128)), bar);
      }
      __syncwarp();
      if (lane < kb) {
        const int k0 = i0 + lane + 1;
        wy_bulk_g2s(Vs + lane * kLdv + k0, y + (int64_t)(i0 + lane) * n + k0, (uint32_t)((n - k0) * sizeof(c128)), bar);
      }
    }
    constexpr int kGap = NB + 16;  // >= (i0 + NB) - 8 c_first: entries between the first active tile and a support
    for (int idx = tid; idx < NB * kGap; idx += 32 * kWyWarps) {
      const int j = idx / kGap, k = c_first * 8 + (idx - j * kGap);
      if (j < kb && k < i0 + j + 1) Vs[j * kLdv + k] = make_double2(0.0, 0.0);
    }
    if (kb < NB || n < ntiles * 8) {
      for (int idx = tid; idx < NB * ntiles * 8; idx += 32 * kWyWarps) {
        const int j = idx / (ntiles * 8), k = idx - j * (ntiles * 8);
        if (k >= c_first * 8 && (j >= kb || k >= n)) Vs[j * kLdv + k] = make_double2(0.0, 0.0);
      }
    }
    wy_mbar_wait(bar, (uint32_t)((nblk - 1 - blk) & 1));
    __syncthreads();

    // ---- W partial = Z conj(V) over the warp's own column tiles
    double wr[ETT][kJt][2], wi[ETT][kJt][2];
#pragma unroll
    for (int a = 0; a < ETT; ++a)
#pragma unroll
      for (int jt = 0; jt < kJt; ++jt) wr[a][jt][0] = wr[a][jt][1] = wi[a][jt][0] = wi[a][jt][1] = 0.0;
#pragma unroll
    for (int lc = 0; lc < LC; ++lc) {
      const int c = warp + 8 * lc;
      if (c < c_first || c >= ntiles) continue;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const c128* vcol = Vs + frow * kLdv + 8 * c + 2 * q + h;  // B[k-slot q][n = frow] of column tile jt
#pragma unroll
        for (int jt = 0; jt < kJt; ++jt) {
          const c128 v = vcol[8 * jt * kLdv];
#pragma unroll
          for (int a = 0; a < ETT; ++a) {
            // W = Z conj(V): Wre += Zre Vre + Zim Vim ; Wim += Zim Vre - Zre Vim
            wy_dmma(wr[a][jt][0], wr[a][jt][1], zr[a][lc][h], v.x);
            wy_dmma(wr[a][jt][0], wr[a][jt][1], zi[a][lc][h], v.y);
            wy_dmma(wi[a][jt][0], wi[a][jt][1], zi[a][lc][h], v.x);
            wy_dmma(wi[a][jt][0], wi[a][jt][1], zr[a][lc][h], -v.y);
          }
        }
      }
    }
    {
      c128* mine = Wp + warp * (kWyEt * kLdw);
#pragma unroll
      for (int a = 0; a < ETT; ++a)
#pragma unroll
        for (int jt = 0; jt < kJt; ++jt)
#pragma unroll
          for (int h = 0; h < 2; ++h)
            mine[(8 * a + frow) * kLdw + 8 * jt + 2 * q + h] = make_double2(wr[a][jt][h], wi[a][jt][h]);
    }
    __syncthreads();
    // ---- W = sum of the partials (kWyEt * NB entries over 256 threads), then W2 = W T^T
    constexpr int kTotal = kWyEt * NB;
    constexpr int kEnt = (kTotal + 32 * kWyWarps - 1) / (32 * kWyWarps);
    c128 s[kEnt];
#pragma unroll
    for (int u = 0; u < kEnt; ++u) {
      const int ent = tid + 256 * u, e = ent / NB, j = ent % NB;
      c128 acc = make_double2(0.0, 0.0);
      if (ent < kTotal) {
#pragma unroll
        for (int w = 0; w < kWyWarps; ++w) acc = cadd(acc, Wp[w * (kWyEt * kLdw) + e * kLdw + j]);
      }
      s[u] = acc;
    }
    __syncthreads();
#pragma unroll
    for (int u = 0; u < kEnt; ++u) {
      const int ent = tid + 256 * u;
      if (ent < kTotal) Wf[(ent / NB) * kLdw + (ent % NB)] = s[u];
    }
    __syncthreads();
    if (warp < ETT * kJt) {
      // W2 = W T^T as one 8 x 8 complex tile per warp: rows a = warp / kJt, columns jt = warp % kJt, K = NB
      const int a = warp / kJt, jt = warp % kJt;
      double ar[2] = {0.0, 0.0}, ai[2] = {0.0, 0.0};
#pragma unroll
      for (int ks = 0; ks < NB / 4; ++ks) {
        const c128 wv = Wf[(8 * a + frow) * kLdw + 4 * ks + q];   // A[row = frow][slot q] = W[e][l = 4ks + q]
        const c128 tv = Ts[(8 * jt + frow) * kLdw + 4 * ks + q];  // B[slot q][n = frow] = T[j = 8jt + frow][l]
        wy_dmma(ar[0], ar[1], wv.x, tv.x);
        wy_dmma(ar[0], ar[1], -wv.y, tv.y);
        wy_dmma(ai[0], ai[1], wv.x, tv.y);
        wy_dmma(ai[0], ai[1], wv.y, tv.x);
      }
#pragma unroll
      for (int h = 0; h < 2; ++h) W2[(8 * a + frow) * kLdw + 8 * jt + 2 * q + h] = make_double2(ar[h], ai[h]);
    }
    __syncthreads();
    // ---- Z -= W2 V^T on the warp's own column tiles.  The tile loop is the OUTER one and skips whole tiles with a
    // warp-uniform branch: with the reflector loop outside, ptxas predicated the DMMAs of the tiles above the block's
    // support instead of branching around them, and predicated-off DMMAs still go through the tensor pipe (ncu
    // counted 392 instead of 272 (tile, block) pairs of DMMAs per CTA at n = 256)
    c128 wa[NB / 4][ETT];
#pragma unroll
    for (int js = 0; js < NB / 4; ++js)
#pragma unroll
      for (int a = 0; a < ETT; ++a) wa[js][a] = W2[(8 * a + frow) * kLdw + 4 * js + q];  // A[row = frow][k-slot q]
#pragma unroll
    for (int lc = 0; lc < LC; ++lc) {
      const int c = warp + 8 * lc;
      if (c < c_first || c >= ntiles) continue;
#pragma unroll
      for (int js = 0; js < NB / 4; ++js) {
        const c128 v = Vs[(4 * js + q) * kLdv + 8 * c + frow];  // B[k-slot q][n = frow] = V[k = 8c+frow][j]
#pragma unroll
        for (int a = 0; a < ETT; ++a) {
          // Zre -= W2re Vre - W2im Vim ; Zim -= W2re Vim + W2im Vre
          wy_dmma(zr[a][lc][0], zr[a][lc][1], -wa[js][a].x, v.x);
          wy_dmma(zr[a][lc][0], zr[a][lc][1], wa[js][a].y, v.y);
          wy_dmma(zi[a][lc][0], zi[a][lc][1], -wa[js][a].x, v.y);
          wy_dmma(zi[a][lc][0], zi[a][lc][1], -wa[js][a].y, v.x);
        }
      }
    }
    __syncthreads();  // Vs / W buffers are rewritten by the next block
  }

  // rows = eigenvectors of H: the conjugate undoes the conj(H) view of the tridiagonalisation
  c128* out = A_all + b * (int64_t)n * n;
#pragma unroll
  for (int a = 0; a < ETT; ++a)
#pragma unroll
    for (int lc = 0; lc < LC; ++lc)
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int e = e0 + 8 * a + frow, k = 8 * (warp + 8 * lc) + 2 * q + h;
        if (e < n && k < n) out[(int64_t)e * n + k] = make_double2(zr[a][lc][h], -zi[a][lc][h]);
      }
}

template <int NB, int LC, int ETT>
int launch_back_wy(int n, int64_t count, c128* A, const double* qt, const c128* y, const c128* tau, c128* wy_t,
                   cudaStream_t st) {
  using Cfg = WyCfg<NB, LC, ETT>;
  cudaError_t e = cudaFuncSetAttribute(backtransform_wy_kernel<NB, LC, ETT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)Cfg::kSmem);
  if (e != cudaSuccess) return (int)e;
  wy_tfactor_kernel<NB><<<dim3((unsigned)((n - 1 + NB - 1) / NB), (unsigned)count), 256, 0, st>>>(n, y, tau, wy_t);
  SQB_LAUNCH_CHECK();
  backtransform_wy_kernel<NB, LC, ETT><<<dim3((unsigned)((n + Cfg::kEt - 1) / Cfg::kEt), (unsigned)count),
                                         32 * kWyWarps, Cfg::kSmem, st>>>(n, A, qt, y, wy_t);
  SQB_LAUNCH_CHECK();
  return SQB_OK;
}

}  // namespace
