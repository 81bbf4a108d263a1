/*
 * slate_b200.h -- C ABI of libslate_b200.so: the B200 (sm_100a) implementation of slate_quantum's
 * Bloch build -> batched Hermitian eigh -> eigenbasis evolution hot path.
 *
 * Conventions
 *  - All data pointers are DEVICE pointers owned by the caller (the Python layer owns them through
 *    torch tensors) unless the parameter name ends in `_host`.
 *  - complex128 is interleaved (re, im) doubles: `sqb_c128`.
 *  - Every call only ENQUEUES work on `stream` (a cudaStream_t passed as void*; NULL = legacy default
 *    stream) and returns immediately; nothing is allocated behind the caller's back - scratch space is
 *    passed in as `workspace` with the size reported by the matching *_workspace_bytes function.
 *  - Return value: 0 = OK, < 0 = bad argument (SQB_E_*), > 0 = a cudaError_t from a launch.
 *    Numerical failures of the eigensolver are reported per matrix in `info[]`, not in the return value.
 *  - Re-entrant / thread-safe: no global mutable state.  A few environment variables are read once per process
 *    (tuning / fallback switches, never needed for correctness): SQB_EIGH_TRIDIAG_SOLVER=ql selects the
 *    QL + rotation-replay tridiagonal stage instead of divide and conquer; SQB_FFT_INPLACE=0 uses the two-buffer FFT pass for 256/512-point strided lines; SQB_EVOLVE=gemm|nufft forces the route of
 *    sqb_evolve_bloch_uniform (default: NUFFT for n <= 512 and t_count >= 192; this one is read at EVERY call);
 *    SQB_EVOLVE_TILES_PER_CTA=<k> sets how
 *    many 64-row time tiles a CTA of the uniform-grid GEMM evolve kernel walks (default 16); SQB_BACKTRANSFORM=reflector
 *    selects the reflector-by-reflector back-transformation for every n (default: blocked compact-WY DMMA
 *    kernel for 64 < n <= 512); SQB_WY_VARIANT=<1|2> picks the blocked back-transformation's tile shape (1: 16-reflector
 *    blocks, two CTAs per SM, default; 2: 32-reflector blocks); SQB_TRIDIAG=column selects the column-by-column
 *    tridiagonalisation instead of the blocked panel kernel (default for 64 < n <= 512), SQB_TRIDIAG_NB=<4|8|16> its
 *    panel width (default 8); SQB_FFT_RADIX16=0 keeps power-of-two FFT lines on radix 8 / 4 / 2 stages (default:
 *    lines of 128 points and more lead with radix-16 stages), SQB_FFT_LINES16=<k> caps the adjacent lines a CTA of
 *    that variant takes (default 32); SQB_FOLD_FUSED=0 runs sqb_fold_transform as per-axis FFT passes plus the
 *    twiddle pass instead of the fused single pass; SQB_CELL_FFT_CLUSTER=0 runs the cell FFT of 64 x 64 block grids as
 *    one pass per axis instead of the one-pass cluster kernel (=1: its pull variant).
 *
 * Each entry point cites the reference interface (Matt-Ord/slate_quantum, paths relative to the
 * reference root) whose arithmetic it replaces.  That arithmetic lives in the un-vendored dependency
 * slate_core@4aa5106 and is reached through the cited call sites.
 */
#ifndef SLATE_B200_H
#define SLATE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct sqb_c128 {
  double re, im;
} sqb_c128;

#define SQB_MAX_DIMS 4

enum {
  SQB_OK = 0,
  SQB_E_BADARG = -1,     /* null pointer / non-positive size / nd out of range            */
  SQB_E_UNSUPPORTED = -2, /* size outside what the kernels implement (documented per call) */
  SQB_E_WORKSPACE = -3,  /* workspace too small                                           */
};

/* ABI version (bumped on any signature change) and a human-readable description of an error code. */
int sqb_abi_version(void);
const char* sqb_error_string(int code);

/* ------------------------------------------------------------------------------------------------
 * K1  Bloch block builder.
 * Replaces: slate_quantum/bloch/build.py:146-167 `kinetic_hamiltonian` (python loop over k-points
 * at :156-161 + densification `with_operator_basis(transformed...)` at :123-128), using
 * slate_quantum/operator/_build/_hamiltonian.py:32-48,72-94,121-146 for the kinetic diagonal.
 *
 *   H[b, k, k'] = E_b[k] * delta(k,k') + vtable[(k' - k) mod shape] / sqrt(N)
 *
 * nd        number of axes (1..SQB_MAX_DIMS); N = prod(shape), n_k = prod(repeat)
 * dk_host   [nd*nd] row i = reciprocal vector dk_i (Cartesian components), host memory
 * vtable    [N] device: potential Fourier components zero-padded to `shape`, FFT order per axis,
 *           C order over axes, NOT yet divided by sqrt(N)
 * blocks    [block_begin, block_begin + block_count) in meshgrid(indexing="ij") C order of the
 *           FFT-ordered fractions (bloch/build.py:40-42,138-143)
 * H_out     [block_count, N, N] row = ket k, col = bra k'
 * ---------------------------------------------------------------------------------------------- */
int sqb_bloch_build(int nd, const int* shape_host, const int* repeat_host, const double* dk_host,
                    const sqb_c128* vtable, double mass, double hbar, int64_t block_begin,
                    int64_t block_count, sqb_c128* H_out, void* stream);

/* Kinetic diagonals only: E_out[block_count, N] complex128 (imag 0), same block order.
 * Replaces operator/_build/_hamiltonian.py:149-158 `kinetic_energy` over a list of fractions. */
int sqb_bloch_kinetic(int nd, const int* shape_host, const int* repeat_host, const double* dk_host,
                      double mass, double hbar, int64_t block_begin, int64_t block_count,
                      sqb_c128* E_out, void* stream);

/* Kinetic diagonal of ONE cell for an explicit Bloch fraction (NULL = zero): E_out[N].
 * Replaces operator/_build/_hamiltonian.py:149-158 `kinetic_energy(metadata, mass, bloch_fraction)`. */
int sqb_kinetic_energy(int nd, const int* shape_host, const double* dk_host, const double* fraction_host,
                       double mass, double hbar, sqb_c128* E_out, void* stream);

/* ------------------------------------------------------------------------------------------------
 * K5  Bloch index permutation (BlochShiftedBasis per axis + BlochTransposedBasis).
 * Replaces: bloch/_shifted_basis.py:62-84 / 87-111 and bloch/_transposed_basis.py:102-133 / 136-167.
 * Vectors are [lead, M] with M = prod(shape)*prod(repeat).
 *   direction 0 (= __from_inner__):  supercell-momentum order -> Bloch [(blk...),(in...)] order
 *   direction 1 (= __into_inner__):  Bloch order -> supercell-momentum order
 * in != out.
 * ---------------------------------------------------------------------------------------------- */
int sqb_bloch_permute(int nd, const int* shape_host, const int* repeat_host, int direction,
                      int64_t lead, const sqb_c128* in, sqb_c128* out, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Time-reversal pairing of Bloch blocks (no counterpart in the reference, which diagonalises every block:
 * bloch/_linalg.py:72-84 -> operator/_linalg.py:45-83).  On an ORTHOGONAL lattice with a Hermitian-symmetric potential
 * table (a real V(x): every physical potential) the blocks of k and -k satisfy H_{-k}[pi g, pi g'] = conj(H_k[g, g'])
 * with s = q + R g -> -s (mod R N per axis) on the supercell momentum index, so only the blocks of half the k-grid
 * need the eigensolver.  sqb_bloch_time_reversal_mirror fills the eigensystems of blocks [first, first + count) of the
 * FULL [n_k, n, n] transform / [n_k, n] eigenvalue arrays from their partners (which must lie outside that range and
 * hold rows = eigenvectors): V[b'][e][pi g] = conj(V[b][e][g]), w[b'] = w[b].  The caller is responsible for the two
 * preconditions (pipeline.BlochSolver.diagonalise(time_reversal=True) checks the lattice and verifies the table).
 * ---------------------------------------------------------------------------------------------- */
int sqb_bloch_time_reversal_mirror(int nd, const int* shape_host, const int* repeat_host, int64_t first, int64_t count,
                                   sqb_c128* transform, double* w, void* stream);

/* ------------------------------------------------------------------------------------------------
 * K4  complex128 FFT, norm="ortho", over the trailing `nd` axes of a [batch, dims...] C-order array.
 * Replaces: slate_core TransformedBasis conversions (np.fft.fft / ifft(norm="ortho") per axis;
 * convention pinned by tests/test_operator.py:197-204) reached from Array.with_basis
 * (operator/_operator.py:70-74, state/_state.py:37-43).
 *   sign = -1 : forward  (np.fft.fftn,  position -> momentum on a ket index)
 *   sign = +1 : backward (np.fft.ifftn, momentum -> position on a ket index)
 * Any axis length: lengths that factor into primes <= 31 run on the shared-memory passes, the others as a chirp-z
 * (Bluestein) convolution on the power-of-two passes (workspace sized by sqb_fft_workspace_bytes).  in == out allowed.
 * ---------------------------------------------------------------------------------------------- */
size_t sqb_fft_workspace_bytes(int nd, const int* dims_host, int64_t batch);
int sqb_fft_c2c(int nd, const int* dims_host, int64_t batch, int sign, const sqb_c128* in,
                sqb_c128* out, void* workspace, size_t workspace_bytes, void* stream);
/* Same transform along ONE axis of a [outer, len, inner] array (used for operator indices). */
int sqb_fft_axis(int64_t outer, int len, int64_t inner, int sign, const sqb_c128* in, sqb_c128* out,
                 void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------
 * K4b  Bloch "cell FFT": the supercell FFT of Bloch-ordered data, factorised.
 * With s_a = k_a R_a + q_a (supercell momentum) and x_a = c_a N_a + j_a (supercell position)
 *   exp(2 pi i s_a x_a/L_a) = exp(2 pi i k_a j_a/N_a) * exp(2 pi i q_a j_a/L_a) * exp(2 pi i q_a c_a/R_a)
 * so TransformedBasis o BlochShiftedBasis o BlochTransposedBasis (the chain of bloch/build.py:79-88 that a
 * StateList.with_state_basis(position basis) walks, state/_state.py:153-165) = N-point FFT inside each
 * block, a twiddle, and an R-point FFT across blocks.
 *
 * sqb_fold_transform: folded[b,e,j] = exp(+2 pi i sum_a q_a(b) j_a / L_a) * ifftn_ortho_N(transform[b,e,:])[j]
 *   (eigenvectors moved ONCE to the in-cell position basis with the Bloch twiddle applied; feeding `folded`
 *   to sqb_evolve_bloch yields states in the "cell layout" [T, n_k, N(j)]).  Blocks
 *   [block_begin, block_begin+batch) of the k-grid `repeat`.  workspace: sqb_fft_workspace_bytes(nd, shape, batch*N).
 * sqb_cell_fft: [lead, n_k, N] cell layout <-> [lead, M] supercell position order, ortho over n_k.
 *   direction 1: cell -> position (sign +1).  For nd > 1 `in_scratch` is CLOBBERED (used in place for the
 *                first nd-1 axes); the last axis scatters into `out`.
 *   direction 0: position -> cell (sign -1); `in_scratch` is only read.
 *   in_scratch != out for nd > 1.  Every repeat[a] <= 4096 with prime factors <= 31.
 * ---------------------------------------------------------------------------------------------- */
size_t sqb_cell_fft_workspace_bytes(int nd, const int* repeat_host);
int sqb_cell_fft(int nd, const int* shape_host, const int* repeat_host, int direction, int64_t lead,
                 sqb_c128* in_scratch, sqb_c128* out, void* workspace, size_t workspace_bytes, void* stream);
/* sqb_cell_fft_moments: sqb_cell_fft (direction 1) with K6 FUSED into its last pass - every position-basis value
 * is reduced into the observable sums of sqb_measure_moments for EVERY axis instead of being stored, so the
 * position-basis state list (68.7 GB at cfg3) is never written (the consumers are
 * operator/_measure.py:191-305 on the lists of state/_state.py:153-165).  Orthogonal axes.
 *   stage 0: run the in-place passes of axes 0..nd-2 on `in_scratch` first (clobbers it);
 *   stage 1: `in_scratch` already holds them (a second reduction of the SAME tile, e.g. the variance around the
 *            periodic position obtained from the stage-0 sums).
 *   spacing_host / length_host [nd]: grid spacing and supercell length per axis;  offsets [nd][lead] (device) or
 *   NULL: per-state shifts added to x;  wrapped: fold x into [-length/2, length/2).
 *   out [lead][nd][5] = {sum |psi|^2, sum x |psi|^2, sum x^2 |psi|^2, sum cos(2 pi n/N) |psi|^2, sum sin |psi|^2}.
 * Deterministic (per-CTA partial sums, reduced in a fixed order). */
size_t sqb_cell_fft_moments_workspace_bytes(int nd, const int* shape_host, const int* repeat_host, int64_t lead);
int sqb_cell_fft_moments(int nd, const int* shape_host, const int* repeat_host, int64_t lead, int stage,
                         sqb_c128* in_scratch, const double* spacing_host, const double* length_host,
                         const double* offsets, int wrapped, double* out, void* workspace, size_t workspace_bytes,
                         void* stream);
/* sqb_cell_fft_sharded: direction 1 (cell -> position) reading the RECEIVE BUFFER of the k-shard -> time-shard
 * all-to-all directly: in_scratch is [ranks][lead][n_k/ranks][N] (rank r contributed blocks
 * [r n_k/ranks, (r+1) n_k/ranks), i.e. block axis 0 split in `ranks` contiguous pieces; repeat[0] % ranks == 0).
 * Saves the transposition copy between the collective and the FFT. */
int sqb_cell_fft_sharded(int nd, const int* shape_host, const int* repeat_host, int64_t lead, int ranks,
                         sqb_c128* in_scratch, sqb_c128* out, void* workspace, size_t workspace_bytes,
                         void* stream);
int sqb_fold_transform(int nd, const int* shape_host, const int* repeat_host, int64_t block_begin,
                       int64_t batch, const sqb_c128* transform, sqb_c128* folded, void* workspace,
                       size_t workspace_bytes, void* stream);

/* sqb_bloch_cell_vectors: the in-cell half of the same factorisation for VECTORS in Bloch order (what
 * BlochTransposedBasis.__into_fundamental__ / __from_fundamental__ walk for a state or a StateList,
 * bloch/_transposed_basis.py:102-167 + _shifted_basis.py:62-111 + the TransformedBasis leaves):
 *   direction 1: [lead, n_k, N] Bloch momentum order -> cell layout [lead, n_k, N(j)]
 *                (in-cell FFT of every block with direction `sign`, then the twiddle exp(sign 2 pi i q j / L))
 *   direction 0: the inverse.  sqb_cell_fft finishes (or starts) the job with the R-point passes.
 * sign = +1 for ket indices (momentum -> position is ifft), -1 for bra (dual) indices.  in == out allowed.
 * workspace: sqb_fft_workspace_bytes(nd, shape, lead * n_k). */
int sqb_bloch_cell_vectors(int nd, const int* shape_host, const int* repeat_host, int direction, int sign,
                           int64_t lead, const sqb_c128* in, sqb_c128* out, void* workspace,
                           size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------
 * K2  Batched Hermitian eigensolver, complex128, one problem per k-block.
 * Replaces: slate_core.linalg.into_diagonal_hermitian (np.linalg.eigh / LAPACK zheevd per block),
 * call site operator/_linalg.py:56 (`into_diagonal_hermitian`), bloch/_linalg.py:72-84.
 *
 * A_inout  [batch, n, n] in:  H_b (row-major, Hermitian; only the UPPER triangle H[r<=c] is read)
 *                        out: transform_b[e, k] = component k of eigenvector e (ROWS = eigenvectors,
 *                             i.e. np.linalg.eigh(H_b)[1].T - the layout of the reference's
 *                             EigenstateBasis transform, bloch/_linalg.py:35-57)
 * w        [batch, n] eigenvalues, ascending within each matrix (quirk Q3: not globally sorted)
 * info     [batch] 0 = converged; k > 0 = the implicit-QL iteration failed on k eigenvalues or ran
 *          out of rotation storage (results for that matrix are undefined)
 * n <= 512 (SQB_E_UNSUPPORTED above: one SM holds a matrix; the Python layer reduces larger matrices to batches of
 * 512 x 512 sub-problems of this call, slate_quantum_b200/core/jacobi.py).  Algorithm: Householder
 * tridiagonalisation (64 < n <= 512: blocked LAPACK-zlatrd panels - read-only matrix-vector products inside a panel
 * and ONE rank-2NB trailing update per panel on the FP64 tensor cores; n <= 64: column by column); divide and conquer on the
 * real tridiagonal (QL leaves of <= 32 rows, dlaed2-style deflation, secular equation per root, Gu/Eisenstat
 * vectors, merge products on the FP64 tensor cores); Householder back-transformation (blocked compact WY on the
 * FP64 tensor cores for 64 < n <= 512, reflector by reflector otherwise).  With the environment
 * variable SQB_EIGH_TRIDIAG_SOLVER=ql the tridiagonal stage is implicit-shift QL with recorded Givens
 * rotations replayed on shared-memory strips instead (info then also reports rotation-storage overflow).
 *
 * sqb_tridiag_eigh_batched is the tridiagonal stage on its own: T_b = tridiag(e_b, d_b, e_b) real symmetric,
 * d [batch, n], e [batch, n] (e[b, i] couples rows i and i+1; e[b, n-1] is ignored), w [batch, n] ascending,
 * q_rows [batch, n, n] with ROWS = eigenvectors.  batch <= 65535 per call.
 * ---------------------------------------------------------------------------------------------- */
size_t sqb_eigh_workspace_bytes(int n, int64_t batch);
int sqb_eigh_batched(int n, int64_t batch, sqb_c128* A_inout, double* w, int* info, void* workspace,
                     size_t workspace_bytes, void* stream);
size_t sqb_tridiag_eigh_workspace_bytes(int n, int64_t batch);
int sqb_tridiag_eigh_batched(int n, int64_t batch, const double* d, const double* e, double* w,
                             double* q_rows, int* info, void* workspace, size_t workspace_bytes,
                             void* stream);

/* ------------------------------------------------------------------------------------------------
 * K3  Eigenbasis projection, phase evolution, back-transformation.
 * Replaces: dynamics/schrodinger.py:31-56 `_solve_schrodinger_equation_diagonal`
 *   (:45-47 projection via State.with_basis -> ExplicitUnitaryBasis, :51-53 np.exp broadcast) and the
 *   caller-side StateList.with_state_basis back-transform (state/_state.py:153-165).
 *
 * transform [batch, n, n] rows = eigenvectors (output layout of sqb_eigh_batched)
 * sqb_project:       c[b, e]      = sum_k conj(transform[b,e,k]) * psi[b,k]
 * sqb_evolve_eigen:  a[t, b, e]   = c[b,e] * exp(-1j * w[b,e] * times[t] / hbar)     (the reference's
 *                    returned StateList raw data, [T, batch*n])
 * sqb_evolve_bloch:  out[t, b, k] = sum_e a[t,b,e] * transform[b,e,k]  with a[] generated on the fly
 *                    (never materialised): FP64 tensor-core (DMMA) complex GEMM per block.
 *                    Rows t in [t_begin, t_begin+t_count) of `times` are produced; `out` points at
 *                    row t_begin's storage, row stride = out_row_stride elements (>= batch*n).
 * ---------------------------------------------------------------------------------------------- */
/* sqb_evolve_bloch_uniform: same result as sqb_evolve_bloch for a UNIFORM grid times[i] = times[0] + i*dt (the
 * caller guarantees it, e.g. SpacedTimeMetadata).  The phases factorise as
 *   exp(-1j w t_{64 q + g + 8 r}) = tilebase[q] * pow8[g] * step8^r
 * (three small per-eigenvalue tables built by one pre-kernel in `workspace`,
 * sqb_evolve_uniform_workspace_bytes(n, batch, t_count) bytes), so the GEMM's producer warps use complex
 * multiplications only, and the complex product is evaluated with the 3M scheme (3 real tensor-core products
 * instead of 4).  Accuracy: <= ~10 extra roundings per element, normwise 1e-15.
 * The workspace also holds the plane Re+Im of `transform` (the third operand of 3M).  It only depends on
 * `transform`: pass reuse_transform_plane = 1 on later calls with the SAME transform, n, batch and workspace
 * (e.g. the next time tile) to skip recomputing it.
 *
 * Two routes, same result to ~1e-13 of sum_e |c_e V[e, j]| (sqb_evolve_uniform_route tells which one a call takes:
 * 1 = NUFFT, 0 = GEMM; SQB_EVOLVE = gemm | nufft overrides the heuristic n <= 512, t_count >= 192):
 *   GEMM  - the 3M tensor-core product above, 6 N flop per output element;
 *   NUFFT - the sum over eigenstates is a sum of N complex exponentials in the row index whose angles
 *           (w dt / hbar) mod 2 pi are shared by every column of a block: per tile of 512 rows the sources are
 *           gathered onto a 1024-point grid with a 14-point "exponential of semicircle" window, one 1024-point FFT
 *           per column in shared memory, scale by the inverse window transform (csrc/evolve_nufft.cuh;
 *           ~110 flop per output element whatever N).  A tile takes every K-th row of a super-tile of K x 512 rows
 *           (K a power of two chosen on the device from the spread of w dt / hbar), so that narrow-band spectra
 *           spread over the gather grid.
 * sqb_evolve_uniform_workspace_bytes is the size THIS call needs: the NUFFT route needs its plan only (~37 KB per
 * block at n = 256), the GEMM route adds the Re+Im plane and the phase tables (8 n^2 + 16 n (9 + t_count / 64) bytes
 * per block).  A workspace sized for the larger of the calls it will see can be shared by both routes: it records
 * whether its Re+Im plane matches the transform, so calls with different t_count (different routes) follow the same
 * reuse_transform_plane protocol; a call whose workspace is too small returns SQB_E_WORKSPACE. */
int sqb_evolve_uniform_route(int n, int64_t batch, int64_t t_count);
size_t sqb_evolve_uniform_workspace_bytes(int n, int64_t batch, int64_t t_count);
int sqb_evolve_bloch_uniform(int n, int64_t batch, const sqb_c128* transform, const double* w,
                             const sqb_c128* c, const double* times, int64_t t_count, double dt, double hbar,
                             sqb_c128* out, int64_t out_row_stride, void* workspace, size_t workspace_bytes,
                             int reuse_transform_plane, void* stream);
int sqb_project(int n, int64_t batch, const sqb_c128* transform, const sqb_c128* psi, sqb_c128* c,
                void* stream);
/* Generic explicit-basis conversion of a vector list [lead, batch*n] (ExplicitUnitaryBasis.__into_inner__ /
 * __from_inner__, reached from state/_state.py:102-107 and :153-165):
 *   mode 0: out[l,b,k] = sum_e in[l,b,e] transform[b,e,k]        (eigenbasis -> inner basis)
 *   mode 1: out[l,b,e] = sum_k conj(transform[b,e,k]) in[l,b,k]  (inner basis -> eigenbasis)      */
int sqb_apply_transform(int n, int64_t batch, int64_t lead, int mode, const sqb_c128* transform,
                        const sqb_c128* in, sqb_c128* out, void* stream);
int sqb_evolve_eigen(int64_t m, const double* w, const sqb_c128* c, const double* times,
                     int64_t n_times, double hbar, sqb_c128* out, void* stream);
int sqb_evolve_bloch(int n, int64_t batch, const sqb_c128* transform, const double* w,
                     const sqb_c128* c, const double* times, int64_t t_count, double hbar,
                     sqb_c128* out, int64_t out_row_stride, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Measured-peak micro-benchmarks (used by bench.py for the roofline denominators; not on the path).
 *   sqb_peak_dmma : `iters` rounds of independent DMMA.8x8x4 chains on every SM (variant 0/1/2 = 4/8/16
 *                   chains per warp; the loop is sensitive to chain count and code placement, so the
 *                   caller times every variant and takes the best); *flops_host = real FP64 flops executed.
 *   sqb_peak_dfma : same with plain DFMA.
 * ---------------------------------------------------------------------------------------------- */
int sqb_peak_dmma(int variant, int64_t iters, double* sink, double* flops_host, void* stream);
int sqb_peak_dfma(int64_t iters, double* sink, double* flops_host, void* stream);

/* ------------------------------------------------------------------------------------------------
 * K6  Observables of position-basis state lists ("next" row of the scope table).
 * Replaces the diagonal-operator expectations of slate_quantum.operator.measure: all_x
 * (operator/_measure.py:191-210, x operator operator/_build/_position.py:228-242), all_periodic_x (:242-268,
 * scattering operator :213-239), all_variance_x (:271-292 with variance_x :140-163) and all_coherent_width
 * (:295-305); normalisation as state.normalize_each (state/_state.py:294-308).
 *
 * states  [lead, M] complex128 in the position basis of a grid `grid_host` (C order, M = prod grid).
 * out     [lead, 5] doubles per state t:
 *           { sum |psi|^2, sum x |psi|^2, sum x^2 |psi|^2, sum cos(2 pi n/N) |psi|^2, sum sin(2 pi n/N) |psi|^2 }
 *         with n the index along `axis`, N = grid_host[axis], x = n * spacing + offsets[t] (offsets may be
 *         NULL = 0), folded into [-length/2, length/2) when `wrapped` != 0 (length = |delta_x[axis]|).
 * From these: <x> = out1/out0, periodic <x> = mod(atan2(out4, out3), 2 pi) * length / (2 pi), variance =
 * out2/out0 of a second call with offsets[t] = -periodic <x>_t and wrapped = 1.
 * Orthogonal axes only.  One CTA per state, deterministic summation order (axes of up to 4096 points use a
 * shared-memory table of positions and phases, longer ones evaluate them per element).
 * ---------------------------------------------------------------------------------------------- */
int sqb_measure_moments(int nd, const int* grid_host, int64_t lead, const sqb_c128* states, int axis,
                        double spacing, double length, const double* offsets, int wrapped, double* out,
                        void* stream);

/* sqb_measure_moments_lattice: the same five sums for a GENERAL (non-orthogonal) cell, where the Cartesian
 * component `axis` of the position mixes every lattice axis (operator/_build/_position.py:228-242 through
 * slate_core fundamental_stacked_x_points): x = sum_a frac_a delta_x[a][axis], frac_a = n_a / N_a plus the
 * fractional image of the per-state Cartesian offset along `axis`, folded into [-1/2, 1/2) when `wrapped`.
 * delta_x_host: [nd * nd] row a = lattice vector a (Cartesian).  The scattering sums use the index along `axis`. */
int sqb_measure_moments_lattice(int nd, const int* grid_host, int64_t lead, const sqb_c128* states, int axis,
                                const double* delta_x_host, const double* offsets, int wrapped, double* out,
                                void* stream);

/* ------------------------------------------------------------------------------------------------
 * K7  Operator / state algebra ("next" rows of the scope table).
 * sqb_zgemm_batched replaces the dense einsum contractions of slate_quantum.operator.matmul / commute
 * (operator/_linalg.py:108-137), matmul_list_operator / matmul_operator_list / get_commutator_operator_list
 * (:172-215), operator.apply / expectation / expectation_of_each (operator/_operator.py:165-215):
 *
 *     C[b] = alpha * op(A[b]) * op(B[b]) + beta * C[b]         b in [0, batch)
 *
 * A(i, k) = a[b*a_batch_stride + i*a_row_stride + k*a_col_stride] (complex conjugated when conj_a != 0), i < m,
 * k < k; B(k, j) likewise, j < n; strides in ELEMENTS, any sign-free combination (row-major, transposed,
 * batch stride 0 = the same operator for every list entry), so a transpose or dagger is a stride swap and never
 * a copy.  C is row-major with leading dimension ldc >= n; it may not overlap A or B.  alpha / beta point to
 * {re, im} on the HOST.  FP64 tensor cores (DMMA.8x8x4), 4 real products per complex MAC; 8 m n k real flop per
 * batch entry.
 *
 * sqb_rowwise_vdot    out[r] = sum_i conj(x[r, i]) y[r, i]      (state.inner_product_each, state/_state.py:208-228;
 *                     the last step of expectation_of_each)
 * sqb_rowwise_weighted_abs2  out[r] = sum_i weights[i] |x[r, i]|^2: expectation_of_each of an operator that is
 *                     DIAGONAL in the basis the states are written in (potentials, x, kinetic energy, k): one pass over
 *                     the state list instead of a dense [M, M] operator
 * sqb_scale_columns   out[r, i] = weights[i] x[r, i]: apply_to_each of a diagonal operator (operator/_operator.py:219-226)
 * sqb_abs2            out[i] = |x[i]|^2                          (state.get_all_occupations, state/_state.py:265-308)
 * sqb_normalize_rows  out[r, :] = x[r, :] / sqrt(|norm2[r]|)     (state.normalize_each, state/_state.py:242-262)
 * ---------------------------------------------------------------------------------------------- */
int sqb_zgemm_batched(int m, int n, int k, int64_t batch, const sqb_c128* a, int64_t a_row_stride,
                      int64_t a_col_stride, int64_t a_batch_stride, int conj_a, const sqb_c128* b,
                      int64_t b_row_stride, int64_t b_col_stride, int64_t b_batch_stride, int conj_b,
                      const double* alpha, const double* beta, sqb_c128* c, int64_t ldc, int64_t c_batch_stride,
                      void* stream);
int sqb_rowwise_vdot(int64_t rows, int64_t len, const sqb_c128* x, int64_t x_row_stride, const sqb_c128* y,
                     int64_t y_row_stride, sqb_c128* out, void* stream);
int sqb_rowwise_weighted_abs2(int64_t rows, int64_t len, const sqb_c128* x, int64_t x_row_stride,
                              const sqb_c128* weights, sqb_c128* out, void* stream);
int sqb_scale_columns(int64_t rows, int64_t len, const sqb_c128* x, int64_t x_row_stride, const sqb_c128* weights,
                      sqb_c128* out, void* stream);
int sqb_abs2(int64_t count, const sqb_c128* x, double* out, void* stream);
int sqb_normalize_rows(int64_t rows, int64_t len, const sqb_c128* x, const sqb_c128* norm2, sqb_c128* out,
                       void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SLATE_B200_H */
