"""Functional wrappers over the C ABI taking torch CUDA tensors (device-buffer ownership only).

Every function enqueues on torch's current CUDA stream and returns device tensors; nothing here
computes on the CPU.  Calling without a CUDA device raises.
"""

from __future__ import annotations

import os

import ctypes
from typing import Sequence

import numpy as np
import torch

from slate_quantum_b200 import _lib


LAUNCHES = 0  # kernels of libslate_b200.so launched through this module (bench.py reports it)


def _count(n: int) -> None:
    global LAUNCHES  # noqa: PLW0603
    LAUNCHES += n


def _fft_launches(dims: Sequence[int]) -> int:
    # per axis: twiddle-table kernel + pass kernel; lines longer than 4096 take two passes (four-step)
    return sum(2 if d <= 4096 else 4 for d in dims)


def _require_cuda() -> None:
    if not torch.cuda.is_available():
        msg = "slate_quantum_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback."
        raise _lib.SlateB200Error(msg)


def _stream() -> ctypes.c_void_p:
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t: torch.Tensor) -> ctypes.c_void_p:
    return ctypes.c_void_p(t.data_ptr())


def _c128(t: torch.Tensor, name: str) -> torch.Tensor:
    if not (t.is_cuda and t.dtype == torch.complex128 and t.is_contiguous()):
        msg = f"{name} must be a contiguous CUDA complex128 tensor"
        raise TypeError(msg)
    if t.is_conj() or t.is_neg():
        # a lazy .conj() / negation is a view bit the raw pointer cannot carry: the kernel would read unconjugated data
        msg = f"{name} carries a lazy conj/neg bit; materialise it first (torch.conj_physical / .resolve_conj())"
        raise TypeError(msg)
    return t


def _f64(t: torch.Tensor, name: str) -> torch.Tensor:
    if not (t.is_cuda and t.dtype == torch.float64 and t.is_contiguous()):
        msg = f"{name} must be a contiguous CUDA float64 tensor"
        raise TypeError(msg)
    if t.is_neg():
        msg = f"{name} carries a lazy neg bit; materialise it first (.resolve_neg())"
        raise TypeError(msg)
    return t


def free_bytes() -> int:
    """HBM this process can still allocate: free device memory plus what torch's caching allocator holds unused."""
    _require_cuda()
    free, _ = torch.cuda.mem_get_info()
    return int(free + torch.cuda.memory_reserved() - torch.cuda.memory_allocated())


def to_device(a: np.ndarray | torch.Tensor, dtype: torch.dtype) -> torch.Tensor:
    _require_cuda()
    if isinstance(a, torch.Tensor):
        return a.to(device="cuda", dtype=dtype).contiguous()
    return torch.from_numpy(np.ascontiguousarray(a)).to(device="cuda", dtype=dtype)


def bloch_build(
    shape: Sequence[int],
    repeat: Sequence[int],
    dk: np.ndarray,
    vtable: torch.Tensor,
    mass: float,
    hbar: float,
    block_begin: int = 0,
    block_count: int | None = None,
    out: torch.Tensor | None = None,
) -> torch.Tensor:
    """K1: H[b,k,k'] for blocks [block_begin, block_begin+block_count). See include/slate_b200.h."""
    _require_cuda()
    lib = _lib.load()
    nd = len(shape)
    n = int(np.prod(shape))
    n_k = int(np.prod(repeat))
    block_count = n_k - block_begin if block_count is None else block_count
    _c128(vtable, "vtable")
    if vtable.numel() != n:
        msg = "vtable must hold prod(shape) entries"
        raise ValueError(msg)
    if out is None:
        out = torch.empty((block_count, n, n), dtype=torch.complex128, device="cuda")
    _c128(out, "out")
    rc = lib.sqb_bloch_build(
        nd, _lib.int_array(shape), _lib.int_array(repeat), _lib.double_array(np.asarray(dk).ravel()),
        _ptr(vtable), float(mass), float(hbar), block_begin, block_count, _ptr(out), _stream(),
    )
    _lib.check(rc, "sqb_bloch_build")
    _count(1)
    return out


def bloch_kinetic(
    shape: Sequence[int], repeat: Sequence[int], dk: np.ndarray, mass: float, hbar: float,
    block_begin: int = 0, block_count: int | None = None,
) -> torch.Tensor:
    _require_cuda()
    lib = _lib.load()
    n = int(np.prod(shape))
    n_k = int(np.prod(repeat))
    block_count = n_k - block_begin if block_count is None else block_count
    out = torch.empty((block_count, n), dtype=torch.complex128, device="cuda")
    rc = lib.sqb_bloch_kinetic(
        len(shape), _lib.int_array(shape), _lib.int_array(repeat), _lib.double_array(np.asarray(dk).ravel()),
        float(mass), float(hbar), block_begin, block_count, _ptr(out), _stream(),
    )
    _lib.check(rc, "sqb_bloch_kinetic")
    _count(1)
    return out


def kinetic_energy(
    shape: Sequence[int], dk: np.ndarray, fraction: np.ndarray | None, mass: float, hbar: float
) -> torch.Tensor:
    """Kinetic diagonal E[N] of one cell for an explicit Bloch fraction (None = 0)."""
    _require_cuda()
    lib = _lib.load()
    out = torch.empty((int(np.prod(shape)),), dtype=torch.complex128, device="cuda")
    frac = None if fraction is None else _lib.double_array(np.asarray(fraction, dtype=np.float64).ravel())
    rc = lib.sqb_kinetic_energy(
        len(shape), _lib.int_array(shape), _lib.double_array(np.asarray(dk).ravel()), frac, float(mass), float(hbar),
        _ptr(out), _stream(),
    )
    _lib.check(rc, "sqb_kinetic_energy")
    _count(1)
    return out


def bloch_permute(
    shape: Sequence[int], repeat: Sequence[int], vectors: torch.Tensor, direction: int, out: torch.Tensor | None = None
) -> torch.Tensor:
    """K5 on the last axis of `vectors` ([..., M]).  direction 0: supercell -> Bloch; 1: Bloch -> supercell."""
    _require_cuda()
    lib = _lib.load()
    _c128(vectors, "vectors")
    m = int(np.prod(shape)) * int(np.prod(repeat))
    if vectors.shape[-1] != m:
        msg = f"last axis must have {m} entries"
        raise ValueError(msg)
    out = torch.empty_like(vectors) if out is None else _c128(out, "out")
    lead = vectors.numel() // m
    rc = lib.sqb_bloch_permute(
        len(shape), _lib.int_array(shape), _lib.int_array(repeat), direction, lead, _ptr(vectors), _ptr(out), _stream()
    )
    _lib.check(rc, "sqb_bloch_permute")
    _count(1)
    return out


def fft_c2c(x: torch.Tensor, dims: Sequence[int], sign: int, out: torch.Tensor | None = None) -> torch.Tensor:
    """K4 over the trailing len(dims) axes of a [batch, prod(dims)]-like tensor, norm="ortho"."""
    _require_cuda()
    lib = _lib.load()
    _c128(x, "x")
    total = int(np.prod(dims))
    if x.numel() % total:
        msg = "tensor size is not a multiple of prod(dims)"
        raise ValueError(msg)
    batch = x.numel() // total
    out = torch.empty_like(x) if out is None else _c128(out, "out")
    dims_c = _lib.int_array(dims)
    ws_bytes = lib.sqb_fft_workspace_bytes(len(dims), dims_c, batch)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device="cuda")
    rc = lib.sqb_fft_c2c(len(dims), dims_c, batch, sign, _ptr(x), _ptr(out), _ptr(ws), ws_bytes, _stream())
    _lib.check(rc, "sqb_fft_c2c")
    _count(_fft_launches(dims))
    return out


def fft_axis(x: torch.Tensor, outer: int, length: int, inner: int, sign: int) -> torch.Tensor:
    _require_cuda()
    lib = _lib.load()
    _c128(x, "x")
    out = torch.empty_like(x)
    dims_c = _lib.int_array([length])
    ws_bytes = lib.sqb_fft_workspace_bytes(1, dims_c, outer * inner)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device="cuda")
    rc = lib.sqb_fft_axis(outer, length, inner, sign, _ptr(x), _ptr(out), _ptr(ws), ws_bytes, _stream())
    _lib.check(rc, "sqb_fft_axis")
    _count(_fft_launches([length]))
    return out


def cell_fft(
    shape: Sequence[int], repeat: Sequence[int], x: torch.Tensor, direction: int, out: torch.Tensor | None = None,
    ranks: int = 1,
) -> torch.Tensor:
    """K4b: [lead, n_k, N] cell layout <-> [lead, M] position order. direction 1 CLOBBERS `x` when nd > 1.

    ranks > 1 (direction 1 only): `x` is the all-to-all receive buffer [ranks, lead, n_k/ranks, N]."""
    _require_cuda()
    lib = _lib.load()
    _c128(x, "x")
    m = int(np.prod(shape)) * int(np.prod(repeat))
    lead = x.numel() // m
    if out is None:
        out = torch.empty((lead, m), dtype=torch.complex128, device="cuda")
    _c128(out, "out")
    rep_c = _lib.int_array(repeat)
    ws_bytes = lib.sqb_cell_fft_workspace_bytes(len(repeat), rep_c)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device="cuda")
    if ranks > 1:
        rc = lib.sqb_cell_fft_sharded(
            len(shape), _lib.int_array(shape), rep_c, lead, ranks, _ptr(x), _ptr(out), _ptr(ws), ws_bytes, _stream()
        )
    else:
        rc = lib.sqb_cell_fft(
            len(shape), _lib.int_array(shape), rep_c, direction, lead, _ptr(x), _ptr(out), _ptr(ws), ws_bytes,
            _stream(),
        )
    _lib.check(rc, "sqb_cell_fft")
    one_pass = (  # the cluster kernel of csrc/fft.cu (cell_fft_impl): table + one launch instead of a pass per axis
        direction == 1 and ranks == 1 and len(shape) == 2 and tuple(repeat) == (64, 64) and shape[1] % 8 == 0
        and x.data_ptr() != out.data_ptr() and os.environ.get("SQB_CELL_FFT_CLUSTER", "2") != "0"
    )
    _count(2 if one_pass else 2 * len(shape))
    return out


def cell_fft_moments(
    shape: Sequence[int], repeat: Sequence[int], x: torch.Tensor, spacing: Sequence[float], length: Sequence[float],
    stage: int = 0, offsets: torch.Tensor | None = None, wrapped: bool = False,
) -> torch.Tensor:
    """K4b + K6 fused: observable sums [lead, nd, 5] of the position-basis states of the cell-layout tile `x`
    ([lead, n_k, N]) for every axis, WITHOUT writing those states.  stage 0 clobbers `x` (in-place passes of the
    leading axes); stage 1 reduces the same tile again (`offsets` [nd, lead], `wrapped`)."""
    _require_cuda()
    lib = _lib.load()
    _c128(x, "x")
    nd = len(shape)
    m = int(np.prod(shape)) * int(np.prod(repeat))
    lead = x.numel() // m
    if offsets is not None:
        _f64(offsets, "offsets")
        if tuple(offsets.shape) != (nd, lead):
            msg = "offsets must be [nd, lead]"
            raise ValueError(msg)
    out = torch.empty((lead, nd, 5), dtype=torch.float64, device="cuda")
    shape_c, rep_c = _lib.int_array(shape), _lib.int_array(repeat)
    ws_bytes = lib.sqb_cell_fft_moments_workspace_bytes(nd, shape_c, rep_c, lead)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device="cuda")
    rc = lib.sqb_cell_fft_moments(
        nd, shape_c, rep_c, lead, int(stage), _ptr(x), _lib.double_array(np.asarray(spacing, dtype=np.float64)),
        _lib.double_array(np.asarray(length, dtype=np.float64)), _ptr(offsets) if offsets is not None else None,
        int(bool(wrapped)), _ptr(out), _ptr(ws), ws_bytes, _stream(),
    )
    _lib.check(rc, "sqb_cell_fft_moments")
    _count((2 * (nd - 1) if stage == 0 else 0) + nd + 3)
    return out


def fold_transform(
    shape: Sequence[int], repeat: Sequence[int], transform: torch.Tensor, block_begin: int = 0,
    out: torch.Tensor | None = None,
) -> torch.Tensor:
    """Eigenvectors -> in-cell position basis with the Bloch twiddle folded in (see include/slate_b200.h)."""
    _require_cuda()
    lib = _lib.load()
    _c128(transform, "transform")
    batch, n, _ = transform.shape
    out = torch.empty_like(transform) if out is None else _c128(out, "out")
    shape_c = _lib.int_array(shape)
    ws_bytes = lib.sqb_fft_workspace_bytes(len(shape), shape_c, batch * n)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device="cuda")
    rc = lib.sqb_fold_transform(
        len(shape), shape_c, _lib.int_array(repeat), block_begin, batch, _ptr(transform), _ptr(out), _ptr(ws),
        ws_bytes, _stream(),
    )
    _lib.check(rc, "sqb_fold_transform")
    _count(_fft_launches(shape) + 1)
    return out


def is_smooth(length: int, max_prime: int = 31) -> bool:
    """True when `length` factors into primes <= max_prime (what the shared-memory FFT passes implement)."""
    rem = int(length)
    if rem < 1:
        return False
    for p in range(2, max_prime + 1):
        while rem % p == 0:
            rem //= p
    return rem == 1


def bloch_cell_vectors(
    shape: Sequence[int], repeat: Sequence[int], x: torch.Tensor, direction: int, out: torch.Tensor | None = None,
    sign: int = 1,
) -> torch.Tensor:
    """In-cell half of the Bloch cell FFT for vectors: [lead, n_k, N] Bloch momentum order <-> cell layout.

    direction 1: in-cell FFT of every block + Bloch twiddle (cell_fft(..., 1) then gives the position basis);
    direction 0: the inverse (after cell_fft(..., 0)).  `out` may be `x`."""
    _require_cuda()
    lib = _lib.load()
    _c128(x, "x")
    n = int(np.prod(shape))
    n_k = int(np.prod(repeat))
    if x.numel() % (n * n_k):
        msg = "vectors must hold whole supercell states"
        raise ValueError(msg)
    lead = x.numel() // (n * n_k)
    out = torch.empty_like(x) if out is None else _c128(out, "out")
    shape_c = _lib.int_array(shape)
    ws_bytes = lib.sqb_fft_workspace_bytes(len(shape), shape_c, lead * n_k)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device="cuda")
    rc = lib.sqb_bloch_cell_vectors(
        len(shape), shape_c, _lib.int_array(repeat), int(direction), int(sign), lead, _ptr(x), _ptr(out), _ptr(ws),
        ws_bytes, _stream(),
    )
    _lib.check(rc, "sqb_bloch_cell_vectors")
    _count(_fft_launches(shape) + -(-n_k // 65535))
    return out


EIGH_KERNEL_MAX_N = 512  # sqb_eigh_batched keeps a matrix on one SM (include/slate_b200.h); larger: core/jacobi.py


def eigh_batched(a: torch.Tensor, workspace: torch.Tensor | None = None) -> tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """K2 IN PLACE: `a` [batch,n,n] is overwritten by the transform (rows = eigenvectors).

    Returns (w [batch,n] float64 ascending per block, a, info [batch] int32).
    """
    _require_cuda()
    lib = _lib.load()
    _c128(a, "a")
    batch, n, n2 = a.shape
    if n != n2:
        msg = "matrices must be square"
        raise ValueError(msg)
    if n > EIGH_KERNEL_MAX_N:
        # one SM cannot hold such a matrix: two-sided block Jacobi over batched K2 sub-problems + K7 GEMMs
        from slate_quantum_b200.core.jacobi import eigh_block_jacobi

        w = torch.empty((batch, n), dtype=torch.float64, device="cuda")
        for i in range(batch):
            wi, vi, _ = eigh_block_jacobi(a[i])
            w[i] = wi
            a[i] = vi
        return w, a, torch.zeros((batch,), dtype=torch.int32, device="cuda")
    w = torch.empty((batch, n), dtype=torch.float64, device="cuda")
    info = torch.empty((batch,), dtype=torch.int32, device="cuda")
    if workspace is None:
        ws_bytes = lib.sqb_eigh_workspace_bytes(n, batch)
        free = free_bytes()
        ws_bytes = min(ws_bytes, max(int(free * 0.6), lib.sqb_eigh_workspace_bytes(n, 1)))
        workspace = torch.empty(ws_bytes, dtype=torch.uint8, device="cuda")
    rc = lib.sqb_eigh_batched(n, batch, _ptr(a), _ptr(w), _ptr(info), _ptr(workspace), workspace.numel(), _stream())
    _lib.check(rc, "sqb_eigh_batched")
    per = lib.sqb_eigh_workspace_bytes(n, 1) - 4096
    chunk = max(1, min(batch, (workspace.numel() - 4096) // per, 65535))
    if batch >= 512 and chunk >= 256:  # two chunks in flight on two internal streams (csrc/eigh.cu)
        chunk = min(chunk // 2, (batch + 1) // 2)
    _count(_eigh_launches_per_chunk(n) * -(-batch // chunk))
    return w, a, info


def time_reversal_mirror(
    shape: Sequence[int], repeat: Sequence[int], first: int, count: int, transform: torch.Tensor, w: torch.Tensor
) -> None:
    """Fill the eigensystems of blocks [first, first + count) of the FULL [n_k, n, n] / [n_k, n] arrays from their
    time-reversal partners: V[b'][e][pi g] = conj(V[b][e][g]), w[b'] = w[b] (orthogonal lattice, real V(x))."""
    _require_cuda()
    _c128(transform, "transform")
    _f64(w, "w")
    nd = len(shape)
    sh = (ctypes.c_int * nd)(*[int(x) for x in shape])
    rp = (ctypes.c_int * nd)(*[int(x) for x in repeat])
    rc = _lib.load().sqb_bloch_time_reversal_mirror(nd, sh, rp, int(first), int(count), _ptr(transform), _ptr(w), _stream())
    _lib.check(rc, "sqb_bloch_time_reversal_mirror")
    _count(-(-int(count) // 65535))


def _dc_depth(n: int) -> int:
    depth = 0
    while ((n + (1 << depth) - 1) >> depth) > 32:
        depth += 1
    return depth


def _eigh_launches_per_chunk(n: int) -> int:
    """Kernels per chunk of matrices: tridiagonalisation + back-transformation + the tridiagonal solver."""
    import os

    if os.environ.get("SQB_EIGH_TRIDIAG_SOLVER") == "ql":
        return 4  # ql + replay
    wy = 1 if (64 < n <= 512 and os.environ.get("SQB_BACKTRANSFORM") != "reflector") else 0  # T-factor kernel
    return 2 + wy + 1 + 2 * _dc_depth(n)  # leaves + (set-up, GEMM) per merge level


def tridiag_eigh_batched(d: torch.Tensor, e: torch.Tensor) -> tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """Divide-and-conquer stage of K2 on its own: real symmetric tridiagonals (d [batch,n], e [batch,n]).

    Returns (w [batch,n] ascending, q [batch,n,n] with rows = eigenvectors, info [batch]).
    """
    _require_cuda()
    lib = _lib.load()
    if d.dtype != torch.float64 or e.dtype != torch.float64 or d.shape != e.shape or d.dim() != 2:
        msg = "d and e must be float64 [batch, n]"
        raise ValueError(msg)
    if not (d.is_cuda and e.is_cuda and d.is_contiguous() and e.is_contiguous()):
        msg = "d and e must be contiguous CUDA tensors"
        raise ValueError(msg)
    batch, n = d.shape
    w = torch.empty((batch, n), dtype=torch.float64, device="cuda")
    q = torch.empty((batch, n, n), dtype=torch.float64, device="cuda")
    info = torch.empty((batch,), dtype=torch.int32, device="cuda")
    workspace = torch.empty(lib.sqb_tridiag_eigh_workspace_bytes(n, batch), dtype=torch.uint8, device="cuda")
    rc = lib.sqb_tridiag_eigh_batched(
        n, batch, _ptr(d), _ptr(e), _ptr(w), _ptr(q), _ptr(info), _ptr(workspace), workspace.numel(), _stream()
    )
    _lib.check(rc, "sqb_tridiag_eigh_batched")
    _count(1 + 2 * _dc_depth(n))
    return w, q, info


def project(transform: torch.Tensor, psi: torch.Tensor) -> torch.Tensor:
    _require_cuda()
    lib = _lib.load()
    _c128(transform, "transform")
    _c128(psi, "psi")
    batch, n, _ = transform.shape
    c = torch.empty((batch, n), dtype=torch.complex128, device="cuda")
    rc = lib.sqb_project(n, batch, _ptr(transform), _ptr(psi), _ptr(c), _stream())
    _lib.check(rc, "sqb_project")
    _count(1)
    return c


def apply_transform(transform: torch.Tensor, vectors: torch.Tensor, mode: int) -> torch.Tensor:
    """Generic explicit-basis conversion of [lead, batch*n] lists (mode 0: eigen -> inner, 1: inner -> eigen)."""
    _require_cuda()
    lib = _lib.load()
    _c128(transform, "transform")
    _c128(vectors, "vectors")
    batch, n, _ = transform.shape
    lead = vectors.numel() // (batch * n)
    out = torch.empty_like(vectors)
    rc = lib.sqb_apply_transform(n, batch, lead, mode, _ptr(transform), _ptr(vectors), _ptr(out), _stream())
    _lib.check(rc, "sqb_apply_transform")
    _count(1)
    return out


def evolve_eigen(w: torch.Tensor, c: torch.Tensor, times: torch.Tensor, hbar: float) -> torch.Tensor:
    _require_cuda()
    lib = _lib.load()
    _f64(w, "w")
    _c128(c, "c")
    _f64(times, "times")
    m = w.numel()
    out = torch.empty((times.numel(), m), dtype=torch.complex128, device="cuda")
    rc = lib.sqb_evolve_eigen(m, _ptr(w), _ptr(c), _ptr(times), times.numel(), float(hbar), _ptr(out), _stream())
    _lib.check(rc, "sqb_evolve_eigen")
    _count(1)
    return out


def uniform_step(times: np.ndarray | Sequence[float]) -> float | None:
    """dt if `times` (host values) is a uniform grid to within a few ulp, else None."""
    t = np.asarray(times, dtype=np.float64)
    if t.size < 2:
        return None
    dt = (t[-1] - t[0]) / (t.size - 1)
    err = np.abs(t - (t[0] + dt * np.arange(t.size))).max()
    return float(dt) if err <= 4 * np.finfo(np.float64).eps * max(np.abs(t).max(), abs(dt)) else None


def evolve_uniform_workspace_bytes(n: int, batch: int, t_count: int) -> int:
    return int(_lib.load().sqb_evolve_uniform_workspace_bytes(n, batch, t_count))


def evolve_uniform_route(n: int, batch: int, t_count: int) -> str:
    """Which kernel `evolve_bloch(..., uniform_dt=dt)` runs for this shape: "nufft" or "gemm" (SQB_EVOLVE overrides)."""
    return "nufft" if _lib.load().sqb_evolve_uniform_route(n, batch, t_count) else "gemm"


def evolve_bloch(
    transform: torch.Tensor, w: torch.Tensor, c: torch.Tensor, times: torch.Tensor, hbar: float,
    out: torch.Tensor | None = None, uniform_dt: float | None = None,
    workspace: torch.Tensor | None = None, reuse_plane: bool = False,
) -> torch.Tensor:
    """K3c: out[t, b, k] for every time in `times` (fused phase x eigenvector DMMA GEMM).

    uniform_dt: pass the grid spacing when `times` is uniform (see `uniform_step`): blocks of <= 512 states and >= 192
    rows then take the NUFFT route (gather + 1024-point FFT per 512-row tile, csrc/evolve_nufft.cuh), the rest the 3M
    tensor-core GEMM with table-driven phases (`evolve_uniform_route`).  workspace / reuse_plane: a caller-owned workspace (>= evolve_uniform_workspace_bytes)
    lets consecutive calls on the SAME transform (time tiles) skip recomputing the Re+Im plane."""
    _require_cuda()
    lib = _lib.load()
    _c128(transform, "transform")
    _f64(w, "w")
    _c128(c, "c")
    _f64(times, "times")
    batch, n, _ = transform.shape
    t_count = times.numel()
    if out is None:
        out = torch.empty((t_count, batch * n), dtype=torch.complex128, device="cuda")
    # `out` may be a column slice of a wider [T, width] buffer (row stride > batch*n): the blocks of one rank
    # inside the states of all ranks
    if not (out.is_cuda and out.dtype == torch.complex128):
        msg = "out must be a CUDA complex128 tensor"
        raise TypeError(msg)
    if out.is_contiguous():
        row_stride = batch * n
    elif out.dim() == 2 and out.stride(1) == 1 and out.shape[1] == batch * n and out.stride(0) >= batch * n:
        row_stride = out.stride(0)
    else:
        msg = "out must be contiguous or a column slice [T, batch*n] of a wider row-major buffer"
        raise TypeError(msg)
    if uniform_dt is None:
        rc = lib.sqb_evolve_bloch(
            n, batch, _ptr(transform), _ptr(w), _ptr(c), _ptr(times), t_count, float(hbar), _ptr(out), row_stride,
            _stream(),
        )
        _lib.check(rc, "sqb_evolve_bloch")
        _count(-(-batch // 65535))
    else:
        ws_bytes = lib.sqb_evolve_uniform_workspace_bytes(n, batch, t_count)
        if workspace is None or workspace.numel() < ws_bytes:
            workspace, reuse_plane = torch.empty(ws_bytes, dtype=torch.uint8, device="cuda"), False
        rc = lib.sqb_evolve_bloch_uniform(
            n, batch, _ptr(transform), _ptr(w), _ptr(c), _ptr(times), t_count, float(uniform_dt), float(hbar),
            _ptr(out), row_stride, _ptr(workspace), workspace.numel(), int(bool(reuse_plane)), _stream(),
        )
        _lib.check(rc, "sqb_evolve_bloch_uniform")
        # GEMM route: Re+Im plane (returns at once when reused) + phase tables + GEMM; NUFFT route: window transform +
        # band + plan + one transform kernel per 65535 blocks
        _count((3 if lib.sqb_evolve_uniform_route(n, batch, t_count) else 2) - (-batch // 65535))
    return out


def measure_moments(
    states: torch.Tensor,
    grid: Sequence[int],
    axis: int,
    spacing: float,
    length: float,
    offsets: torch.Tensor | None = None,
    wrapped: bool = False,
) -> torch.Tensor:
    """K6: per-state sums {|psi|^2, x|psi|^2, x^2|psi|^2, cos(2 pi n/N)|psi|^2, sin(2 pi n/N)|psi|^2} -> [lead, 5].

    `states` [lead, prod(grid)] complex128 in the position basis; x = n_axis * spacing + offsets[t], folded into
    [-length/2, length/2) when `wrapped` (reference operator/_measure.py:191-305).
    """
    _require_cuda()
    lib = _lib.load()
    _c128(states, "states")
    m = int(np.prod(grid))
    if states.numel() % m:
        msg = "states must hold whole vectors over the grid"
        raise ValueError(msg)
    lead = states.numel() // m
    if offsets is not None:
        _f64(offsets, "offsets")
        if offsets.numel() != lead:
            msg = "one offset per state"
            raise ValueError(msg)
    out = torch.empty((lead, 5), dtype=torch.float64, device="cuda")
    g = (ctypes.c_int * len(grid))(*[int(x) for x in grid])
    rc = lib.sqb_measure_moments(
        len(grid), g, lead, _ptr(states), int(axis), float(spacing), float(length),
        _ptr(offsets) if offsets is not None else None, int(bool(wrapped)), _ptr(out), _stream(),
    )
    _lib.check(rc, "sqb_measure_moments")
    _count(1)
    return out


def measure_moments_lattice(
    states: torch.Tensor, grid: Sequence[int], axis: int, delta_x: np.ndarray, offsets: torch.Tensor | None = None,
    wrapped: bool = False,
) -> torch.Tensor:
    """K6 for a general (non-orthogonal) cell: same five sums as `measure_moments`, x = Cartesian component `axis`
    of the grid position for lattice vectors `delta_x` ([nd, nd], row a = vector a)."""
    _require_cuda()
    lib = _lib.load()
    _c128(states, "states")
    m = int(np.prod(grid))
    if states.numel() % m:
        msg = "states must hold whole vectors over the grid"
        raise ValueError(msg)
    lead = states.numel() // m
    if offsets is not None:
        _f64(offsets, "offsets")
        if offsets.numel() != lead:
            msg = "one offset per state"
            raise ValueError(msg)
    nd = len(grid)
    out = torch.empty((lead, 5), dtype=torch.float64, device="cuda")
    g = (ctypes.c_int * nd)(*[int(x) for x in grid])
    rc = lib.sqb_measure_moments_lattice(
        nd, g, lead, _ptr(states), int(axis), _lib.double_array(np.asarray(delta_x, dtype=np.float64)[:nd, :nd].ravel()),
        _ptr(offsets) if offsets is not None else None, int(bool(wrapped)), _ptr(out), _stream(),
    )
    _lib.check(rc, "sqb_measure_moments_lattice")
    _count(1)
    return out


def _gemm_operand(t: torch.Tensor, name: str) -> tuple[torch.Tensor, int, int, int, int]:
    """(tensor, row stride, col stride, batch stride, conj) of a [m, k] or [batch, m, k] complex128 CUDA view.
    Transposed (`.mT`), daggered (`.mH`), conjugated and broadcast (`expand`) views are passed as they are."""
    if not (t.is_cuda and t.dtype == torch.complex128 and t.dim() in (2, 3)):
        msg = f"{name} must be a 2-d or 3-d CUDA complex128 tensor"
        raise TypeError(msg)
    if t.is_neg() or any(s < 0 for s in t.stride()):
        t = t.resolve_neg().contiguous()
    bs = t.stride(0) if t.dim() == 3 else 0
    return t, t.stride(-2), t.stride(-1), bs, int(t.is_conj())


def zgemm(
    a: torch.Tensor, b: torch.Tensor, *, alpha: complex = 1.0, beta: complex = 0.0, out: torch.Tensor | None = None,
) -> torch.Tensor:
    """K7: out[..] = alpha * a @ b + beta * out for complex128 [m, k] @ [k, n] or batched [batch, m, k] @ [batch, k, n]
    (a 2-d operand is shared by every batch entry).  Operands may be any strided / conjugated torch views."""
    _require_cuda()
    lib = _lib.load()
    a, a_rs, a_cs, a_bs, a_cj = _gemm_operand(a, "a")
    b, b_rs, b_cs, b_bs, b_cj = _gemm_operand(b, "b")
    m, k = a.shape[-2:]
    k2, n = b.shape[-2:]
    if k != k2:
        msg = f"inner dimensions differ: {tuple(a.shape)} @ {tuple(b.shape)}"
        raise ValueError(msg)
    batched = a.dim() == 3 or b.dim() == 3
    batch = 1
    if batched:
        sizes = {t.shape[0] for t in (a, b) if t.dim() == 3}
        if len(sizes) != 1:
            msg = f"batch sizes differ: {tuple(a.shape)} @ {tuple(b.shape)}"
            raise ValueError(msg)
        batch = sizes.pop()
    shape = (batch, m, n) if batched else (m, n)
    if out is None:
        if beta != 0:
            msg = "beta != 0 needs `out`"
            raise ValueError(msg)
        out = torch.empty(shape, dtype=torch.complex128, device="cuda")
    elif tuple(out.shape) != shape or not (out.is_cuda and out.dtype == torch.complex128 and out.is_contiguous()):
        msg = f"out must be a contiguous CUDA complex128 tensor of shape {shape}"
        raise TypeError(msg)
    al = (ctypes.c_double * 2)(complex(alpha).real, complex(alpha).imag)
    be = (ctypes.c_double * 2)(complex(beta).real, complex(beta).imag)
    rc = lib.sqb_zgemm_batched(
        m, n, k, batch, _ptr(a), a_rs, a_cs, a_bs, a_cj, _ptr(b), b_rs, b_cs, b_bs, b_cj, al, be, _ptr(out), n, m * n,
        _stream(),
    )
    _lib.check(rc, "sqb_zgemm_batched")
    _count(-(-batch // 65535) if m and n else 0)
    return out


def rowwise_vdot(x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
    """out[r] = sum_i conj(x[r, i]) * y[r, i] for [rows, len] complex128 tensors (unit stride along i)."""
    _require_cuda()
    for t, nm in ((x, "x"), (y, "y")):
        if not (t.is_cuda and t.dtype == torch.complex128 and t.dim() == 2 and (t.shape[1] <= 1 or t.stride(1) == 1)):
            msg = f"{nm} must be a [rows, len] CUDA complex128 tensor with unit stride along len"
            raise TypeError(msg)
    if x.shape != y.shape:
        msg = "x and y must have the same shape"
        raise ValueError(msg)
    x, y = x.resolve_conj(), y.resolve_conj()
    rows, length = x.shape
    out = torch.empty((rows,), dtype=torch.complex128, device="cuda")
    rc = _lib.load().sqb_rowwise_vdot(rows, length, _ptr(x), x.stride(0), _ptr(y), y.stride(0), _ptr(out), _stream())
    _lib.check(rc, "sqb_rowwise_vdot")
    _count(1 if rows else 0)
    return out


def rowwise_weighted_abs2(x: torch.Tensor, weights: torch.Tensor) -> torch.Tensor:
    """out[r] = sum_i weights[i] * |x[r, i]|^2 for [rows, len] complex128 states and [len] complex128 weights."""
    _require_cuda()
    if not (x.is_cuda and x.dtype == torch.complex128 and x.dim() == 2 and (x.shape[1] <= 1 or x.stride(1) == 1)):
        msg = "x must be a [rows, len] CUDA complex128 tensor with unit stride along len"
        raise TypeError(msg)
    _c128(weights, "weights")
    if weights.numel() != x.shape[1]:
        msg = "one weight per column"
        raise ValueError(msg)
    x = x.resolve_conj()
    out = torch.empty((x.shape[0],), dtype=torch.complex128, device="cuda")
    rc = _lib.load().sqb_rowwise_weighted_abs2(
        x.shape[0], x.shape[1], _ptr(x), x.stride(0), _ptr(weights), _ptr(out), _stream()
    )
    _lib.check(rc, "sqb_rowwise_weighted_abs2")
    _count(1 if x.shape[0] else 0)
    return out


def scale_columns(x: torch.Tensor, weights: torch.Tensor) -> torch.Tensor:
    """out[r, i] = weights[i] * x[r, i] (a diagonal operator applied to every state of a list)."""
    _require_cuda()
    if not (x.is_cuda and x.dtype == torch.complex128 and x.dim() == 2 and (x.shape[1] <= 1 or x.stride(1) == 1)):
        msg = "x must be a [rows, len] CUDA complex128 tensor with unit stride along len"
        raise TypeError(msg)
    _c128(weights, "weights")
    if weights.numel() != x.shape[1]:
        msg = "one weight per column"
        raise ValueError(msg)
    x = x.resolve_conj()
    out = torch.empty(tuple(x.shape), dtype=torch.complex128, device="cuda")
    rc = _lib.load().sqb_scale_columns(x.shape[0], x.shape[1], _ptr(x), x.stride(0), _ptr(weights), _ptr(out), _stream())
    _lib.check(rc, "sqb_scale_columns")
    _count(1 if x.numel() else 0)
    return out


def abs2(x: torch.Tensor) -> torch.Tensor:
    """|x|^2 elementwise (float64, same shape)."""
    _require_cuda()
    _c128(x, "x")
    out = torch.empty(x.shape, dtype=torch.float64, device="cuda")
    rc = _lib.load().sqb_abs2(x.numel(), _ptr(x), _ptr(out), _stream())
    _lib.check(rc, "sqb_abs2")
    _count(1 if x.numel() else 0)
    return out


def normalize_rows(x: torch.Tensor, norm2: torch.Tensor | None = None) -> torch.Tensor:
    """x[r, :] / sqrt(|norm2[r]|) with norm2 = <x_r|x_r> unless given."""
    _require_cuda()
    _c128(x, "x")
    if x.dim() != 2:
        msg = "x must be [rows, len]"
        raise TypeError(msg)
    if norm2 is None:
        norm2 = rowwise_vdot(x, x)
    _c128(norm2, "norm2")
    out = torch.empty_like(x)
    rc = _lib.load().sqb_normalize_rows(x.shape[0], x.shape[1], _ptr(x), _ptr(norm2), _ptr(out), _stream())
    _lib.check(rc, "sqb_normalize_rows")
    _count(1 if x.numel() else 0)
    return out
