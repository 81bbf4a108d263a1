"""Schrodinger evolution by eigen-decomposition (reference: slate_quantum/dynamics/schrodinger.py:31-73).

The reference materialises ``coefficients[None, :] * np.exp(-1j * E * t / hbar)`` ([T, M] complex128) and
leaves the back-transformation to the caller's ``with_state_basis`` (state/_state.py:153-165).  Here both
StateLists are LAZY:

* ``solve_schrodinger_equation_decomposition`` returns an ``EvolvedStateList`` holding (c, E, t) on the GPU;
  ``raw_data`` materialises the eigenbasis amplitudes with one kernel (sqb_evolve_eigen).
* ``EvolvedStateList.with_state_basis(basis)`` returns a ``TiledEvolvedStates``: the states are produced one
  TIME TILE at a time by the fused phase x eigenvector DMMA GEMM (sqb_evolve_bloch[_uniform]) - the eigenbasis
  array is never written.  For the supercell POSITION basis of a Bloch Hamiltonian the eigenvectors are first
  folded into the in-cell position basis with the Bloch twiddle (sqb_fold_transform, once per eigensolve) so a
  tile only needs the R-point FFTs across blocks (sqb_cell_fft) - the route bench.py measures.  A tile is sized to
  the free HBM; ``device_data`` assembles the full [T, M] list when it fits, ``tiles()`` / ``host_tiles()`` /
  ``copy_to_host()`` stream it when it does not (cfg3: 68.7 GB of states + tiles; cfg5: 92 GB).
"""

from __future__ import annotations

from collections import deque
from typing import Any, Callable, Iterator

import numpy as np
import torch

from slate_quantum_b200 import kernels
from slate_quantum_b200.bloch._transposed_basis import BlochTransposedBasis
from slate_quantum_b200.configs import HBAR as hbar_value
from slate_quantum_b200.core import array
from slate_quantum_b200.core import basis as _basis
from slate_quantum_b200.core.basis import Basis, TupleBasis, _strip
from slate_quantum_b200.operator._linalg import into_diagonal_hermitian
from slate_quantum_b200.operator._operator import Operator
from slate_quantum_b200.state._state import State, StateList


# Rows of a time tile are kept a multiple of this: a CTA of the uniform-grid evolve kernel walks 8 consecutive
# 64-row tiles (csrc/evolve.cu), so tiles of k * 512 rows have no ragged tail.
_TILE_QUANTUM = 512


def _tile_rows(n_times: int, row_bytes: int, n_buffers: int, reserve_bytes: int = 0, fraction: float = 0.8) -> int:
    """Largest tile (rows) such that `n_buffers` [rows, M] buffers fit in `fraction` of the free HBM."""
    free = kernels.free_bytes()
    budget = max(int(free * fraction) - reserve_bytes, 0)
    rows = budget // max(n_buffers * row_bytes, 1)
    if rows >= n_times:
        return n_times
    if rows >= _TILE_QUANTUM:
        return rows // _TILE_QUANTUM * _TILE_QUANTUM
    if rows < 1:
        msg = (
            f"not enough free HBM for one evolved state per buffer ({row_bytes / 1e9:.2f} GB each, "
            f"{free / 1e9:.1f} GB free)"
        )
        raise MemoryError(msg)
    return rows


# Pinned host staging buffers are expensive to create (cudaHostAlloc of 1 GiB takes ~0.1 s): host_tiles() keeps the
# ones it made, keyed by shape, and reuses them in later calls.
_PINNED_POOL: dict[tuple[int, int], list[torch.Tensor]] = {}


def _pinned_buffers(rows: int, width: int, count: int) -> list[torch.Tensor]:
    pool = _PINNED_POOL.setdefault((rows, width), [])
    while len(pool) < count:
        pool.append(torch.empty((rows, width), dtype=torch.complex128).pin_memory())
    return pool[:count]


class TiledEvolvedStates(StateList):
    """``EvolvedStateList.with_state_basis(basis)``: [T, M] states in `basis`, generated tile by tile on the GPU.

    kind "bloch": the eigenbasis' own inner (Bloch-ordered momentum) basis - K3c only;
    kind "position": supercell position basis - K3c on the folded eigenvectors + K4b cell FFT;
    kind "momentum": supercell momentum basis - K3c + the K5 permutation;
    kind "generic": K3c + the inner basis' conversion of every tile.
    """

    def __init__(self, basis: Basis, source: "EvolvedStateList", kind: str, target: Basis) -> None:
        self._basis = basis
        self._host = None
        self._device = None
        self._source = source
        self._kind = kind
        self._target = target
        eig = source._eigenbasis()  # noqa: SLF001
        self._blocks = eig._blocks()  # noqa: SLF001  [n_blocks, n, n], rows = eigenvectors
        self._inner = _strip(eig.inner)
        self._eig = eig
        nb, n, _ = self._blocks.shape
        self._w = source._eigenvalues.reshape(nb, n).contiguous()  # noqa: SLF001
        self._c = source._coefficients.reshape(nb, n).contiguous()  # noqa: SLF001
        self._evolve_ws: torch.Tensor | None = None
        self._ws_primed = False

    # ------------------------------------------------------------------ sizes
    @property
    def n_times(self) -> int:
        return int(self._source._times.numel())  # noqa: SLF001

    @property
    def state_size(self) -> int:
        return int(self._blocks.shape[0] * self._blocks.shape[1])

    @property
    def row_bytes(self) -> int:
        return self.state_size * 16

    def _n_scratch(self) -> int:
        """[rows, M] scratch buffers a tile needs besides its output."""
        return {"bloch": 0, "position": 1, "momentum": 1, "generic": 1}[self._kind]

    # ------------------------------------------------------------------ one tile
    def _transform(self) -> torch.Tensor:
        if self._kind != "position":
            return self._blocks
        # eigenvectors in the in-cell position basis x Bloch twiddle, cached on the eigenbasis (once per eigh)
        cache = self._eig.transform()
        folded = getattr(cache, "_sqb_folded", None)
        if folded is None:
            shape, repeat = self._inner._dims()  # noqa: SLF001
            folded = kernels.fold_transform(shape, repeat, self._blocks)
            cache._sqb_folded = folded  # noqa: SLF001
        return folded

    def _evolve(self, t0: int, tt: int, out: torch.Tensor) -> torch.Tensor:
        src = self._source
        times = src._times[t0 : t0 + tt]  # noqa: SLF001
        tr = self._transform()
        dt = src._uniform_dt  # noqa: SLF001
        if dt is None:
            return kernels.evolve_bloch(tr, self._w, self._c, times, src._hbar, out=out)  # noqa: SLF001
        nb, n, _ = tr.shape
        need = kernels.evolve_uniform_workspace_bytes(n, nb, tt)
        if self._evolve_ws is None or self._evolve_ws.numel() < need:
            self._evolve_ws = torch.empty(need, dtype=torch.uint8, device="cuda")
            self._ws_primed = False
        res = kernels.evolve_bloch(
            tr, self._w, self._c, times, src._hbar, out=out, uniform_dt=dt,  # noqa: SLF001
            workspace=self._evolve_ws, reuse_plane=self._ws_primed,
        )
        self._ws_primed = True  # the Re+Im plane of THIS transform now sits in the workspace
        return res

    def _fill(self, t0: int, tt: int, out: torch.Tensor, scratch: torch.Tensor | None) -> torch.Tensor:
        """Rows [t0, t0+tt) of the list into `out` ([tt, M], contiguous)."""
        if self._kind == "bloch":
            return self._evolve(t0, tt, out)
        assert scratch is not None
        states = self._evolve(t0, tt, scratch[:tt])
        shape, repeat = self._inner._dims() if hasattr(self._inner, "_dims") else (None, None)  # noqa: SLF001
        if self._kind == "position":
            return kernels.cell_fft(shape, repeat, states, 1, out=out)
        if self._kind == "momentum":
            return kernels.bloch_permute(shape, repeat, states, 1, out=out)
        out.copy_(self._inner.convert_vector_into(states, self._target, -1).reshape(tt, -1))
        return out

    # ------------------------------------------------------------------ streaming access
    def tile_rows(self, n_buffers: int | None = None, reserve_bytes: int = 0) -> int:
        n_buffers = (1 + self._n_scratch()) if n_buffers is None else n_buffers
        return _tile_rows(self.n_times, self.row_bytes, n_buffers, reserve_bytes)

    def tiles(self, rows: int | None = None) -> Iterator[tuple[int, torch.Tensor]]:
        """Yield (t0, states[tt, M]) device tiles in time order.  The tensor is REUSED by the next iteration."""
        if self._device is not None:  # already materialised
            yield 0, self._device.reshape(self.n_times, self.state_size)
            return
        self._transform()  # fold before sizing the tiles
        rows = self.tile_rows() if rows is None else int(rows)
        m = self.state_size
        out = torch.empty((rows, m), dtype=torch.complex128, device="cuda")
        scratch = torch.empty((rows, m), dtype=torch.complex128, device="cuda") if self._n_scratch() else None
        for t0 in range(0, self.n_times, rows):
            tt = min(rows, self.n_times - t0)
            yield t0, self._fill(t0, tt, out[:tt], scratch)

    def cell_tiles(self, rows: int | None = None) -> Iterator[tuple[int, torch.Tensor]]:
        """kind "position" only: yield (t0, tile[tt, M]) in the CELL layout (block, in-cell position) - the evolve
        output before the across-block FFT.  Consumers that only reduce the position-basis states
        (kernels.cell_fft_moments: K6 fused into the last FFT pass) never need them written.  The tensor is reused
        and may be clobbered by the consumer."""
        if self._kind != "position":
            msg = "cell_tiles() needs the supercell position basis of a Bloch Hamiltonian"
            raise ValueError(msg)
        self._transform()
        rows = self.tile_rows(1) if rows is None else int(rows)
        scratch = torch.empty((rows, self.state_size), dtype=torch.complex128, device="cuda")
        for t0 in range(0, self.n_times, rows):
            tt = min(rows, self.n_times - t0)
            yield t0, self._evolve(t0, tt, scratch[:tt])

    def bloch_dims(self) -> tuple[tuple[int, ...], tuple[int, ...]]:
        return self._inner._dims()  # noqa: SLF001

    def host_tiles(
        self, rows: int | None = None, chunk_bytes: int = 1 << 30, n_host_buffers: int = 3,
    ) -> Iterator[tuple[int, torch.Tensor]]:
        """Yield (t0, states[rr, M]) as PINNED HOST tensors in time order (device -> host overlapped with the
        evolution of the next tile).  A yielded tensor stays valid for the next `n_host_buffers - 1` iterations."""
        self._transform()
        m, row_bytes = self.state_size, self.row_bytes
        host_rows = max(1, chunk_bytes // row_bytes)
        if rows is None:
            # two device tiles ping-pong (+ one scratch); keep them modest so the first D2H starts early
            cap = max(host_rows, (4 << 30) // row_bytes)
            rows = min(self.tile_rows(2 + self._n_scratch()), cap, self.n_times)
        host_rows = min(host_rows, rows)
        dev = [torch.empty((rows, m), dtype=torch.complex128, device="cuda") for _ in range(2)]
        scratch = torch.empty((rows, m), dtype=torch.complex128, device="cuda") if self._n_scratch() else None
        stage = _pinned_buffers(host_rows, m, n_host_buffers)
        copy_stream = torch.cuda.Stream()
        cur = torch.cuda.current_stream()
        tile_free = [None, None]
        pending: deque = deque()  # (event, t0, host view) in time order
        k = 0
        # tile boundaries: the first tiles ramp up (64 rows = one time tile of the evolve kernel, then x 2) so that the
        # device -> host stream - the bottleneck of the whole call at config sizes - starts as soon as a few rows exist
        bounds, t_next, ramp = [], 0, min(rows, 64)
        while t_next < self.n_times:
            tt = min(ramp, self.n_times - t_next)
            bounds.append((t_next, tt))
            t_next += tt
            ramp = min(rows, ramp * 2)
        for i_tile, (t0, tt) in enumerate(bounds):
            buf = dev[i_tile & 1]
            if tile_free[i_tile & 1] is not None:
                cur.wait_event(tile_free[i_tile & 1])  # the D2H of the tile that used this buffer is done
            self._fill(t0, tt, buf[:tt], scratch)
            ready = torch.cuda.Event()
            ready.record(cur)
            copy_stream.wait_event(ready)
            for r0 in range(0, tt, host_rows):
                rr = min(host_rows, tt - r0)
                while len(pending) >= n_host_buffers:  # the staging buffer about to be reused must be consumed
                    ev, pt0, view = pending.popleft()
                    ev.synchronize()
                    yield pt0, view
                with torch.cuda.stream(copy_stream):
                    view = stage[k % n_host_buffers][:rr]
                    view.copy_(buf[r0 : r0 + rr], non_blocking=True)
                    ev = torch.cuda.Event()
                    ev.record(copy_stream)
                pending.append((ev, t0 + r0, view))
                k += 1
            done = torch.cuda.Event()
            done.record(copy_stream)
            tile_free[i_tile & 1] = done
        while pending:
            ev, pt0, view = pending.popleft()
            ev.synchronize()
            yield pt0, view
        cur.wait_stream(copy_stream)

    def copy_to_host(self, out: Any, rows: int | None = None) -> Any:
        """Stream the whole list into a host array `out` ([T, M] or flat; numpy or torch, pinned or pageable)."""
        target = torch.from_numpy(out) if isinstance(out, np.ndarray) else out
        target = target.reshape(self.n_times, self.state_size)
        if target.is_pinned():
            for t0, tile in self.tiles(rows):
                target[t0 : t0 + tile.shape[0]].copy_(tile, non_blocking=True)
            torch.cuda.synchronize()
            return out
        for t0, view in self.host_tiles(rows):
            target[t0 : t0 + view.shape[0]].copy_(view)
        return out

    def for_each_tile(self, fn: Callable[[int, torch.Tensor], Any], rows: int | None = None) -> list[Any]:
        """fn(t0, device tile) for every tile; the results in time order (K6-style reductions at scale)."""
        return [fn(t0, tile) for t0, tile in self.tiles(rows)]

    # ------------------------------------------------------------------ Array protocol
    @property
    def device_data(self) -> torch.Tensor:
        """The full [T * M] list on the device (raises MemoryError when it cannot fit; use tiles() then)."""
        if self._device is None:
            self._transform()
            total = self.n_times * self.row_bytes
            free = kernels.free_bytes()
            if total + self._n_scratch() * self.row_bytes > free:
                msg = (
                    f"the evolved state list needs {total / 1e9:.1f} GB but {free / 1e9:.1f} GB of HBM are free: "
                    "iterate tiles() / host_tiles() or copy_to_host() instead of materialising it"
                )
                raise MemoryError(msg)
            torch.cuda.empty_cache()  # one big block: hand the allocator's cached pieces back first
            full = torch.empty((self.n_times, self.state_size), dtype=torch.complex128, device="cuda")
            rows = self.n_times
            scratch = None
            if self._n_scratch():
                rows = _tile_rows(self.n_times, self.row_bytes, 1)
                scratch = torch.empty((rows, self.state_size), dtype=torch.complex128, device="cuda")
            for t0 in range(0, self.n_times, rows):
                tt = min(rows, self.n_times - t0)
                self._fill(t0, tt, full[t0 : t0 + tt], scratch)
            self._device = full.reshape(-1)
        return self._device

    @property
    def raw_data(self) -> np.ndarray:
        if self._host is None:
            if self._device is not None:
                self._host = self._device.detach().cpu().numpy()
            else:
                host = np.empty(self.n_times * self.state_size, dtype=np.complex128)
                self.copy_to_host(host)
                self._host = host
        return self._host

    @property
    def is_on_device(self) -> bool:
        return True

    def state(self, index: int) -> State:
        """One state of the list (a single time row is evolved; nothing else is materialised)."""
        index = int(index) % self.n_times
        m = self.state_size
        out = torch.empty((1, m), dtype=torch.complex128, device="cuda")
        scratch = torch.empty((1, m), dtype=torch.complex128, device="cuda") if self._n_scratch() else None
        # a one-row tile of a uniform grid: the general kernel (the table-driven one wants the grid's own rows)
        src = self._source
        dt, src._uniform_dt = src._uniform_dt, None  # noqa: SLF001
        try:
            self._fill(index, 1, out, scratch)
        finally:
            src._uniform_dt = dt  # noqa: SLF001
        return State(self._target, out.reshape(-1))

    def __getitem__(self, index: Any) -> Any:
        if isinstance(index, tuple) and isinstance(index[0], (int, np.integer)) and index[1] == slice(None):
            return self.state(index[0])
        if isinstance(index, (int, np.integer)):
            return self.state(index)
        msg = "only list[i] / list[i, :] indexing is implemented"
        raise NotImplementedError(msg)

    def __iter__(self):
        for _, tile in self.tiles():
            for row in tile:
                yield State(self._target, row.clone())

    def with_state_basis(self, basis: Basis) -> StateList:
        if basis == self._target:
            return self
        return self._source.with_state_basis(basis)

    def with_basis(self, basis: Basis) -> StateList:
        tup = _basis.as_tuple(basis)
        if tup.children[0] == _basis.as_tuple(self.basis).children[0]:
            return self.with_state_basis(tup.children[1])
        return StateList(self.basis, self.device_data).with_basis(basis)


class EvolvedStateList(StateList):
    """StateList over (times, eigenbasis) whose [T, M] data is generated on demand."""

    def __init__(
        self, basis: Basis, coefficients: torch.Tensor, eigenvalues: torch.Tensor, times: torch.Tensor, hbar: float,
        uniform_dt: float | None = None,
    ) -> None:
        self._basis = basis
        self._host = None
        self._device = None
        self._coefficients = coefficients
        self._eigenvalues = eigenvalues  # [n_blocks, n] float64
        self._times = times
        self._hbar = hbar
        self._uniform_dt = uniform_dt  # spacing of a uniform time grid: enables the table-driven 3M GEMM

    @property
    def device_data(self) -> torch.Tensor:
        if self._device is None:
            self._device = kernels.evolve_eigen(
                self._eigenvalues.reshape(-1), self._coefficients.reshape(-1), self._times, self._hbar
            ).reshape(-1)
        return self._device

    @property
    def raw_data(self) -> np.ndarray:
        if self._host is None:
            self._host = self.device_data.detach().cpu().numpy().reshape(self._times.numel(), -1)
        return self._host

    @property
    def is_on_device(self) -> bool:
        return True

    def _eigenbasis(self):
        return _strip(_basis.as_tuple(self.basis).children[1])

    def _kind_for(self, inner: Basis, basis: Basis) -> str:
        target = _strip(basis)
        if inner == target:
            return "bloch"
        if isinstance(inner, BlochTransposedBasis) and inner._standard_children() and not inner._is_dual_any():  # noqa: SLF001
            if target == _strip(_basis.from_metadata(inner.metadata())) and inner._cell_path():  # noqa: SLF001
                return "position"
            if target == inner.supercell_momentum_basis():
                return "momentum"
        return "generic"

    def with_state_basis(self, basis: Basis) -> StateList:
        """Back-transformation of the evolved states (reference state/_state.py:153-165), lazily and in time
        tiles: see TiledEvolvedStates."""
        eig = self._eigenbasis()
        if eig == _strip(basis):
            return self
        if eig.is_dual:
            return StateList(self.basis, self.device_data).with_state_basis(basis)
        inner = _strip(eig.inner)
        times_basis = _basis.as_tuple(self.basis).children[0]
        out_basis = TupleBasis((times_basis, basis)).upcast()
        return TiledEvolvedStates(out_basis, self, self._kind_for(inner, basis), basis)

    def with_basis(self, basis: Basis) -> StateList:
        tup = _basis.as_tuple(basis)
        if tup.children[0] == _basis.as_tuple(self.basis).children[0]:
            return self.with_state_basis(tup.children[1])
        return StateList(self.basis, self.device_data).with_basis(basis)


def _time_values(times: Basis) -> np.ndarray:
    """reference schrodinger.py:50: take the values from the metadata object, never recompute them."""
    values = np.array(list(times.metadata().values))
    points = getattr(_strip(times), "points", np.arange(times.size))
    return values[points]


def _solve_schrodinger_equation_diagonal(
    initial_state: State, times: Basis, hamiltonian: Any, *, hbar: float = hbar_value
) -> StateList:
    """reference schrodinger.py:31-56."""
    eigenbasis = hamiltonian.basis.inner.children[0]
    coefficients = initial_state.with_basis(eigenbasis).device_data.reshape(-1)
    n_total = coefficients.numel()
    eig_real = getattr(hamiltonian, "eigenvalues_real", None)
    if eig_real is None:
        eig_real = hamiltonian.device_data.real.contiguous()
    blocks = _strip(eigenbasis)._blocks()  # noqa: SLF001
    eig_real = eig_real.reshape(blocks.shape[0], n_total // blocks.shape[0]).contiguous()
    host_times = _time_values(times).astype(np.float64)
    time_values = kernels.to_device(host_times, torch.float64)
    out_basis = TupleBasis((times, eigenbasis)).upcast()
    return EvolvedStateList(out_basis, coefficients, eig_real, time_values, hbar, kernels.uniform_step(host_times))


def solve_schrodinger_equation_decomposition(
    initial_state: State, times: Basis, hamiltonian: Operator, *, hbar: float = hbar_value
) -> StateList:
    """Solve the Schrodinger equation by directly finding eigenstates (reference schrodinger.py:59-73)."""
    diagonal = into_diagonal_hermitian(hamiltonian)
    cast = array.cast_basis(diagonal, diagonal.basis.inner)
    cast.eigenvalues_real = diagonal.eigenvalues_real
    return _solve_schrodinger_equation_diagonal(initial_state, times, cast, hbar=hbar)


def solve_schrodinger_equation(*_args: Any, **_kwargs: Any) -> StateList:
    """The qutip ODE path of the reference (schrodinger.py:76-123) is out of scope."""
    msg = "solve_schrodinger_equation (qutip sesolve) is not part of the B200 hot path; use the decomposition solver"
    raise NotImplementedError(msg)


__all__ = [
    "BlochTransposedBasis",
    "EvolvedStateList",
    "TiledEvolvedStates",
    "solve_schrodinger_equation",
    "solve_schrodinger_equation_decomposition",
]
