"""Observables on states / state lists (reference: slate_quantum/operator/_measure.py:105-420).

Every quantity here is a diagonal expectation in the position basis, ``sum_x f(x) |psi(x)|^2 / sum_x |psi(x)|^2``
(the reference normalises with ``state.normalize_each`` and contracts a dense diagonal operator).  One CUDA
kernel (``sqb_measure_moments``) reads the ``[T, M]`` state list ONCE and returns the sums for x, x^2 and the
fundamental scattering phase together; nothing is computed on the CPU and only ``[T]`` floats ever leave the
GPU - which is what makes the observables of a 68.7 GB evolved state list cheap to bring back to the host.

Position observables work on any cell (orthogonal axes: table-driven kernel; otherwise the lattice kernel that
decodes every index); the momentum observables need an axis orthogonal to the others.
"""

from __future__ import annotations

from typing import Any

import numpy as np
import torch

from slate_quantum_b200 import kernels
from slate_quantum_b200.core import basis as _basis
from slate_quantum_b200.core.array import Array
from slate_quantum_b200.core.metadata import TupleMetadata, volume
from slate_quantum_b200.state._state import State, StateList


def _tiles_in(states: Array, target: Any) -> Any:
    """Device tensors [rows, M] covering the states in basis `target`, in list order.

    An ordinary Array gives one tensor.  The lazy evolved lists of dynamics.schrodinger (anything with `tiles()`)
    are walked one time tile at a time, so the observables of a list that does not fit in HBM (cfg3: 68.7 GB)
    are reduced on the GPU without the list ever existing."""
    converted = states.with_state_basis(target) if isinstance(states, StateList) else states.with_basis(target)
    size = target.size
    if hasattr(converted, "tiles") and getattr(converted, "_device", None) is None:
        for _, tile in converted.tiles():
            yield tile.reshape(-1, size)
        return
    data = converted.device_data
    if not data.is_complex():
        data = data.to(torch.complex128)
    yield data.reshape(-1, size).contiguous()


def _reduce_position(states: Array, space: TupleMetadata, fn: Any) -> torch.Tensor:
    """cat(fn(tile)) over the position-basis tiles of `states` (fn: [rows, M] -> [rows])."""
    position = _basis.from_metadata(space).upcast()
    return torch.cat([fn(tile) for tile in _tiles_in(states, position)])


def _reduce_momentum(states: Array, space: TupleMetadata, fn: Any) -> torch.Tensor:
    momentum = _basis.transformed_from_metadata(space).upcast()
    return torch.cat([fn(tile) for tile in _tiles_in(states, momentum)])


def _position_data(states: Array, space: TupleMetadata) -> torch.Tensor:
    """[lead, M] device tensor of the states in the fundamental (position) basis of ``space``."""
    return torch.cat(list(_tiles_in(states, _basis.from_metadata(space).upcast())))


def _cell(space: TupleMetadata) -> tuple[np.ndarray, tuple[int, ...]]:
    nd = len(space.children)
    delta_x = np.asarray(volume.fundamental_stacked_delta_x(space), dtype=np.float64)[:nd, :nd]
    return delta_x, tuple(int(c.fundamental_size) for c in space.children)


def _is_orthogonal_axis(space: TupleMetadata, axis: int) -> bool:
    """x along `axis` depends on that axis' index alone (the table-driven kernel applies)."""
    a, _ = _cell(space)
    off_axis = a[axis].copy()
    off_axis[axis] = 0.0
    tol = 1e-12 * np.abs(a).max()
    return bool(np.abs(off_axis).max(initial=0.0) <= tol and np.abs(a[:, axis]).sum() - abs(a[axis, axis]) <= tol)


def _axis_geometry(space: TupleMetadata, axis: int) -> tuple[tuple[int, ...], float, float]:
    a, grid = _cell(space)
    if not _is_orthogonal_axis(space, axis):
        msg = "this observable needs an axis orthogonal to the others on the B200 path"
        raise NotImplementedError(msg)
    return grid, float(a[axis, axis]) / grid[axis], float(np.linalg.norm(a[axis]))


def _moments(pos: torch.Tensor, space: TupleMetadata, axis: int, offsets: torch.Tensor | None, wrapped: bool) -> torch.Tensor:
    """[lead, 5] sums {|psi|^2, x |psi|^2, x^2 |psi|^2, cos, sin} of position-basis states; orthogonal axes use the
    table-driven kernel, general cells (e.g. cfg4's hexagonal lattice) the lattice kernel."""
    if _is_orthogonal_axis(space, axis):
        grid, spacing, length = _axis_geometry(space, axis)
        return kernels.measure_moments(pos, grid, axis, spacing, length, offsets, wrapped)
    delta_x, grid = _cell(space)
    return kernels.measure_moments_lattice(pos, grid, axis, delta_x, offsets, wrapped)


def _list_array(states: StateList, values: torch.Tensor) -> Array:
    list_basis = _basis.from_metadata(states.basis.metadata().children[0])
    return Array(list_basis, values)


def _all_x(pos: torch.Tensor, space: TupleMetadata, axis: int, offset: float, wrapped: bool) -> torch.Tensor:
    offsets = torch.full((pos.shape[0],), float(offset), dtype=torch.float64, device="cuda") if offset else None
    mom = _moments(pos, space, axis, offsets, wrapped)
    return mom[:, 1] / mom[:, 0]


def _all_periodic_x(pos: torch.Tensor, space: TupleMetadata, axis: int) -> torch.Tensor:
    length = float(np.linalg.norm(_cell(space)[0][axis]))
    mom = _moments(pos, space, axis, None, False)
    angle = torch.atan2(mom[:, 4], mom[:, 3])
    return torch.remainder(angle, 2 * np.pi) * (length / (2 * np.pi))


def _all_variance_x(pos: torch.Tensor, space: TupleMetadata, axis: int) -> torch.Tensor:
    centre = _all_periodic_x(pos, space, axis)
    mom = _moments(pos, space, axis, (-centre).contiguous(), True)
    return mom[:, 2] / mom[:, 0]


def _position_bundle(states: StateList, space: TupleMetadata) -> dict[str, list[torch.Tensor]]:
    """<x>, periodic <x> and the variance around it for EVERY axis, from ONE walk over the position-basis tiles.

    The reference computes each observable from the materialised list; here a lazy evolved list would be re-evolved
    per observable, so the first request reduces everything while each tile is resident (two K6 passes per axis and
    tile) and caches the [T] results on the list object."""
    cached = states.__dict__.get("_sqb_measure_cache")
    if cached is not None:
        return cached
    nd = len(space.children)
    position = _basis.from_metadata(space).upcast()
    fused = _fused_bundle(states, space, position)
    if fused is not None:
        states.__dict__["_sqb_measure_cache"] = fused
        return fused
    parts: dict[str, list[list[torch.Tensor]]] = {k: [[] for _ in range(nd)] for k in ("x", "periodic", "variance")}
    for tile in _tiles_in(states, position):
        for axis in range(nd):
            length = float(np.linalg.norm(_cell(space)[0][axis]))
            mom = _moments(tile, space, axis, None, False)
            centre = torch.remainder(torch.atan2(mom[:, 4], mom[:, 3]), 2 * np.pi) * (length / (2 * np.pi))
            mom2 = _moments(tile, space, axis, (-centre).contiguous(), True)
            parts["x"][axis].append(mom[:, 1] / mom[:, 0])
            parts["periodic"][axis].append(centre)
            parts["variance"][axis].append(mom2[:, 2] / mom2[:, 0])
    bundle = {k: [torch.cat(v[axis]) for axis in range(nd)] for k, v in parts.items()}
    states.__dict__["_sqb_measure_cache"] = bundle
    return bundle


def _fused_bundle(states: StateList, space: TupleMetadata, position: Any) -> dict[str, list[torch.Tensor]] | None:
    """The bundle for a LAZY evolved list of a Bloch Hamiltonian on an orthogonal cell: K6 is fused into the last
    pass of the cell FFT (sqb_cell_fft_moments), twice per tile (sums, then the variance around the periodic
    position), so the position-basis states are never written - neither the 68.7 GB list nor a tile of it."""
    import os

    # Opt-in (SQB_MEASURE_FUSED=1).  Measured at cfg3 (profiles/r02_summary.md): the fused last pass reads half the
    # bytes of the plain pass but runs at 3 CTAs / SM (80 registers with the observable sums) and is latency bound
    # like the plain pass (~15 k cycles per 2048-element CTA), so one fused pass costs 37 ms per 4096 states against
    # 27 ms (scatter) + 2 x 10 ms (read passes) - 478 vs 451 ms for the whole observables step.  It halves the tile
    # memory (no position tile), which is what matters when the states barely fit.
    if os.environ.get("SQB_MEASURE_FUSED", "0") != "1":
        return None
    converted = states.with_state_basis(position) if isinstance(states, StateList) else None
    if converted is None or not hasattr(converted, "cell_tiles") or getattr(converted, "_kind", "") != "position":
        return None
    if getattr(converted, "_device", None) is not None:
        return None  # already materialised: plain read passes
    nd = len(space.children)
    if not all(_is_orthogonal_axis(space, a) for a in range(nd)):
        return None
    shape, repeat = converted.bloch_dims()
    geom = [_axis_geometry(space, a) for a in range(nd)]
    spacing = [g[1] for g in geom]
    length = [g[2] for g in geom]
    parts: dict[str, list[list[torch.Tensor]]] = {k: [[] for _ in range(nd)] for k in ("x", "periodic", "variance")}
    scale = torch.tensor([ln / (2 * np.pi) for ln in length], dtype=torch.float64, device="cuda")
    for _, tile in converted.cell_tiles():
        mom = kernels.cell_fft_moments(shape, repeat, tile, spacing, length, stage=0)          # [rows, nd, 5]
        centre = torch.remainder(torch.atan2(mom[:, :, 4], mom[:, :, 3]), 2 * np.pi) * scale[None, :]
        mom2 = kernels.cell_fft_moments(shape, repeat, tile, spacing, length, stage=1,
                                        offsets=(-centre).T.contiguous(), wrapped=True)
        for axis in range(nd):
            parts["x"][axis].append(mom[:, axis, 1] / mom[:, axis, 0])
            parts["periodic"][axis].append(centre[:, axis])
            parts["variance"][axis].append(mom2[:, axis, 2] / mom2[:, axis, 0])
    return {k: [torch.cat(v[axis]) for axis in range(nd)] for k, v in parts.items()}


def _momentum_data(states: Array, space: TupleMetadata) -> torch.Tensor:
    """[lead, M] device tensor of the states in the transformed (momentum) basis of ``space`` (K4 on the GPU)."""
    return torch.cat(list(_tiles_in(states, _basis.transformed_from_metadata(space).upcast())))


def _k_geometry(space: TupleMetadata, axis: int) -> tuple[tuple[int, ...], float, float]:
    grid, spacing, _ = _axis_geometry(space, axis)  # also checks orthogonality
    dk = 2 * np.pi / (spacing * grid[axis])
    return grid, dk, dk * grid[axis]


def _all_k(mom_data: torch.Tensor, space: TupleMetadata, axis: int) -> torch.Tensor:
    # index i <-> k = i dk folded into [-K/2, K/2): the FFT order of the reference's k points
    grid, dk, big_k = _k_geometry(space, axis)
    mom = kernels.measure_moments(mom_data, grid, axis, dk, big_k, None, True)
    return mom[:, 1] / mom[:, 0]


def _all_variance_k(mom_data: torch.Tensor, space: TupleMetadata, axis: int) -> torch.Tensor:
    grid, dk, big_k = _k_geometry(space, axis)
    centre = _all_k(mom_data, space, axis)
    mom = kernels.measure_moments(mom_data, grid, axis, dk, big_k, (-centre).contiguous(), True)
    return mom[:, 2] / mom[:, 0]


def all_k(states: StateList, *, axis: int) -> Array:
    """Momentum of all states (reference operator/_measure.py:323-336)."""
    space = states.basis.metadata().children[1]
    return _list_array(states, _reduce_momentum(states, space, lambda mom: _all_k(mom, space, axis)))


def all_variance_k(states: StateList, *, axis: int) -> Array:
    """<(K - <K>)^2> of all states (reference operator/_measure.py:361-378)."""
    space = states.basis.metadata().children[1]
    return _list_array(states, _reduce_momentum(states, space, lambda mom: _all_variance_k(mom, space, axis)))


def all_uncertainty(states: StateList, *, axis: int) -> Array:
    """sqrt(var_x var_k) of all states (reference operator/_measure.py:400-420)."""
    space = states.basis.metadata().children[1]
    var_x = _reduce_position(states, space, lambda pos: _all_variance_x(pos, space, axis))
    var_k = _reduce_momentum(states, space, lambda mom: _all_variance_k(mom, space, axis))
    return _list_array(states, torch.sqrt(var_x * var_k))


def k(state: State, *, axis: int) -> float:
    """reference operator/_measure.py:307-320."""
    space = state.basis.metadata()
    return float(_all_k(_momentum_data(state, space), space, axis)[0])


def variance_k(state: State, *, axis: int) -> float:
    """reference operator/_measure.py:338-358."""
    space = state.basis.metadata()
    return float(_all_variance_k(_momentum_data(state, space), space, axis)[0])


def uncertainty(state: State, *, axis: int) -> float:
    """reference operator/_measure.py:381-397."""
    return float(np.sqrt(variance_x(state, axis=axis) * variance_k(state, axis=axis)))


def all_x(states: StateList, *, axis: int, offset: float = 0, wrapped: bool = False) -> Array:
    """Position of all states (reference operator/_measure.py:191-210)."""
    space = states.basis.metadata().children[1]
    if not offset and not wrapped:
        return _list_array(states, _position_bundle(states, space)["x"][axis])
    return _list_array(states, _reduce_position(states, space, lambda pos: _all_x(pos, space, axis, offset, wrapped)))


def all_periodic_x(states: StateList, *, axis: int) -> Array:
    """Periodic position coordinate of all states (reference operator/_measure.py:242-268)."""
    space = states.basis.metadata().children[1]
    return _list_array(states, _position_bundle(states, space)["periodic"][axis])


def all_variance_x(states: StateList, *, axis: int) -> Array:
    """<x^2> around each state's periodic position (reference operator/_measure.py:271-292)."""
    space = states.basis.metadata().children[1]
    return _list_array(states, _position_bundle(states, space)["variance"][axis])


def all_coherent_width(states: StateList, *, axis: int) -> Array:
    """Width of a Gaussian wavepacket, sqrt(2 variance) (reference operator/_measure.py:295-305)."""
    space = states.basis.metadata().children[1]
    return _list_array(states, torch.sqrt(2.0 * _position_bundle(states, space)["variance"][axis]))


def x(state: State, *, axis: int, offset: float = 0, wrapped: bool = False) -> float:
    """reference operator/_measure.py:105-121."""
    space = state.basis.metadata()
    return float(_all_x(_position_data(state, space), space, axis, offset, wrapped)[0])


def periodic_x(state: State, *, axis: int) -> float:
    """reference operator/_measure.py:124-138."""
    space = state.basis.metadata()
    return float(_all_periodic_x(_position_data(state, space), space, axis)[0])


def variance_x(state: State, *, axis: int) -> float:
    """reference operator/_measure.py:140-163."""
    space = state.basis.metadata()
    return float(_all_variance_x(_position_data(state, space), space, axis)[0])


def coherent_width(state: State, *, axis: int) -> float:
    """reference operator/_measure.py:166-188."""
    return float(np.sqrt(2.0 * variance_x(state, axis=axis)))


__all__: list[Any] = [
    "all_coherent_width",
    "all_k",
    "all_uncertainty",
    "all_variance_k",
    "all_periodic_x",
    "all_variance_x",
    "all_x",
    "coherent_width",
    "k",
    "periodic_x",
    "uncertainty",
    "variance_k",
    "variance_x",
    "x",
]


# ---------------------------------------------------------------------------------------- generic f(x) / f(k)
def _expectation_over_norm(operator: Any, states: StateList) -> torch.Tensor:
    """real(<psi|O|psi> / <psi|psi>) per state: the reference normalises the states first (state.normalize_each),
    which for a linear expectation is a division of the two one-pass reductions."""
    from slate_quantum_b200.operator._operator import expectation_of_each
    from slate_quantum_b200.state._state import inner_product_each

    value = expectation_of_each(operator, states).device_data
    norm = inner_product_each(states, states).device_data
    return (value / norm).real.contiguous()


def _as_list(state: State) -> StateList:
    return StateList.from_states([state])


def all_potential_from_function(
    states: StateList, fn: Any, *, wrapped: bool = False, offset: tuple[float, ...] | None = None
) -> Array:
    """<f(x)> of every (normalised) state for a generic function of the Cartesian position
    (reference operator/_measure.py:41-61); one weighted pass over the position-basis states."""
    from slate_quantum_b200.operator import build as _build

    space = states.basis.metadata().children[1]
    operator = _build.potential_from_function(space, fn=fn, wrapped=wrapped, offset=offset)
    return _list_array(states, _expectation_over_norm(operator, states))


def potential_from_function(
    state: State, fn: Any, *, wrapped: bool = False, offset: tuple[float, ...] | None = None
) -> float:
    """reference operator/_measure.py:19-38."""
    return float(all_potential_from_function(_as_list(state), fn, wrapped=wrapped, offset=offset).raw_data[0])


def all_momentum_from_function(
    states: StateList, fn: Any, *, wrapped: bool = False, offset: tuple[float, ...] | None = None
) -> Array:
    """<f(k)> of every (normalised) state for a generic function of the Cartesian wavevector (the list form of
    reference operator/_measure.py:64-83); one weighted pass over the momentum-basis states."""
    from slate_quantum_b200.operator import build as _build

    space = states.basis.metadata().children[1]
    operator = _build.momentum_from_function(space, fn=fn, wrapped=wrapped, offset=offset)
    return _list_array(states, _expectation_over_norm(operator, states))


def momentum_from_function(
    state: State, fn: Any, *, wrapped: bool = False, offset: tuple[float, ...] | None = None
) -> float:
    """reference operator/_measure.py:64-83."""
    return float(all_momentum_from_function(_as_list(state), fn, wrapped=wrapped, offset=offset).raw_data[0])
