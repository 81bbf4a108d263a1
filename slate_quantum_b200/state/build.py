"""State builders (reference: slate_quantum/state/_build.py:79-143)."""

from __future__ import annotations

from typing import Any, Callable

import numpy as np

from slate_quantum_b200.core import basis
from slate_quantum_b200.core.metadata import TupleMetadata, volume
from slate_quantum_b200.state._state import State, normalize


def coherent(
    metadata: TupleMetadata, x_0: tuple[float, ...], k_0: tuple[float, ...], sigma_0: tuple[float, ...]
) -> State:
    """Gaussian wavepacket exp(i k0.(x-x0) - |x-x0|^2 / 2 sigma^2), periodic displacement (reference :79-99)."""
    a = volume.fundamental_stacked_delta_x(metadata)
    nd = len(metadata.children)
    a = a[:, :nd]
    fractions = np.meshgrid(
        *(np.arange(c.fundamental_size) / c.fundamental_size for c in metadata.children), indexing="ij"
    )
    frac = np.array([f.ravel() for f in fractions])
    frac0 = np.linalg.solve(a.T, np.asarray(x_0, dtype=np.float64)[:nd])
    wrapped = (frac - frac0[:, None] + 0.5) % 1.0 - 0.5
    disp = np.einsum("ij,ik->jk", a, wrapped)
    distance = np.linalg.norm([d / s for d, s in zip(disp, sigma_0, strict=False)], axis=0)
    phi = np.einsum("ij,i->j", disp, np.asarray(k_0, dtype=np.float64)[:nd])
    data = np.exp(1j * phi - np.square(distance) / 2)
    norm = np.sqrt(np.sum(np.square(np.abs(data))))
    return State(basis.from_metadata(metadata).upcast(), data / norm)


def position(metadata: TupleMetadata, idx: tuple[int, ...]) -> State:
    data = np.zeros(metadata.shape, dtype=np.complex128)
    data[tuple(i % n for (i, n) in zip(idx, metadata.shape, strict=True))] = 1.0
    return State(basis.from_metadata(metadata).upcast(), data.ravel())


def momentum(metadata: TupleMetadata, idx: tuple[int, ...]) -> State:
    data = np.zeros(metadata.shape, dtype=np.complex128)
    data[tuple(i % n for (i, n) in zip(idx, metadata.shape, strict=True))] = 1.0
    return State(basis.transformed_from_metadata(metadata).upcast(), data.ravel())


def from_function(
    metadata: TupleMetadata,
    fn: Callable[[tuple[np.ndarray, ...]], np.ndarray],
    *,
    normalize_state: bool = True,
    offset: tuple[float, ...] | None = None,
    wrapped: bool = False,
) -> State:
    positions = volume.fundamental_stacked_x_points(metadata, offset=offset, wrapped=wrapped)
    result = State(basis.from_metadata(metadata).upcast(), np.asarray(fn(positions), dtype=np.complex128).ravel())
    return normalize(result) if normalize_state else result


__all__: list[Any] = ["coherent", "from_function", "momentum", "position"]
