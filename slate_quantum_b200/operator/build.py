"""Operator builders on the hot path (reference: slate_quantum/operator/_build/_hamiltonian.py, _potential.py).

These are host-side constructors of *sparse descriptions* (a few Fourier components, a kinetic
diagonal); the dense arithmetic happens on the GPU when the operators are converted or handed to
`slate_quantum_b200.bloch.build.kinetic_hamiltonian`.
"""

from __future__ import annotations

from typing import Any, Callable

import numpy as np

from slate_quantum_b200._util import outer_product
from slate_quantum_b200.configs import HBAR as hbar_si
from slate_quantum_b200.core import basis
from slate_quantum_b200.core.basis import (
    AsUpcast,
    Basis,
    CroppedBasis,
    SplitBasis,
    TruncatedBasis,
    Truncation,
    TupleBasis,
    transformed_from_metadata,
)
from slate_quantum_b200 import kernels
from slate_quantum_b200.core.metadata import TupleMetadata, fundamental_stacked_dk, volume
from slate_quantum_b200.metadata import repeat_volume_metadata
from slate_quantum_b200.operator._diagonal import momentum_operator_basis, position_operator_basis, with_outer_basis
from slate_quantum_b200.operator._operator import Operator


# --------------------------------------------------------------------------------- kinetic term
def _require_periodic(metadata: TupleMetadata) -> None:
    if any(c.interpolation != "Fourier" for c in metadata.children):
        msg = "only periodic ('Fourier') axes are implemented on the B200 hot path"
        raise NotImplementedError(msg)


def kinetic_basis(metadata: TupleMetadata, *, is_dual: Any = None) -> Basis:
    """reference _hamiltonian.py:100-118 (periodic axes only)."""
    duals = is_dual if isinstance(is_dual, tuple) else tuple(False for _ in metadata.children)
    children = tuple(transformed_from_metadata(c, is_dual=d) for c, d in zip(metadata.children, duals, strict=False))
    return AsUpcast(TupleBasis(children, metadata.extra), metadata)


def kinetic_energy(
    metadata: TupleMetadata, mass: float, bloch_fraction: np.ndarray | None = None, *, hbar: float = hbar_si
) -> Operator:
    """Kinetic part of the Hamiltonian, diagonal in the momentum basis (reference _hamiltonian.py:149-158).

    The diagonal (Bloch phase, Cartesian-row Brillouin-zone wrap of _hamiltonian.py:72-94, hbar^2 k^2 / 2m)
    is evaluated by the CUDA kernel behind `sqb_kinetic_energy`; this function only assembles the basis.
    """
    _require_periodic(metadata)
    shape = tuple(c.fundamental_size for c in metadata.children)
    energy = kernels.kinetic_energy(shape, fundamental_stacked_dk(metadata), bloch_fraction, mass, hbar)
    momentum_basis = kinetic_basis(metadata)
    return Operator(AsUpcast(momentum_operator_basis(momentum_basis), TupleMetadata((metadata, metadata))), energy)


def kinetic_hamiltonian(
    potential: Operator, mass: float, bloch_fraction: np.ndarray | None = None, *, hbar: float = hbar_si
) -> Operator:
    """SplitBasis(potential, kinetic) Hamiltonian with concatenated raw data (reference _hamiltonian.py:161-182)."""
    metadata = potential.basis.metadata().children[0]
    kinetic = kinetic_energy(metadata, mass, bloch_fraction, hbar=hbar)
    split = SplitBasis(potential.basis, kinetic.basis)
    return Operator(
        AsUpcast(split, TupleMetadata((metadata, metadata))),
        np.concatenate([potential.raw_data, kinetic.raw_data], axis=None),
    )


# --------------------------------------------------------------------------------- potentials
def potential(outer_basis: Basis, data: np.ndarray) -> Operator:
    """reference _potential.py:46-50."""
    return Operator(position_operator_basis(outer_basis), np.asarray(data).ravel())


_build_potential = potential


def _cropped(metadata: TupleMetadata, n_terms: tuple[int | None, ...]) -> tuple[Basis, Basis]:
    transformed_basis = basis.transformed_from_metadata(metadata)
    cropped = basis.with_modified_children(
        transformed_basis, lambda i, y: y if n_terms[i] is None else CroppedBasis(n_terms[i], y).upcast()
    ).upcast()
    return transformed_basis, cropped


def cos_potential(metadata: TupleMetadata, height: float, *, offset: float = 0) -> Operator:
    """reference _potential.py:86-108: height * prod_a (1 + offset + cos(2 pi x_a / L_a)) / 2."""
    transformed_basis, cropped = _cropped(metadata, (3,) * len(metadata.children))
    n_dim = len(metadata.children)
    data = outer_product(*(np.array([2 * (1 + offset), 1, 1], dtype=np.complex128),) * n_dim)
    return potential(cropped, 0.25**n_dim * height * data * np.sqrt(transformed_basis.size))


def sin_potential(metadata: TupleMetadata, height: float, *, offset: float = 0) -> Operator:
    """reference _potential.py:111-131."""
    transformed_basis, cropped = _cropped(metadata, (3,) * len(metadata.children))
    n_dim = len(metadata.children)
    data = outer_product(*(np.array([2 * (1 + offset), 1j, -1j]),) * n_dim)
    return potential(cropped, 0.25**n_dim * height * data * np.sqrt(transformed_basis.size))


def _square_wave_points(n_terms: int, lanczos_factor: float) -> np.ndarray:
    prefactor = np.sinc(2 * np.fft.fftfreq(n_terms)) ** lanczos_factor
    return prefactor * np.fft.ifft(np.sign(np.cos(np.linspace(0, 2 * np.pi, n_terms, endpoint=False))))


def square_potential(
    metadata: TupleMetadata, height: float, *, n_terms: tuple[int, ...] | None = None, lanczos_factor: float = 0
) -> Operator:
    """reference _potential.py:143-164."""
    terms = (None,) * len(metadata.children) if n_terms is None else n_terms
    transformed_basis, cropped = _cropped(metadata, terms)
    data = outer_product(*(_square_wave_points(n, lanczos_factor) for n in basis.as_tuple(cropped).shape))
    return potential(cropped, 0.5 * height * data * np.sqrt(transformed_basis.size))


def fcc_potential(metadata: TupleMetadata, height: float) -> Operator:
    """reference _potential.py:167-190 (2-D only, as in the reference)."""
    transformed_basis, cropped = _cropped(metadata, (3,) * len(metadata.children))
    n_dim = len(metadata.children)
    assert n_dim == 2  # noqa: PLR2004
    data = np.array([[3, 1, 1], [1, 1, 0], [1, 0, 1]]).astype(np.complex128)
    return potential(cropped, (1 / 3) ** n_dim * data * height * np.sqrt(transformed_basis.size))


def potential_from_function(
    metadata: TupleMetadata,
    fn: Callable[[tuple[np.ndarray, ...]], np.ndarray],
    *,
    wrapped: bool = False,
    offset: tuple[float, ...] | None = None,
) -> Operator:
    """reference _potential.py:193-216."""
    positions = volume.fundamental_stacked_x_points(metadata, offset=offset, wrapped=wrapped)
    points = np.asarray(fn(positions)).ravel().astype(np.complex128)
    return potential(basis.from_metadata(metadata).upcast(), points)


def repeat_potential(potential_: Operator, shape: tuple[int, ...]) -> Operator:
    """Repeat a unit-cell potential over a supercell (reference _potential.py:56-83)."""
    unit_meta = potential_.basis.inner.outer_recast.metadata()
    transformed_basis = basis.transformed_from_metadata(unit_meta)
    as_transformed = with_outer_basis(potential_, transformed_basis)
    converted_basis = basis.transformed_from_metadata(repeat_volume_metadata(unit_meta, shape))
    repeat_basis = basis.with_modified_children(
        converted_basis,
        lambda i, y: TruncatedBasis(Truncation(transformed_basis.children[i].size, shape[i], 0), y).upcast(),
    ).upcast()
    return _build_potential(repeat_basis, as_transformed.raw_data * np.sqrt(np.prod(shape)))


# --------------------------------------------------------------------------------- x / k / p operators
def momentum(outer_basis: Basis, data: np.ndarray) -> Operator:
    """Operator diagonal in the momentum basis (reference _build/_momentum.py:30-34)."""
    return Operator(momentum_operator_basis(outer_basis), np.asarray(data).ravel())


def x(metadata: TupleMetadata, *, axis: int, offset: float = 0, wrapped: bool = False) -> Operator:
    """Cartesian position along `axis`, diagonal in the position basis (reference _build/_position.py:228-242)."""
    inner_offset = tuple(offset if i == axis else 0 for i in range(len(metadata.children)))
    points = volume.fundamental_stacked_x_points(metadata, offset=inner_offset, wrapped=wrapped)[axis]
    return potential(basis.from_metadata(metadata).upcast(), points.astype(np.complex128))


def k(metadata: TupleMetadata, *, axis: int) -> Operator:
    """Cartesian wavevector along `axis`, diagonal in the momentum basis (reference _build/_momentum.py:37-48)."""
    _require_periodic(metadata)
    points = volume.fundamental_stacked_k_points(metadata)[axis].astype(np.complex128)
    return momentum(basis.transformed_from_metadata(metadata).upcast(), points)


def k_squared(metadata: TupleMetadata, *, axis: int) -> Operator:
    """reference _build/_momentum.py:51-62."""
    _require_periodic(metadata)
    points = volume.fundamental_stacked_k_points(metadata)[axis].astype(np.complex128)
    return momentum(basis.transformed_from_metadata(metadata).upcast(), points**2)


def p(metadata: TupleMetadata, *, axis: int, hbar: float = hbar_si) -> Operator:
    """hbar k (reference _build/_momentum.py:65-79)."""
    _require_periodic(metadata)
    points = volume.fundamental_stacked_k_points(metadata)[axis].astype(np.complex128)
    return momentum(basis.transformed_from_metadata(metadata).upcast(), (hbar * points).astype(np.complex128))


def momentum_from_function(
    metadata: TupleMetadata,
    fn: Callable[[tuple[np.ndarray, ...]], np.ndarray],
    *,
    wrapped: bool = False,
    offset: tuple[float, ...] | None = None,
) -> Operator:
    """f(k) as an operator diagonal in the momentum basis (reference _build/_momentum.py:82-98)."""
    _require_periodic(metadata)
    positions = volume.fundamental_stacked_k_points(metadata, offset=offset, wrapped=wrapped)
    points = np.asarray(fn(positions)).ravel().astype(np.complex128)
    return momentum(basis.transformed_from_metadata(metadata).upcast(), points)
