"""Bloch Hamiltonian builders (reference: slate_quantum/bloch/build.py:32-167).

`kinetic_hamiltonian` is the K1 entry point: instead of the reference's Python loop over k-points
(bloch/build.py:156-161) followed by a per-operator densification inside slate_core (:123-128), every
k-block is written by ONE launch of the block-builder kernel, straight into the [n_k, N, N] raw data of
the BlockDiagonalBasis operator.
"""

from __future__ import annotations

import numpy as np

from slate_quantum_b200 import kernels
from slate_quantum_b200.bloch._shifted_basis import BlochShiftedBasis
from slate_quantum_b200.bloch._transposed_basis import BlochTransposedBasis
from slate_quantum_b200.configs import HBAR as hbar_si
from slate_quantum_b200.core import basis
from slate_quantum_b200.core import metadata as _metadata
from slate_quantum_b200.core.basis import BlockDiagonalBasis, TupleBasis
from slate_quantum_b200.core.metadata import LabeledMetadata, TupleMetadata
from slate_quantum_b200.metadata import RepeatedMetadata
from slate_quantum_b200.operator._diagonal import with_outer_basis
from slate_quantum_b200.operator._operator import Operator, OperatorList


class BlochFractionMetadata(LabeledMetadata):
    """Metadata for the Bloch fraction (reference bloch/build.py:32-49)."""

    def __init__(self, size: int) -> None:
        super().__init__(size)

    @property
    def values(self) -> np.ndarray:
        return _metadata.fundamental_nk_points(self) / self.fundamental_size

    @staticmethod
    def from_repeats(repeat: tuple[int, ...]) -> TupleMetadata:
        return TupleMetadata(tuple(BlochFractionMetadata(n) for n in repeat), None)


def bloch_state_metadata_from_split(fraction_meta: TupleMetadata, state_meta: TupleMetadata) -> TupleMetadata:
    """reference bloch/build.py:57-67."""
    return TupleMetadata(
        tuple(RepeatedMetadata(c, r) for c, r in zip(state_meta.children, fraction_meta.shape, strict=True)), state_meta.extra
    )


def fraction_metadata_from_bloch_state(metadata: TupleMetadata) -> TupleMetadata:
    return BlochFractionMetadata.from_repeats(tuple(c.n_repeats for c in metadata.children))


def basis_from_metadata(metadata: TupleMetadata):
    """Build the Bloch state basis (reference bloch/build.py:79-88)."""
    operator_basis = basis.transformed_from_metadata(metadata)
    return BlochTransposedBasis(
        basis.with_modified_children(operator_basis, lambda _, i: BlochShiftedBasis(i).upcast())
    ).upcast()


def _bloch_operator_basis(fraction_meta: TupleMetadata, cell_meta: TupleMetadata):
    state_meta = bloch_state_metadata_from_split(fraction_meta, cell_meta)
    operator_basis = basis_from_metadata(state_meta)
    n = cell_meta.fundamental_size
    return BlockDiagonalBasis(TupleBasis((operator_basis, operator_basis.dual_basis())), (n, n)).upcast()


def bloch_operator_from_list(operators: OperatorList) -> Operator:
    """Block diagonal Bloch operator from a list of per-k-point operators (reference bloch/build.py:111-135)."""
    meta = operators.basis.metadata()
    operators = operators.with_list_basis(basis.from_metadata(meta.children[0]))
    op_meta = meta.children[1]
    is_dual = basis.as_tuple(operators.basis).children[1].is_dual
    operators = operators.with_operator_basis(basis.transformed_from_metadata(op_meta, is_dual=is_dual).upcast())
    out_basis = _bloch_operator_basis(meta.children[0], op_meta.children[0])
    return Operator(out_basis, operators.device_data.reshape(-1))


def get_sample_fractions(metadata: TupleMetadata) -> tuple[np.ndarray, ...]:
    """Bloch fractions of every k-block, block order = C-order ravel (reference bloch/build.py:138-143)."""
    mesh = np.meshgrid(*(n.values for n in metadata.children), indexing="ij")
    return tuple(nki.ravel() for nki in mesh)


def potential_fourier_table(potential: Operator):
    """The potential's Fourier components zero-padded to the unit-cell grid (device tensor [N])."""
    cell_meta = potential.basis.metadata().children[0]
    transformed = basis.transformed_from_metadata(cell_meta)
    return with_outer_basis(potential, transformed).device_data.reshape(-1).contiguous()


def kinetic_hamiltonian(potential: Operator, mass: float, repeat: tuple[int, ...], *, hbar: float = hbar_si) -> Operator:
    """Kinetic + potential Hamiltonian of every k-block in the Bloch basis (reference bloch/build.py:146-167)."""
    cell_meta = potential.basis.metadata().children[0]
    if any(c.interpolation != "Fourier" for c in cell_meta.children):
        msg = "Bloch Hamiltonians need periodic axes"
        raise NotImplementedError(msg)
    shape = tuple(c.fundamental_size for c in cell_meta.children)
    list_metadata = BlochFractionMetadata.from_repeats(tuple(repeat))
    table = potential_fourier_table(potential)
    dk = _metadata.fundamental_stacked_dk(cell_meta)
    blocks = kernels.bloch_build(shape, tuple(repeat), dk, table, mass, hbar)
    out = Operator(_bloch_operator_basis(list_metadata, cell_meta), blocks.reshape(-1))
    out.bloch_parameters = {"shape": shape, "repeat": tuple(repeat), "dk": dk, "mass": mass, "hbar": hbar}
    return out
