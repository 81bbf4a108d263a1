"""Diagonal operator bases (reference: slate_quantum/operator/_diagonal.py:64-207)."""

from __future__ import annotations

from slate_quantum_b200.core import basis as _basis
from slate_quantum_b200.core.basis import AsUpcast, Basis, DiagonalBasis, RecastBasis, TupleBasis
from slate_quantum_b200.core.metadata import TupleMetadata
from slate_quantum_b200.operator._operator import Operator


def recast_diagonal_basis(inner_recast: Basis, outer_recast: Basis) -> AsUpcast:
    """reference :64-73: the diagonal is stored on the DUAL outer basis."""
    inner = DiagonalBasis(TupleBasis((inner_recast, inner_recast.dual_basis()))).upcast()
    recast = RecastBasis(inner, inner_recast.dual_basis(), outer_recast.dual_basis())
    return AsUpcast(recast, inner.metadata())


recast_diagonal_basis_with_metadata = recast_diagonal_basis


def with_outer_basis(operator: Operator, outer_basis: Basis) -> Operator:
    """reference :98-103."""
    basis = recast_diagonal_basis(operator.basis.inner.inner_recast, outer_basis)
    return operator.with_basis(basis)


def position_operator_basis(basis: Basis) -> AsUpcast:
    """reference :169-176."""
    inner_recast = _basis.from_metadata(basis.metadata(), is_dual=basis.is_dual).upcast()
    recast = recast_diagonal_basis(inner_recast, basis).inner
    return AsUpcast(recast, TupleMetadata((basis.metadata(), basis.metadata())))


def momentum_operator_basis(basis: Basis) -> AsUpcast:
    """reference :200-207."""
    inner_recast = _basis.transformed_from_metadata(basis.metadata(), is_dual=basis.is_dual)
    recast = recast_diagonal_basis(inner_recast, basis).inner
    return AsUpcast(recast, TupleMetadata((basis.metadata(), basis.metadata())))
