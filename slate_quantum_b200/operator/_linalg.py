"""Eigen-decomposition wrappers (reference: slate_quantum/operator/_linalg.py:45-105)."""

from __future__ import annotations

import numpy as np

from slate_quantum_b200._util import with_list_basis
from slate_quantum_b200.core import basis
from slate_quantum_b200.core.basis import DiagonalBasis, FundamentalBasis, TupleBasis
from slate_quantum_b200.core.linalg import into_diagonal_hermitian as into_diagonal_hermitian_array
from slate_quantum_b200.metadata import EigenvalueMetadata
from slate_quantum_b200.operator._operator import Operator
from slate_quantum_b200.state._basis import EigenstateBasis
from slate_quantum_b200.state._state import StateList


def into_diagonal_hermitian(operator: Operator) -> Operator:
    """Eigenstates of a hermitian operator (reference operator/_linalg.py:45-83).

    One batched K2 launch sequence for a BlockDiagonalBasis (Bloch) operator.  The result is
    ``Operator(DiagonalBasis(TupleBasis((eig, eig.dual))).upcast(), eigenvalues)`` with
    ``eig = EigenstateBasis(transform, direction="forward", data_id)``.
    """
    diagonal = into_diagonal_hermitian_array(operator)
    inner_basis = diagonal.basis
    explicit = inner_basis.inner.children[0]
    transform = explicit.transform()
    data_id = explicit.data_id
    inner = EigenstateBasis(transform, direction="forward", data_id=data_id).upcast()
    new_basis = DiagonalBasis(
        TupleBasis(
            (inner, EigenstateBasis(transform, direction="forward", data_id=data_id).dual_basis().upcast()),
            diagonal.basis.metadata().extra,
        )
    ).upcast()
    out = Operator(new_basis, diagonal.device_data)
    out.eigenvalues_real = diagonal.eigenvalues_real
    return out


into_diagonal = into_diagonal_hermitian


def get_eigenstates_hermitian(operator: Operator) -> StateList:
    """Dense list of eigenstates labelled by eigenvalue (reference operator/_linalg.py:86-105)."""
    diagonal = into_diagonal_hermitian(operator)
    states = diagonal.basis.inner.inner.children[0].inner.eigenvectors()
    as_tuple = with_list_basis(states, basis.from_metadata(states.basis.metadata().children[0]))
    out_basis = TupleBasis(
        (FundamentalBasis(EigenvalueMetadata(diagonal.raw_data)), basis.as_tuple(as_tuple.basis).children[1])
    ).upcast()
    return StateList(out_basis, as_tuple.raw_data.astype(np.complex128))
