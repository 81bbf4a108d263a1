"""Mirror of slate_quantum.operator for the hot path (reference: slate_quantum/operator/__init__.py)."""

from slate_quantum_b200.operator import build, measure
from slate_quantum_b200.operator._diagonal import (
    momentum_operator_basis,
    position_operator_basis,
    recast_diagonal_basis,
    with_outer_basis,
)
from slate_quantum_b200.operator._linalg import get_eigenstates_hermitian, into_diagonal, into_diagonal_hermitian
from slate_quantum_b200.operator._operator import Operator, OperatorList, operator_basis

__all__ = [
    "Operator",
    "OperatorList",
    "build",
    "get_eigenstates_hermitian",
    "into_diagonal",
    "into_diagonal_hermitian",
    "measure",
    "momentum_operator_basis",
    "operator_basis",
    "position_operator_basis",
    "recast_diagonal_basis",
    "with_outer_basis",
]
