"""Mirror of slate_quantum.operator for the hot path (reference: slate_quantum/operator/__init__.py)."""

from slate_quantum_b200.operator import build, measure
from slate_quantum_b200.operator._diagonal import (
    momentum_operator_basis,
    position_operator_basis,
    recast_diagonal_basis,
    with_outer_basis,
)
from slate_quantum_b200.operator._linalg import (
    commute,
    dagger,
    dagger_each,
    get_commutator_operator_list,
    get_eigenstates_hermitian,
    into_diagonal,
    into_diagonal_hermitian,
    matmul,
    matmul_list_operator,
    matmul_operator_list,
)
from slate_quantum_b200.operator._operator import (
    Operator,
    OperatorList,
    apply,
    apply_to_each,
    expectation,
    expectation_of_each,
    operator_basis,
)

__all__ = [
    "Operator",
    "OperatorList",
    "apply",
    "apply_to_each",
    "build",
    "commute",
    "dagger",
    "dagger_each",
    "expectation",
    "expectation_of_each",
    "get_commutator_operator_list",
    "get_eigenstates_hermitian",
    "into_diagonal",
    "into_diagonal_hermitian",
    "matmul",
    "matmul_list_operator",
    "matmul_operator_list",
    "measure",
    "momentum_operator_basis",
    "operator_basis",
    "position_operator_basis",
    "recast_diagonal_basis",
    "with_outer_basis",
]
