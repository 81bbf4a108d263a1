"""reference: slate_quantum/bloch/_linalg.py:72-84."""

from __future__ import annotations

from slate_quantum_b200.operator._linalg import into_diagonal_hermitian
from slate_quantum_b200.operator._operator import Operator


def into_diagonal(operator: Operator) -> Operator:
    """Diagonalise a Bloch (block diagonal) operator block by block: one batched K2 call."""
    return into_diagonal_hermitian(operator)
