"""Mirror of the slice of slate_quantum.noise that consumes the Hermitian-operator algebra.

Only the operator-algebra consumers are here (temperature-corrected noise operators, the Hamiltonian shift); the
noise-kernel containers and their fitting / truncation helpers are outside the hot-path scope.
"""

from slate_quantum_b200.noise import build

__all__ = ["build"]
