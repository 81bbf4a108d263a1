"""Mirror of the slice of slate_quantum.noise that consumes the Hermitian-operator algebra and the eigensolver
(SURVEY section 8f rank 4).

`build`: temperature-corrected noise operators, the Hamiltonian shift (K7).  `_kernel`: noise kernels from operator
lists (K7) and their eigenvalue diagonalisation (K2) - the reference's `np.linalg.eigh` call sites
(noise/diagonalize/_eigenvalue.py:53, :103).  Kernel fitting / truncation / isotropic re-parametrisations are
outside the hot-path scope.
"""

from slate_quantum_b200.noise import build
from slate_quantum_b200.noise._kernel import (
    DiagonalNoiseKernel,
    NoiseKernel,
    build_diagonal_kernel,
    diagonal_kernel_from_operators,
    get_periodic_noise_operators_diagonal_eigenvalue,
    get_periodic_noise_operators_eigenvalue,
    noise_kernel_from_operators,
)

__all__ = [
    "DiagonalNoiseKernel",
    "NoiseKernel",
    "build",
    "build_diagonal_kernel",
    "diagonal_kernel_from_operators",
    "get_periodic_noise_operators_diagonal_eigenvalue",
    "get_periodic_noise_operators_eigenvalue",
    "noise_kernel_from_operators",
]
