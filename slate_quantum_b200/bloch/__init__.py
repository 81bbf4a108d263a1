"""Makes use of Bloch's theorem to efficiently calculate eigenstates (mirror of slate_quantum.bloch)."""

from slate_quantum_b200.bloch import build
from slate_quantum_b200.bloch._linalg import into_diagonal
from slate_quantum_b200.bloch._shifted_basis import BlochShiftedBasis
from slate_quantum_b200.bloch._transposed_basis import BlochTransposedBasis
from slate_quantum_b200.core.basis import BlockDiagonalBasis

__all__ = ["BlochShiftedBasis", "BlochTransposedBasis", "BlockDiagonalBasis", "build", "into_diagonal"]
