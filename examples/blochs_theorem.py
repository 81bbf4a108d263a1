"""Counterpart of the reference's examples/blochs_theorem.py on the B200 path (no plotting).

    python examples/blochs_theorem.py        # needs a CUDA device

A periodic potential's Hamiltonian is block diagonal in the Bloch basis: `bloch.build.kinetic_hamiltonian`
writes every k-block with one kernel launch and `get_eigenstates_hermitian` / `bloch.into_diagonal`
diagonalise all blocks with one batched eigensolver call.
"""

import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))

from slate_quantum_b200 import bloch, operator  # noqa: E402
from slate_quantum_b200.configs import HBAR as hbar  # noqa: E402
from slate_quantum_b200.core import basis  # noqa: E402
from slate_quantum_b200.core.metadata import spaced_volume_metadata_from_stacked_delta_x  # noqa: E402

if __name__ == "__main__":
    metadata = spaced_volume_metadata_from_stacked_delta_x((np.array([2 * np.pi]),), (60,))
    rng = np.random.default_rng(0)
    potential = operator.build.potential_from_function(
        metadata, lambda x: 400 * rng.normal(size=x[0].size).astype(np.complex128)
    )
    hamiltonian = bloch.build.kinetic_hamiltonian(potential, hbar**2, (3,))

    # The hamiltonian is block diagonal in the BlochTransposedBasis: raw data = [n_k, N, N] blocks
    blocks = hamiltonian.raw_data.reshape(3, 60, 60)
    print("blocks:", blocks.shape, "hermitian:", np.allclose(blocks, blocks.conj().transpose(0, 2, 1)))

    # into_diagonal is aware that the matrix is block diagonal: one batched eigensolver call
    eigenstates = operator.get_eigenstates_hermitian(hamiltonian)
    energies = eigenstates.basis.metadata().children[0].values.real
    print("lowest energies:", np.sort(energies)[:5])

    # states in the supercell momentum basis
    momentum = basis.transformed_from_metadata(hamiltonian.basis.metadata().children[0]).upcast()
    in_k = eigenstates.with_state_basis(momentum).raw_data.reshape(180, 180)
    print("norms of the first three states:", np.linalg.norm(in_k[:3], axis=1))
