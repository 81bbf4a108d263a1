"""Counterpart of the reference's examples/eigenstates.py (no plotting): eigenstates of a free particle and of a
particle in a cosine well on ONE cell - the dense, non-Bloch use of the eigensolver (batch of one matrix).

    python examples/eigenstates.py           # needs a CUDA device
"""

import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))

from slate_quantum_b200 import operator, state  # noqa: E402
from slate_quantum_b200.configs import HBAR as hbar  # noqa: E402
from slate_quantum_b200.core import basis  # noqa: E402
from slate_quantum_b200.core.metadata import spaced_volume_metadata_from_stacked_delta_x  # noqa: E402

if __name__ == "__main__":
    # Metadata for a 1D volume with 60 points in the x direction
    metadata = spaced_volume_metadata_from_stacked_delta_x((np.array([2 * np.pi]),), (60,))
    position_basis = basis.from_metadata(metadata).upcast()
    momentum_basis = basis.transformed_from_metadata(metadata).upcast()

    potential = operator.build.cos_potential(metadata, 0)
    hamiltonian = operator.build.kinetic_hamiltonian(potential, hbar**2)
    eigenstates = operator.get_eigenstates_hermitian(hamiltonian)
    # The eigenstates of a free particle are plane waves: flat in position space, <= 2 components in momentum space
    for i, s in enumerate(list(eigenstates)[:3]):
        in_x = np.abs(s.with_basis(position_basis).raw_data)
        in_k = np.abs(s.with_basis(momentum_basis).raw_data)
        print(f"free particle state {i}: max|psi(x)| / min|psi(x)| over occupied = "
              f"{in_x.max():.4f} / {in_x.min():.4f}; momentum components > 1e-8: {(in_k > 1e-8).sum()}")

    # Now we modify the potential to have a barrier: the eigenstates localise around the potential minimum
    potential = operator.build.cos_potential(metadata, 10)
    hamiltonian = operator.build.kinetic_hamiltonian(potential, hbar**2)
    eigenstates = operator.get_eigenstates_hermitian(hamiltonian)
    energies = eigenstates.basis.metadata().children[0].values
    print("lowest energies in the well:", np.array2string(np.asarray(energies[:4]).real, precision=6))
    for i, s in enumerate(list(eigenstates)[:3]):
        print(f"well state {i}: <x> = {operator.measure.periodic_x(s, axis=0):.4f}, "
              f"width = {operator.measure.coherent_width(s, axis=0):.4f}, "
              f"<H> = {operator.expectation(hamiltonian, s).real:.6f}, norm = {state.normalization(s):.12f}")
