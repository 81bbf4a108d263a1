"""Counterpart of the reference's examples/schrodinger.py: eigen-decomposition Schrodinger evolution of a
wavepacket in a periodic potential, through the Bloch path (no plotting).

    python examples/schrodinger.py           # needs a CUDA device
"""

import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))

from slate_quantum_b200 import bloch, operator, state  # noqa: E402
from slate_quantum_b200.configs import HBAR as hbar  # noqa: E402
from slate_quantum_b200.core import FundamentalBasis, basis  # noqa: E402
from slate_quantum_b200.core.metadata import Domain, spaced_volume_metadata_from_stacked_delta_x  # noqa: E402
from slate_quantum_b200.dynamics import solve_schrodinger_equation_decomposition  # noqa: E402
from slate_quantum_b200.metadata import SpacedTimeMetadata, repeat_volume_metadata  # noqa: E402

if __name__ == "__main__":
    # unit cell: 2-D square lattice, 16 x 16 points; supercell = 16 x 16 unit cells
    metadata = spaced_volume_metadata_from_stacked_delta_x(
        (np.array([2 * np.pi, 0]), np.array([0, 2 * np.pi])), (16, 16)
    )
    repeat = (16, 16)
    potential = operator.build.cos_potential(metadata, 10.0)
    hamiltonian = bloch.build.kinetic_hamiltonian(potential, hbar**2, repeat)

    supercell = repeat_volume_metadata(metadata, repeat)
    initial_state = state.build.coherent(
        supercell, (16 * np.pi, 16 * np.pi), (0.0, 0.0), (2 * np.pi, 2 * np.pi)
    )
    times = FundamentalBasis(SpacedTimeMetadata(60, domain=Domain(delta=8 * np.pi * hbar)))
    evolution = solve_schrodinger_equation_decomposition(initial_state, times, hamiltonian)

    # the reference returns amplitudes in the eigenbasis: |a_n(t)| does not change with time
    amplitudes = evolution.raw_data
    print("eigenbasis amplitudes:", amplitudes.shape, "max | |a(t)| - |a(0)| | =",
          np.abs(np.abs(amplitudes) - np.abs(amplitudes[0])).max())

    # back in the position basis (fused phase x eigenvector GEMM + Bloch permutation + FFT on the GPU)
    position = evolution.with_state_basis(basis.from_metadata(supercell).upcast()).raw_data.reshape(60, 256, 256)
    density = np.abs(position) ** 2
    print("norm over time:", density.sum(axis=(1, 2))[[0, 29, 59]])
    x = np.arange(256)
    print("<x> drift (grid points):", (density.sum(axis=2) @ x)[[0, 29, 59]])
