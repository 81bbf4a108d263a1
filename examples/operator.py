"""Counterpart of the reference's examples/operator.py: <x> of every position eigenstate is exactly i * dx; plus
the operator algebra on the same grid ([x, k] acting on a smooth state is i times the state).

    python examples/operator.py              # needs a CUDA device
"""

import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))

from slate_quantum_b200 import operator, state  # noqa: E402
from slate_quantum_b200.core.metadata import spaced_volume_metadata_from_stacked_delta_x, volume  # noqa: E402

if __name__ == "__main__":
    metadata = spaced_volume_metadata_from_stacked_delta_x((np.array([2 * np.pi]),), (500,))
    dx = volume.fundamental_stacked_dx(metadata)[0][0]
    # all 500 position eigenstates as one list: one kernel pass instead of 500 dense contractions
    states = state.StateList.from_states([state.build.position(metadata, (i,)) for i in range(metadata.fundamental_size)])
    x = operator.measure.all_x(states, axis=0).raw_data
    assert np.array_equal(x, np.arange(500) * dx)
    for i in (0, 123, 499):  # and one at a time, as the reference does
        assert operator.measure.x(state.build.position(metadata, (i,)), axis=0) == i * dx
    print("<x> of every position eigenstate is exactly i * dx")

    packet = state.build.coherent(metadata, (np.pi,), (0,), (0.3,))
    commutator = operator.commute(operator.build.x(metadata, axis=0), operator.build.k(metadata, axis=0))
    applied = operator.apply(commutator, packet).with_basis(packet.basis).raw_data
    print("max |[x, k] psi - i psi| =", np.abs(applied - 1j * packet.raw_data).max())
    print("<[x, k]> =", operator.expectation(commutator, packet))
