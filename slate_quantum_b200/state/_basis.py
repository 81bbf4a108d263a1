"""EigenstateBasis (reference: slate_quantum/state/_basis.py:25-111)."""

from __future__ import annotations

from slate_quantum_b200.core.basis import ExplicitUnitaryBasis


class EigenstateBasis(ExplicitUnitaryBasis):
    """A basis with data stored as eigenstates; `transform()` rows are the eigenvectors."""

    def eigenvectors(self):
        from slate_quantum_b200.state._state import StateList

        states = super().eigenvectors()
        return StateList(states.basis, states.device_data)
