"""Mirror of slate_quantum.state (reference: slate_quantum/state/_state.py, _basis.py, _build.py)."""

from slate_quantum_b200.state import build
from slate_quantum_b200.state._basis import EigenstateBasis
from slate_quantum_b200.state._state import (
    State,
    StateList,
    inner_product,
    normalization,
    normalize,
)

__all__ = ["EigenstateBasis", "State", "StateList", "build", "inner_product", "normalization", "normalize"]
