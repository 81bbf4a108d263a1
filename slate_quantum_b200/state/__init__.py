"""Mirror of slate_quantum.state (reference: slate_quantum/state/_state.py, _basis.py, _build.py)."""

from slate_quantum_b200.state import build
from slate_quantum_b200.state._basis import EigenstateBasis
from slate_quantum_b200.state._state import (
    State,
    StateList,
    all_inner_product,
    get_all_occupations,
    get_occupations,
    inner_product,
    inner_product_each,
    normalization,
    normalization_each,
    normalize,
    normalize_each,
)

__all__ = [
    "EigenstateBasis",
    "State",
    "StateList",
    "all_inner_product",
    "build",
    "get_all_occupations",
    "get_occupations",
    "inner_product",
    "inner_product_each",
    "normalization",
    "normalization_each",
    "normalize",
    "normalize_each",
]
