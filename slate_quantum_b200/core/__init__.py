"""A minimal, CUDA-backed stand-in for the `slate_core` surface that slate_quantum's hot path touches.

Only what SURVEY.md section 8(b) lists is provided; names and call signatures follow slate_core so that the
mirrored `slate_quantum_b200.{bloch,operator,state,dynamics,metadata}` modules read like the reference.
"""

from slate_quantum_b200.core import array, basis, linalg, metadata
from slate_quantum_b200.core.array import Array
from slate_quantum_b200.core.basis import (
    AsUpcast,
    Basis,
    Ctype,
    FundamentalBasis,
    TransformedBasis,
    TupleBasis,
)
from slate_quantum_b200.core.metadata import BasisMetadata, SimpleMetadata, TupleMetadata

__all__ = [
    "Array",
    "AsUpcast",
    "Basis",
    "BasisMetadata",
    "Ctype",
    "FundamentalBasis",
    "SimpleMetadata",
    "TransformedBasis",
    "TupleBasis",
    "TupleMetadata",
    "array",
    "basis",
    "linalg",
    "metadata",
]
