"""Mirror of slate_quantum._util (reference: slate_quantum/_util/__init__.py, _prod.py)."""

from __future__ import annotations

import numpy as np

from slate_quantum_b200.core import basis as _basis
from slate_quantum_b200.core.array import Array
from slate_quantum_b200.core.basis import Basis, TupleBasis


def outer_product(*arrays: np.ndarray) -> np.ndarray:
    """reference _util/_prod.py:6-10."""
    grids = np.meshgrid(*arrays, indexing="ij")
    return np.prod(grids, axis=0)


def with_list_basis(array: Array, basis: Basis) -> Array:
    """reference _util/__init__.py:24-36."""
    rhs_basis = _basis.as_tuple(array.basis).children[1]
    out_basis = TupleBasis((basis, rhs_basis)).upcast()
    return array.with_basis(out_basis)


__all__ = ["outer_product", "with_list_basis"]
