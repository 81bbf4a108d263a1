"""slate_core.linalg.into_diagonal_hermitian on the B200: one batched eigensolver call (K2)."""

from __future__ import annotations

import uuid

import numpy as np
import torch

from slate_quantum_b200 import kernels
from slate_quantum_b200.core import basis as _basis
from slate_quantum_b200.core.array import Array
from slate_quantum_b200.core.basis import (
    AsUpcast,
    BasisStateMetadata,
    BlockDiagonalBasis,
    DiagonalBasis,
    ExplicitUnitaryBasis,
    FundamentalBasis,
    TupleBasis,
)
from slate_quantum_b200.core.metadata import SimpleMetadata


class EighFailedError(RuntimeError):
    """The batched eigensolver reported non-convergence (info[] != 0) for some block."""


def into_diagonal_hermitian(operator: Array) -> Array:
    """Diagonalise a hermitian operator; block-diagonal operators are diagonalised block by block.

    Replaces slate_core.linalg.into_diagonal_hermitian (call site: reference operator/_linalg.py:56).
    Returns Array(DiagonalBasis(TupleBasis((explicit, explicit.dual))), eigenvalues) where
    ``explicit = ExplicitUnitaryBasis(transform)`` and ``transform`` rows are eigenvectors
    ([n_blocks, N(eigen index), N(component)], type at reference bloch/_linalg.py:35-57).
    Eigenvalues are complex128 with zero imaginary part, ascending within each block (quirk Q3).
    """
    outer = _basis._strip(operator.basis)  # noqa: SLF001
    if isinstance(outer, BlockDiagonalBasis):
        state_basis = _basis.as_tuple(outer.inner).children[0]
        n_blocks, n = outer.n_blocks, outer.block_shape[0]
        assert outer.block_shape[0] == outer.block_shape[1]
        blocks = operator.device_data.reshape(n_blocks, n, n).clone()
    else:
        tup = _basis.as_tuple(_basis.from_metadata(operator.basis.metadata()))
        state_basis = _basis.transformed_from_metadata(operator.basis.metadata().children[0])
        dense_basis = TupleBasis((state_basis, state_basis.dual_basis()))
        del tup
        n_blocks, n = 1, state_basis.size
        blocks = operator.with_basis(dense_basis).device_data.reshape(1, n, n).clone()
    w, v, info = kernels.eigh_batched(blocks)
    bad = int((info != 0).sum().item())
    if bad:
        msg = f"eigensolver failed to converge on {bad} of {n_blocks} blocks"
        raise EighFailedError(msg)
    eig_index = FundamentalBasis(SimpleMetadata(n_blocks * n))
    states = FundamentalBasis(BasisStateMetadata(state_basis))
    tbasis = TupleBasis((eig_index, states))
    transform_basis = BlockDiagonalBasis(tbasis, (n, n)).upcast() if n_blocks > 1 else tbasis.upcast()
    transform = Array(transform_basis, v.reshape(-1))
    explicit = ExplicitUnitaryBasis(transform, direction="forward", data_id=uuid.uuid4())
    out_basis = DiagonalBasis(TupleBasis((explicit, explicit.dual_basis()), operator.basis.metadata().extra))
    eigenvalues = torch.complex(w.reshape(-1), torch.zeros_like(w.reshape(-1)))
    out = Array(out_basis, eigenvalues)
    out.eigenvalues_real = w  # [n_blocks, n] float64 device tensor, used by the fused evolution
    return out


def into_diagonal(operator: Array) -> Array:
    return into_diagonal_hermitian(operator)


__all__ = ["EighFailedError", "into_diagonal", "into_diagonal_hermitian", "AsUpcast", "np"]
