"""Noise kernels and their eigenvalue diagonalisation (reference: slate_quantum/noise/_kernel.py:38-177,
noise/diagonalize/_eigenvalue.py:22-131) - the second consumer of the Hermitian eigensolver.

A noise kernel beta is the correlation of a set of noise operators L_a with weights lambda_a:

    full kernel      beta[(i j), (k l)] = sum_a lambda_a conj(L_a[j, i]) L_a[k, l]        (reference :38-66)
    diagonal kernel  beta[i, j]         = sum_a lambda_a conj(d_a[i]) d_a[j],  L_a = diag(d_a)   (reference :145-177)

Both contractions are ONE complex GEMM on the FP64 tensor cores (K7) with the list index as the inner dimension;
the reference's `np.linalg.eigh` call sites (noise/diagonalize/_eigenvalue.py:53 and :103) run on the batched
eigensolver (K2; matrices above 512 rows take the block-Jacobi path of core/jacobi.py).  Only the containers these
two functions need are mirrored: the kernel is an Array over (operator basis, operator basis), or over the
DiagonalBasis pair for a diagonal kernel.
"""

from __future__ import annotations

import numpy as np
import torch

from slate_quantum_b200 import kernels
from slate_quantum_b200.core import basis as _basis
from slate_quantum_b200.core.array import Array, _complex
from slate_quantum_b200.core.basis import Basis, DiagonalBasis, FundamentalBasis, TupleBasis
from slate_quantum_b200.metadata import EigenvalueMetadata
from slate_quantum_b200.operator._operator import OperatorList


class NoiseKernel(Array):
    """beta[(i j), (k l)] over TupleBasis((operator basis, operator basis)) (reference noise/_kernel.py:32-35)."""


class DiagonalNoiseKernel(Array):
    """beta[i, j] of noise operators that are all diagonal in `state_basis` (reference noise/_kernel.py:89-108)."""

    def __init__(self, basis: Basis, data: object) -> None:
        super().__init__(basis, data)

    @property
    def state_basis(self) -> Basis:
        return _basis.as_tuple(self.basis).children[0]


def _list_values(operators: OperatorList) -> tuple[torch.Tensor, Basis, Basis]:
    tup = _basis.as_tuple(operators.basis)
    values = np.asarray(tup.children[0].metadata().values, dtype=np.complex128)
    return kernels.to_device(values, torch.complex128), tup.children[0], tup.children[1]


def noise_kernel_from_operators(operators: OperatorList) -> NoiseKernel:
    """reference noise/_kernel.py:38-66: einsum("a,aji,akl->ij kl", lambda, conj(L), L) as one GEMM."""
    values, _, op_basis = _list_values(operators)
    meta = operators.basis.metadata().children[1]
    dense_basis = _basis.from_metadata(meta, is_dual=(False, True))
    data = _complex(operators.with_operator_basis(dense_basis).device_data)
    n = _basis.as_tuple(dense_basis).shape[0]
    ops = data.reshape(-1, n, n)
    lhs = (ops.mT.contiguous().reshape(-1, n * n) * values[:, None]).contiguous()  # [a, (i j)] = lambda_a L_a[j, i]
    rhs = ops.reshape(-1, n * n).contiguous()                                       # [a, (k l)]
    out = kernels.zgemm(torch.conj_physical(lhs).mT, rhs)                           # [(i j), (k l)]
    del op_basis
    return NoiseKernel(TupleBasis((dense_basis, dense_basis)), out.reshape(-1))


def build_diagonal_kernel(state_basis: Basis, data: object) -> DiagonalNoiseKernel:
    """reference noise/_kernel.py:110-126 (the recast bookkeeping reduced to the basis of the diagonal index)."""
    return DiagonalNoiseKernel(TupleBasis((state_basis, state_basis.dual_basis())), data)


def diagonal_kernel_from_operators(operators: OperatorList) -> DiagonalNoiseKernel:
    """reference noise/_kernel.py:145-177: einsum("a,ai,aj->ij", lambda, conj(d), d) as one GEMM."""
    values, _, _ = _list_values(operators)
    meta = operators.basis.metadata().children[1]
    state_basis = _basis.from_metadata(meta.children[0])
    diag_basis = DiagonalBasis(TupleBasis((state_basis, state_basis.dual_basis())))
    d = _complex(operators.with_operator_basis(diag_basis).device_data).reshape(values.numel(), -1).contiguous()
    lhs = torch.conj_physical(d * values[:, None]).contiguous()  # lambda real in every use; conj(lambda d) = lambda conj(d)
    out = kernels.zgemm(lhs.mT, d)
    return build_diagonal_kernel(state_basis, out.reshape(-1))


def _eigh_dense(data: torch.Tensor) -> tuple[torch.Tensor, torch.Tensor]:
    n = data.shape[0]
    herm = (0.5 * (data + data.mH)).contiguous()
    scale = float(herm.abs().max().item()) or 1.0
    if float((data - data.mH).abs().max().item()) > 1e-6 * scale:
        msg = "kernel non hermitian"  # the reference asserts the same (noise/diagonalize/_eigenvalue.py:51, :98-100)
        raise AssertionError(msg)
    w, t, info = kernels.eigh_batched(herm.reshape(1, n, n))
    if int(info.abs().sum().item()):
        msg = "noise kernel eigensolve did not converge"
        raise RuntimeError(msg)
    return w[0], t[0]  # rows of t = eigenvectors


def get_periodic_noise_operators_diagonal_eigenvalue(kernel: DiagonalNoiseKernel) -> OperatorList:
    """N independent DIAGONAL noise operators of a diagonal kernel, labelled by the kernel's eigenvalues
    (reference noise/diagonalize/_eigenvalue.py:80-131): data = conj(U^T) of eigh(beta)."""
    state_basis = kernel.state_basis
    n = state_basis.size
    w, t = _eigh_dense(_complex(kernel.device_data).reshape(n, n))
    eigenvalue = FundamentalBasis(EigenvalueMetadata(w.cpu().numpy()))
    op_basis = DiagonalBasis(TupleBasis((state_basis, state_basis.dual_basis())))
    return OperatorList(TupleBasis((eigenvalue, op_basis)).upcast(), torch.conj_physical(t).reshape(-1))


def get_periodic_noise_operators_eigenvalue(kernel: NoiseKernel) -> OperatorList:
    """The n^2 noise operators that diagonalise a full kernel (reference noise/diagonalize/_eigenvalue.py:22-77)."""
    op_basis = _basis.as_tuple(kernel.basis).children[0]
    n = _basis.as_tuple(op_basis).shape[0]
    data = _complex(kernel.device_data).reshape(n, n, n, n).transpose(0, 1).reshape(n * n, n * n).contiguous()
    w, t = _eigh_dense(data)
    eigenvalue = FundamentalBasis(EigenvalueMetadata(w.cpu().numpy()))
    return OperatorList(TupleBasis((eigenvalue, op_basis)).upcast(), torch.conj_physical(t).reshape(-1))
