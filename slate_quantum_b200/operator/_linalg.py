"""Eigen-decomposition wrappers and operator products (reference: slate_quantum/operator/_linalg.py:45-215)."""

from __future__ import annotations

import numpy as np

from slate_quantum_b200 import kernels
from slate_quantum_b200._util import with_list_basis
from slate_quantum_b200.core import basis
from slate_quantum_b200.core.basis import DiagonalBasis, FundamentalBasis, TupleBasis
from slate_quantum_b200.core.linalg import into_diagonal_hermitian as into_diagonal_hermitian_array
from slate_quantum_b200.metadata import EigenvalueMetadata
from slate_quantum_b200.core.array import _complex
from slate_quantum_b200.operator._operator import Operator, OperatorList, dense_operator
from slate_quantum_b200.state._basis import EigenstateBasis
from slate_quantum_b200.state._state import StateList


def into_diagonal_hermitian(operator: Operator) -> Operator:
    """Eigenstates of a hermitian operator (reference operator/_linalg.py:45-83).

    One batched K2 launch sequence for a BlockDiagonalBasis (Bloch) operator.  The result is
    ``Operator(DiagonalBasis(TupleBasis((eig, eig.dual))).upcast(), eigenvalues)`` with
    ``eig = EigenstateBasis(transform, direction="forward", data_id)``.
    """
    diagonal = into_diagonal_hermitian_array(operator)
    inner_basis = diagonal.basis
    explicit = inner_basis.inner.children[0]
    transform = explicit.transform()
    data_id = explicit.data_id
    inner = EigenstateBasis(transform, direction="forward", data_id=data_id).upcast()
    new_basis = DiagonalBasis(
        TupleBasis(
            (inner, EigenstateBasis(transform, direction="forward", data_id=data_id).dual_basis().upcast()),
            diagonal.basis.metadata().extra,
        )
    ).upcast()
    out = Operator(new_basis, diagonal.device_data)
    out.eigenvalues_real = diagonal.eigenvalues_real
    return out


into_diagonal = into_diagonal_hermitian


def get_eigenstates_hermitian(operator: Operator) -> StateList:
    """Dense list of eigenstates labelled by eigenvalue (reference operator/_linalg.py:86-105)."""
    diagonal = into_diagonal_hermitian(operator)
    states = diagonal.basis.inner.inner.children[0].inner.eigenvectors()
    as_tuple = with_list_basis(states, basis.from_metadata(states.basis.metadata().children[0]))
    out_basis = TupleBasis(
        (FundamentalBasis(EigenvalueMetadata(diagonal.raw_data)), basis.as_tuple(as_tuple.basis).children[1])
    ).upcast()
    return StateList(out_basis, as_tuple.raw_data.astype(np.complex128))


# ------------------------------------------------------------------------------------------------ K7 products
def _dense_pair(lhs: Operator, rhs: Operator) -> tuple[TupleBasis, object, TupleBasis, object]:
    """Both operators dense, with rhs's ket basis the dual of lhs's bra basis (the contracted index)."""
    tup_l, a = dense_operator(lhs)
    tup_r, b = dense_operator(rhs)
    if tup_r.children[0] != tup_l.children[1].dual_basis():
        tup_r, b = dense_operator(rhs.with_basis(TupleBasis((tup_l.children[1].dual_basis(), tup_r.children[1]))))
    return tup_l, a, tup_r, b


def matmul(lhs: Operator, rhs: Operator) -> Operator:
    """A_ik B_kj = M_ij (reference operator/_linalg.py:108-119) - one complex GEMM on the FP64 tensor cores."""
    tup_l, a, tup_r, b = _dense_pair(lhs, rhs)
    out = kernels.zgemm(a, b)
    return Operator(TupleBasis((tup_l.children[0], tup_r.children[1])), out.reshape(-1))


def commute(lhs: Operator, rhs: Operator) -> Operator:
    """lhs rhs - rhs lhs (reference operator/_linalg.py:122-137); the second product accumulates into the first."""
    tup, a = dense_operator(lhs)
    square = TupleBasis((tup.children[0], tup.children[0].dual_basis()))
    if tup.children[1] != square.children[1]:
        tup, a = dense_operator(lhs.with_basis(square))
    _, b = dense_operator(rhs if basis._strip(rhs.basis) == square else rhs.with_basis(square))  # noqa: SLF001
    out = kernels.zgemm(a, b)
    kernels.zgemm(b, a, alpha=-1.0, beta=1.0, out=out)
    return Operator(square, out.reshape(-1))


def dagger(operator: Operator) -> Operator:
    """Hermitian conjugate (reference operator/_linalg.py:140-147): a stride swap + conjugation, no arithmetic."""
    tup, a = dense_operator(operator)
    out_basis = TupleBasis((tup.children[1].dual_basis(), tup.children[0].dual_basis()))
    return Operator(out_basis, a.mH.resolve_conj().contiguous().reshape(-1))


def _dense_list(operators: OperatorList) -> tuple[object, TupleBasis, object]:
    """(list basis, fundamental operator basis, [L, n, n] device data) of an operator list."""
    list_basis = basis.as_tuple(operators.basis).children[0]
    meta = operators.basis.metadata().children[1]
    op_basis = basis.from_metadata(meta, is_dual=(False, True))
    dense = operators.with_operator_basis(op_basis)
    n0, n1 = basis.as_tuple(op_basis).shape
    return list_basis, basis.as_tuple(op_basis), _complex(dense.device_data).reshape(-1, n0, n1).contiguous()


def _as_fundamental(operator: Operator, op_basis: TupleBasis) -> object:
    return _complex(operator.with_basis(op_basis).device_data).reshape(op_basis.shape).contiguous()


def _wrap_list(list_basis: object, op_basis: TupleBasis, data: object, like: OperatorList) -> OperatorList:
    return OperatorList(basis.AsUpcast(TupleBasis((list_basis, op_basis)), like.basis.metadata()), data.reshape(-1))


def dagger_each(operators: OperatorList) -> OperatorList:
    """Hermitian conjugate of every operator of a list (reference operator/_linalg.py:150-177)."""
    list_basis, op_basis, data = _dense_list(operators)
    return _wrap_list(list_basis, op_basis, data.mH.resolve_conj().contiguous(), operators)


def matmul_list_operator(lhs: OperatorList, rhs: Operator) -> OperatorList:
    """(A_m)_ik B_kj for every m (reference operator/_linalg.py:172-183): one batched GEMM, B shared (batch stride 0)."""
    list_basis, op_basis, a = _dense_list(lhs)
    return _wrap_list(list_basis, op_basis, kernels.zgemm(a, _as_fundamental(rhs, op_basis)), lhs)


def matmul_operator_list(lhs: Operator, rhs: OperatorList) -> OperatorList:
    """A_ik (B_m)_kj for every m (reference operator/_linalg.py:186-196)."""
    list_basis, op_basis, b = _dense_list(rhs)
    return _wrap_list(list_basis, op_basis, kernels.zgemm(_as_fundamental(lhs, op_basis), b), rhs)


def get_commutator_operator_list(lhs: Operator, rhs: OperatorList) -> OperatorList:
    """[lhs, rhs_m] for every m (reference operator/_linalg.py:199-215; the commutators of
    noise.build.temperature_corrected_operators, noise/build/_build.py:141-155)."""
    list_basis, op_basis, b = _dense_list(rhs)
    a = _as_fundamental(lhs, op_basis)
    out = kernels.zgemm(a, b)
    kernels.zgemm(b, a, alpha=-1.0, beta=1.0, out=out)
    return _wrap_list(list_basis, op_basis, out, rhs)
