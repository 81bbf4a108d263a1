"""Operator / OperatorList containers (reference: slate_quantum/operator/_operator.py:58-395)."""

from __future__ import annotations

from typing import Any, Iterable

import numpy as np

from slate_quantum_b200.core import basis as _basis
from slate_quantum_b200.core.array import Array
from slate_quantum_b200.core.basis import AsUpcast, Basis, FundamentalBasis, TupleBasis, are_dual_shapes


def _assert_operator_basis(basis: Basis) -> None:
    """reference operator/_operator.py:34-40."""
    is_dual = basis.is_dual
    if isinstance(is_dual, bool) or len(is_dual) != 2:  # noqa: PLR2004
        msg = "Basis is not 2d"
        raise TypeError(msg)
    assert are_dual_shapes(is_dual[0], is_dual[1])


def operator_basis(basis: Basis) -> TupleBasis:
    return TupleBasis((basis, basis.dual_basis()))


class Operator(Array):
    """An operator: basis metadata = TupleMetadata((M, M)) (reference :58-75)."""

    def __init__(self, basis: Basis, data: Any) -> None:
        _assert_operator_basis(basis)
        super().__init__(basis, data)

    def with_basis(self, basis: Basis) -> "Operator":
        assert basis.metadata() == self.basis.metadata()
        data = self.basis.convert_vector_into(self.device_data.to(dtype=_c128()).reshape(-1), basis, 0)
        return Operator(basis, data)

    def as_array(self) -> np.ndarray:
        n = self.basis.metadata().children[0].fundamental_size
        fund = self.basis.into_fundamental(self.device_data.to(dtype=_c128()).reshape(-1), 0)
        return fund.detach().cpu().numpy().reshape(n, n)


def _c128():
    import torch

    return torch.complex128


def _assert_operator_list_basis(basis: Basis) -> None:
    is_dual = basis.is_dual
    if isinstance(is_dual, bool) or len(is_dual) != 2:  # noqa: PLR2004
        msg = "Basis is not 2d"
        raise TypeError(msg)


class OperatorList(Array):
    """A list of operators: basis = TupleBasis((list basis, operator basis)) (reference :239-395)."""

    def __init__(self, basis: Basis, data: Any) -> None:
        _assert_operator_list_basis(basis)
        super().__init__(basis, data)

    def with_basis(self, basis: Basis) -> "OperatorList":
        assert basis.metadata() == self.basis.metadata()
        data = self.basis.convert_vector_into(self.device_data.reshape(-1), basis, 0)
        return OperatorList(basis, data)

    def with_operator_basis(self, basis: Basis) -> "OperatorList":
        """reference :268-284."""
        lhs_basis = _basis.as_tuple(self.basis).children[0]
        return self.with_basis(AsUpcast(TupleBasis((lhs_basis, basis)), self.basis.metadata()))

    def with_list_basis(self, basis: Basis) -> "OperatorList":
        """reference :315-327."""
        rhs_basis = _basis.as_tuple(self.basis).children[1]
        return self.with_basis(AsUpcast(TupleBasis((basis, rhs_basis)), self.basis.metadata()))

    def __iter__(self):
        tup = _basis.as_tuple(self.basis)
        data = self.raw_data.reshape(tup.shape)
        return (Operator(tup.children[1], row) for row in data)

    def __len__(self) -> int:
        return _basis.as_tuple(self.basis).children[0].size

    def __getitem__(self, index: Any) -> Operator:
        """list[i] / list[i, :] -> the i-th operator (reference :329-375, integer index only)."""
        tup = _basis.as_tuple(self.basis)
        if isinstance(index, tuple) and len(index) == 2 and index[1] == slice(None):  # noqa: PLR2004
            index = index[0]
        if not isinstance(index, (int, np.integer)):
            msg = "only list[i] / list[i, :] indexing is implemented"
            raise NotImplementedError(msg)
        return Operator(tup.children[1], self.raw_data.reshape(tup.shape)[index])

    @staticmethod
    def from_operators(iter_: Iterable[Operator]) -> "OperatorList":
        """reference :377-395."""
        ops = list(iter_)
        assert all(x.basis == ops[0].basis for x in ops)
        list_basis = FundamentalBasis.from_size(len(ops))
        return OperatorList(TupleBasis((list_basis, ops[0].basis)).upcast(), np.array([x.raw_data for x in ops]))


# ------------------------------------------------------------------------------------------------ K7 contractions
def dense_operator(operator: Operator):
    """(TupleBasis((ket basis, bra basis)), [n0, n1] device matrix).  An operator that is not stored over a tuple
    basis (diagonal, recast, block-diagonal ...) is expanded into the dense fundamental operator basis."""
    from slate_quantum_b200.core.array import _complex

    tup = _basis._strip(operator.basis)  # noqa: SLF001
    if not isinstance(tup, TupleBasis):
        tup = _basis.from_metadata(operator.basis.metadata(), is_dual=(False, True))
        operator = operator.with_basis(tup)
    return tup, _complex(operator.device_data).reshape(tup.shape).contiguous()


_dense_operator = dense_operator


def apply(operator: Operator, state: Any) -> Any:
    """O|psi> in the operator's ket basis (reference operator/_operator.py:210-218)."""
    from slate_quantum_b200 import kernels
    from slate_quantum_b200.core.array import _complex
    from slate_quantum_b200.state._state import State

    tup, o = _dense_operator(operator)
    psi = _complex(state.with_basis(tup.children[1].dual_basis()).device_data).reshape(-1, 1)
    return State(tup.children[0], kernels.zgemm(o, psi).reshape(-1))


def expectation(operator: Operator, state: Any) -> complex:
    """<psi|O|psi> (reference operator/_operator.py:178-193)."""
    from slate_quantum_b200 import kernels
    from slate_quantum_b200.core.array import _complex

    applied = apply(operator, state)
    lhs = _complex(state.with_basis(applied.basis).device_data).reshape(1, -1)
    return complex(kernels.rowwise_vdot(lhs, applied.device_data.reshape(1, -1)).cpu().numpy()[0])


def _diagonal_form(operator: Operator):
    """(ket basis, diagonal) when the operator is stored as a diagonal over (b, b.dual) - potentials, x
    (position basis), kinetic energy, k, p (momentum basis), the output of into_diagonal_hermitian (eigenbasis) -
    else None."""
    from slate_quantum_b200.core.array import _complex
    from slate_quantum_b200.core.basis import DiagonalBasis, RecastBasis

    stored = _basis._strip(operator.basis)  # noqa: SLF001
    data = _complex(operator.device_data).reshape(-1)
    if isinstance(stored, RecastBasis):
        inner = _basis._strip(stored.inner)  # noqa: SLF001
        if not isinstance(inner, DiagonalBasis):
            return None
        data = stored.into_inner(data, 0)
        stored = inner
    if not isinstance(stored, DiagonalBasis):
        return None
    ket, bra = _basis.as_tuple(stored.inner).children
    if bra != ket.dual_basis():
        return None
    return ket, data.contiguous()


def expectation_of_each(operator: Operator, states: Any) -> Array:
    """<psi_a|O|psi_a> for every state of a list (reference operator/_operator.py:196-207).

    Diagonal operators (potential, x: position basis; kinetic energy, k, p: momentum basis; a diagonalised
    Hamiltonian: its eigenbasis) take ONE pass over the states written in that basis, sum_i d_i |psi_ai|^2
    (sqb_rowwise_weighted_abs2) - the only form that scales to the Bloch supercells, where a dense [M, M] operator
    cannot exist; a SplitBasis Hamiltonian (potential + kinetic) is the sum of its two diagonal parts.  Anything
    else: Phi = Psi O^T as one complex GEMM over the whole list (the operator is read through swapped strides),
    then a row-wise <Psi|Phi>."""
    from slate_quantum_b200 import kernels
    from slate_quantum_b200.core.array import _complex
    from slate_quantum_b200.core.basis import SplitBasis

    list_basis = _basis.as_tuple(states.basis).children[0]
    stored = _basis._strip(operator.basis)  # noqa: SLF001
    if isinstance(stored, SplitBasis):
        data = operator.device_data.reshape(-1)
        lhs = expectation_of_each(Operator(stored.lhs, data[: stored.lhs.size]), states)
        rhs = expectation_of_each(Operator(stored.rhs, data[stored.lhs.size :]), states)
        return Array(list_basis, lhs.device_data + rhs.device_data)
    diagonal = _diagonal_form(operator)
    if diagonal is not None:
        import torch

        from slate_quantum_b200.state._state import iter_list_tiles

        ket, diag = diagonal
        parts = [kernels.rowwise_weighted_abs2(psi, diag) for psi in iter_list_tiles(states, ket)]
        return Array(list_basis, parts[0] if len(parts) == 1 else torch.cat(parts))
    tup, o = _dense_operator(operator)
    rhs = states.with_state_basis(tup.children[1].dual_basis())
    psi1 = _complex(rhs.device_data).reshape(list_basis.size, -1)
    phi = kernels.zgemm(psi1, o.mT)
    lhs = states.with_state_basis(tup.children[0])
    psi0 = _complex(lhs.device_data).reshape(list_basis.size, -1)
    return Array(list_basis, kernels.rowwise_vdot(psi0, phi))


def apply_to_each(operator: Operator, states: Any) -> Any:
    """O|psi_a> for every state of a list, in the operator's ket basis (reference operator/_operator.py:219-226):
    a column scaling for operators stored as diagonals, else Psi O^T as one complex GEMM."""
    from slate_quantum_b200 import kernels
    from slate_quantum_b200.core.array import _complex
    from slate_quantum_b200.state._state import StateList

    list_basis = _basis.as_tuple(states.basis).children[0]
    diagonal = _diagonal_form(operator)
    if diagonal is not None:
        ket, diag = diagonal
        psi = _complex(states.with_state_basis(ket).device_data).reshape(list_basis.size, -1)
        return StateList(TupleBasis((list_basis, ket)).upcast(), kernels.scale_columns(psi, diag).reshape(-1))
    tup, o = dense_operator(operator)
    psi = _complex(states.with_state_basis(tup.children[1].dual_basis()).device_data).reshape(list_basis.size, -1)
    return StateList(TupleBasis((list_basis, tup.children[0])).upcast(), kernels.zgemm(psi, o.mT).reshape(-1))
