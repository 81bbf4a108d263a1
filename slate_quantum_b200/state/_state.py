"""State / StateList containers (reference: slate_quantum/state/_state.py:31-179)."""

from __future__ import annotations

from typing import Any

import numpy as np

from slate_quantum_b200 import kernels
from slate_quantum_b200.core import basis as _basis
from slate_quantum_b200.core.array import Array, _complex
from slate_quantum_b200.core.basis import BasisStateMetadata, Basis, FundamentalBasis, TupleBasis


class State(Array):
    """A state vector in some basis (reference state/_state.py:31-46)."""

    def with_basis(self, basis: Basis) -> "State":
        assert basis.metadata() == self.basis.metadata()
        data = self.basis.convert_vector_into(self.device_data.reshape(-1), basis, 0)
        return State(basis, data)


def inner_product(state_0: State, state_1: State) -> complex:
    """<state_0|state_1> (reference state/_state.py:49-55)."""
    rhs = _complex(state_1.with_basis(state_0.basis).device_data).reshape(1, -1)
    lhs = _complex(state_0.device_data).reshape(1, -1)
    return complex(kernels.rowwise_vdot(lhs, rhs).cpu().numpy()[0])


def normalization(state: State) -> float:
    """abs(<state|state>) (reference state/_state.py:58-62)."""
    return float(abs(inner_product(state, state)))


def normalize(state: State) -> State:
    """state / sqrt(<state|state>) (reference state/_state.py:65-68)."""
    data = _complex(state.device_data).reshape(1, -1).contiguous()
    return State(state.basis, kernels.normalize_rows(data).reshape(-1))


def get_occupations(state: State) -> Array:
    """abs(a_n)^2 of the state's own basis coefficients (reference state/_state.py:71-87)."""
    basis = FundamentalBasis(BasisStateMetadata(state.basis))
    return Array(basis, kernels.abs2(_complex(state.device_data).reshape(-1).contiguous()))


class StateList(Array):
    """A list of states: basis = TupleBasis((list basis, state basis)) (reference state/_state.py:90-179)."""

    def __init__(self, basis: Basis, data: Any) -> None:
        assert len(basis.metadata().children) == 2  # noqa: PLR2004
        super().__init__(basis, data)

    def with_basis(self, basis: Basis) -> "StateList":
        assert basis.metadata() == self.basis.metadata()
        data = self.basis.convert_vector_into(self.device_data.reshape(-1), basis, 0)
        return StateList(basis, data)

    def with_state_basis(self, basis: Basis) -> "StateList":
        """reference state/_state.py:153-165."""
        lhs_basis = _basis.as_tuple(self.basis).children[0]
        return self.with_basis(TupleBasis((lhs_basis, basis)).upcast())

    def with_list_basis(self, basis: Basis) -> "StateList":
        """reference state/_state.py:167-179."""
        rhs_basis = _basis.as_tuple(self.basis).children[1]
        return self.with_basis(TupleBasis((basis, rhs_basis)).upcast())

    def __len__(self) -> int:
        return _basis.as_tuple(self.basis).children[0].size

    def __iter__(self):
        tup = _basis.as_tuple(self.basis)
        data = self.raw_data.reshape(tup.shape)
        return (State(tup.children[1], row) for row in data)

    def __getitem__(self, index: Any) -> Any:
        tup = _basis.as_tuple(self.basis)
        data = self.raw_data.reshape(tup.shape)
        if isinstance(index, tuple) and isinstance(index[0], (int, np.integer)) and index[1] == slice(None):
            return State(tup.children[1], data[index[0]])
        if isinstance(index, (int, np.integer)):
            return State(tup.children[1], data[index])
        msg = "only list[i] / list[i, :] indexing is implemented"
        raise NotImplementedError(msg)

    @staticmethod
    def from_states(states: Any) -> "StateList":
        states = list(states)
        assert all(s.basis == states[0].basis for s in states)
        list_basis = _basis.FundamentalBasis.from_size(len(states))
        return StateList(TupleBasis((list_basis, states[0].basis)).upcast(), np.array([s.raw_data for s in states]))


def iter_list_tiles(states: "StateList", state_basis: Basis | None = None):
    """[rows, size] device tensors covering the list in `state_basis` (default: its own), in list order.

    Ordinary lists give one tensor; the lazy evolved lists of dynamics.schrodinger (anything with `tiles()`) are
    walked one time tile at a time so that per-state reductions never need the whole list in HBM."""
    converted = states if state_basis is None else states.with_state_basis(state_basis)
    if hasattr(converted, "tiles") and getattr(converted, "_device", None) is None:
        for _, tile in converted.tiles():
            yield tile
        return
    tup = _basis.as_tuple(converted.basis)
    yield _complex(converted.device_data).reshape(tup.children[0].size, -1).contiguous()


def _list_data(states: StateList) -> tuple[TupleBasis, Any]:
    """(tuple basis, [len(list), state size] device data) of a state list."""
    from slate_quantum_b200.core.array import as_tuple_basis

    as_tuple = as_tuple_basis(states)
    tup = _basis.as_tuple(as_tuple.basis)
    return tup, _complex(as_tuple.device_data).reshape(tup.shape).contiguous()


def inner_product_each(state_0: StateList, state_1: StateList) -> Array:
    """<state_0[j]|state_1[j]> for every j (reference state/_state.py:208-218)."""
    if state_0 is state_1 and hasattr(state_0, "tiles"):
        # <psi_j|psi_j> of a lazy evolved list: one pass per time tile
        import torch

        list_basis = _basis.as_tuple(state_0.basis).children[0]
        return Array(list_basis, torch.cat([kernels.rowwise_vdot(t, t) for t in iter_list_tiles(state_0)]))
    tup, lhs = _list_data(state_0)
    _, rhs = _list_data(state_1.with_basis(tup.upcast()))
    return Array(tup.children[0], kernels.rowwise_vdot(lhs, rhs))


def all_inner_product(state_0: StateList, state_1: StateList) -> Array:
    """<state_0[j]|state_1[k]> for every pair (reference state/_state.py:221-232): one complex GEMM (K7)."""
    tup0, lhs = _list_data(state_0)
    tup1, rhs = _list_data(state_1.with_state_basis(tup0.children[1]))
    return Array(TupleBasis((tup0.children[0], tup1.children[0])), kernels.zgemm(lhs.conj(), rhs.mT).reshape(-1))


def normalization_each(states: StateList) -> Array:
    """abs(<state[j]|state[j]>) (reference state/_state.py:235-246)."""
    product = inner_product_each(states, states)
    return Array(product.basis, product.device_data.abs())


def normalize_each(states: StateList) -> StateList:
    """Every state divided by sqrt(<state|state>) (reference state/_state.py:249-262)."""
    tup, data = _list_data(states)
    return StateList(tup.upcast(), kernels.normalize_rows(data).reshape(-1))


def get_all_occupations(states: StateList) -> Array:
    """abs(a_n(t))^2 for every state of the list, in the list's own state basis (reference state/_state.py:265-308)."""
    tup, data = _list_data(states)
    basis = TupleBasis((tup.children[0], FundamentalBasis(BasisStateMetadata(tup.children[1]))))
    return Array(basis, kernels.abs2(data).reshape(-1))
