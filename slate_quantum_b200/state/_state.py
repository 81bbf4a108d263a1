"""State / StateList containers (reference: slate_quantum/state/_state.py:31-179)."""

from __future__ import annotations

from typing import Any

import numpy as np

from slate_quantum_b200.core import basis as _basis
from slate_quantum_b200.core.array import Array
from slate_quantum_b200.core.basis import Basis, TupleBasis


class State(Array):
    """A state vector in some basis (reference state/_state.py:31-46)."""

    def with_basis(self, basis: Basis) -> "State":
        assert basis.metadata() == self.basis.metadata()
        data = self.basis.convert_vector_into(self.device_data.reshape(-1), basis, 0)
        return State(basis, data)


def inner_product(state_0: State, state_1: State) -> complex:
    """<state_0|state_1> (reference state/_state.py:49-55)."""
    rhs = state_1.with_basis(state_0.basis).raw_data
    return complex(np.vdot(state_0.raw_data, rhs))


def normalization(state: State) -> float:
    return float(np.real(inner_product(state, state)))


def normalize(state: State) -> State:
    return State(state.basis, state.raw_data / np.sqrt(normalization(state)))


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
