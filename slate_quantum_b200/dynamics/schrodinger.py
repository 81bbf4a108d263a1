"""Schrodinger evolution by eigen-decomposition (reference: slate_quantum/dynamics/schrodinger.py:31-73).

The reference materialises ``coefficients[None, :] * np.exp(-1j * E * t / hbar)`` ([T, M] complex128) and
leaves the back-transformation to the caller's ``with_state_basis``.  Here the returned StateList is LAZY:
it holds (c, E, t) on the GPU; ``raw_data`` materialises the eigenbasis amplitudes with one kernel
(sqb_evolve_eigen) and ``with_state_basis`` to the Bloch / momentum / position basis of a Bloch Hamiltonian
runs the fused phase x eigenvector DMMA GEMM (sqb_evolve_bloch) without ever writing the eigenbasis array.
"""

from __future__ import annotations

from typing import Any

import numpy as np
import torch

from slate_quantum_b200 import kernels
from slate_quantum_b200.bloch._transposed_basis import BlochTransposedBasis
from slate_quantum_b200.configs import HBAR as hbar_value
from slate_quantum_b200.core import array
from slate_quantum_b200.core import basis as _basis
from slate_quantum_b200.core.basis import Basis, TupleBasis, _strip
from slate_quantum_b200.operator._linalg import into_diagonal_hermitian
from slate_quantum_b200.operator._operator import Operator
from slate_quantum_b200.state._state import State, StateList


class EvolvedStateList(StateList):
    """StateList over (times, eigenbasis) whose [T, M] data is generated on demand."""

    def __init__(
        self, basis: Basis, coefficients: torch.Tensor, eigenvalues: torch.Tensor, times: torch.Tensor, hbar: float,
        uniform_dt: float | None = None,
    ) -> None:
        self._basis = basis
        self._host = None
        self._device = None
        self._coefficients = coefficients
        self._eigenvalues = eigenvalues  # [n_blocks, n] float64
        self._times = times
        self._hbar = hbar
        self._uniform_dt = uniform_dt  # spacing of a uniform time grid: enables the table-driven 3M GEMM

    @property
    def device_data(self) -> torch.Tensor:
        if self._device is None:
            self._device = kernels.evolve_eigen(
                self._eigenvalues.reshape(-1), self._coefficients.reshape(-1), self._times, self._hbar
            ).reshape(-1)
        return self._device

    @property
    def raw_data(self) -> np.ndarray:
        if self._host is None:
            self._host = self.device_data.detach().cpu().numpy().reshape(self._times.numel(), -1)
        return self._host

    def _eigenbasis(self):
        return _strip(_basis.as_tuple(self.basis).children[1])

    def with_state_basis(self, basis: Basis) -> StateList:
        """Fused back-transformation when `basis` shares the eigenbasis' inner (Bloch) basis family."""
        eig = self._eigenbasis()
        inner = _strip(eig.inner)
        blocks = eig._blocks()  # noqa: SLF001
        times_basis = _basis.as_tuple(self.basis).children[0]
        out_basis = TupleBasis((times_basis, basis)).upcast()
        if eig.is_dual:
            return super().with_state_basis(basis)
        states = kernels.evolve_bloch(blocks, self._eigenvalues.reshape(blocks.shape[0], -1).contiguous(),
                                      self._coefficients.reshape(blocks.shape[0], -1).contiguous(),
                                      self._times, self._hbar, uniform_dt=self._uniform_dt)
        data = inner.convert_vector_into(states, basis, -1) if inner != basis else states
        return StateList(out_basis, data.reshape(-1))

    def with_basis(self, basis: Basis) -> StateList:
        tup = _basis.as_tuple(basis)
        if tup.children[0] == _basis.as_tuple(self.basis).children[0]:
            return self.with_state_basis(tup.children[1])
        return StateList(self.basis, self.device_data).with_basis(basis)


def _time_values(times: Basis) -> np.ndarray:
    """reference schrodinger.py:50: take the values from the metadata object, never recompute them."""
    values = np.array(list(times.metadata().values))
    points = getattr(_strip(times), "points", np.arange(times.size))
    return values[points]


def _solve_schrodinger_equation_diagonal(
    initial_state: State, times: Basis, hamiltonian: Any, *, hbar: float = hbar_value
) -> StateList:
    """reference schrodinger.py:31-56."""
    eigenbasis = hamiltonian.basis.inner.children[0]
    coefficients = initial_state.with_basis(eigenbasis).device_data.reshape(-1)
    n_total = coefficients.numel()
    eig_real = getattr(hamiltonian, "eigenvalues_real", None)
    if eig_real is None:
        eig_real = hamiltonian.device_data.real.contiguous()
    blocks = _strip(eigenbasis)._blocks()  # noqa: SLF001
    eig_real = eig_real.reshape(blocks.shape[0], n_total // blocks.shape[0]).contiguous()
    host_times = _time_values(times).astype(np.float64)
    time_values = kernels.to_device(host_times, torch.float64)
    out_basis = TupleBasis((times, eigenbasis)).upcast()
    return EvolvedStateList(out_basis, coefficients, eig_real, time_values, hbar, kernels.uniform_step(host_times))


def solve_schrodinger_equation_decomposition(
    initial_state: State, times: Basis, hamiltonian: Operator, *, hbar: float = hbar_value
) -> StateList:
    """Solve the Schrodinger equation by directly finding eigenstates (reference schrodinger.py:59-73)."""
    diagonal = into_diagonal_hermitian(hamiltonian)
    cast = array.cast_basis(diagonal, diagonal.basis.inner)
    cast.eigenvalues_real = diagonal.eigenvalues_real
    return _solve_schrodinger_equation_diagonal(initial_state, times, cast, hbar=hbar)


def solve_schrodinger_equation(*_args: Any, **_kwargs: Any) -> StateList:
    """The qutip ODE path of the reference (schrodinger.py:76-123) is out of scope."""
    msg = "solve_schrodinger_equation (qutip sesolve) is not part of the B200 hot path; use the decomposition solver"
    raise NotImplementedError(msg)


__all__ = ["BlochTransposedBasis", "solve_schrodinger_equation", "solve_schrodinger_equation_decomposition"]
