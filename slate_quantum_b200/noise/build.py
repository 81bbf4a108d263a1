"""Noise-operator builders on top of K7 (reference: slate_quantum/noise/build/_build.py:141-175,
noise/build/_caldeira_leggett.py:173-184)."""

from __future__ import annotations

import numpy as np

from slate_quantum_b200 import kernels
from slate_quantum_b200.configs import HBAR as hbar
from slate_quantum_b200.core import basis as _basis
from slate_quantum_b200.core.basis import TupleBasis
from slate_quantum_b200.core.metadata import TupleMetadata
from slate_quantum_b200.metadata import eigenvalue_basis
from slate_quantum_b200.operator import build as _operator_build
from slate_quantum_b200.operator._linalg import _dense_list, commute, get_commutator_operator_list
from slate_quantum_b200.operator._operator import Operator, OperatorList

Boltzmann = 1.380649e-23  # J / K, exact (scipy.constants.Boltzmann)


def caldeira_leggett_operators(metadata: TupleMetadata) -> OperatorList:
    """The single Caldeira-Leggett noise operator of a 1-d system: x, labelled by eigenvalue 1
    (reference noise/build/_caldeira_leggett.py:173-184)."""
    assert len(metadata.fundamental_shape) == 1
    operators = OperatorList.from_operators([_operator_build.x(metadata, axis=0)])
    inner = _basis.as_tuple(operators.basis).children[1]
    return OperatorList(TupleBasis((eigenvalue_basis(np.array([1])), inner)).upcast(), operators.raw_data)


def temperature_corrected_operators(
    hamiltonian: Operator, operators: OperatorList, temperature: float, eta: float = 1.0
) -> OperatorList:
    """sqrt(2 eta kT / hbar^2) L_m - sqrt(eta / (8 kT hbar^2)) [H, L_m] for every noise operator L_m
    (reference noise/build/_build.py:141-155; like the reference, the result is multiplied by hbar's scale,
    i.e. not divided back).  The commutators are two batched complex GEMMs (K7) over the whole list."""
    commutator = get_commutator_operator_list(hamiltonian, operators)
    kb_t = Boltzmann * temperature
    correction = commutator * (-1 * np.sqrt(eta / (8 * kb_t * hbar**2)))
    scaled = operators * np.sqrt(2 * eta * kb_t / hbar**2)
    return correction + scaled


def hamiltonian_shift(hamiltonian: Operator, operators: OperatorList, eta: float) -> Operator:
    """i eta / (4 hbar) [H, sum_m L_m^dagger L_m] (reference noise/build/_build.py:158-172).  The sum over the list
    is folded into the GEMM's inner dimension: (L as [(m k), j])^H @ (L as [(m k), l])."""
    _, op_basis, data = _dense_list(operators)
    n = op_basis.shape[0]
    flat = data.reshape(-1, n)
    shift_product = Operator(op_basis, kernels.zgemm(flat.mH, flat).reshape(-1))
    return commute(hamiltonian, shift_product) * (1j * eta / (4 * hbar))
