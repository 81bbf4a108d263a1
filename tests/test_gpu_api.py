"""The reference's own tests for the hot path, re-run against the drop-in API (GPU).

Each test is the counterpart of a test in /root/reference/tests (cited), with `slate_quantum` ->
`slate_quantum_b200` and `slate_core` -> `slate_quantum_b200.core`; tolerances are the reference's.
"""

import itertools

import numpy as np
import pytest

from oracle import bloch_oracle as o

pytestmark = pytest.mark.gpu

hbar = o.HBAR_SI
SHAPES_1D = [(x,) for x in [3, 5]]
SHAPES_2D = [*itertools.product([3, 4], [3, 4]), (5, 5)]
GRID = [*itertools.product(SHAPES_1D, SHAPES_1D), *itertools.product(SHAPES_2D, SHAPES_2D)]


def _api():
    from slate_quantum_b200 import bloch, dynamics, operator, state
    from slate_quantum_b200.core import TupleBasis, basis, metadata

    return bloch, dynamics, operator, state, TupleBasis, basis, metadata


@pytest.mark.parametrize(("shape", "repeat"), GRID)
def test_build_bloch_operator(cuda, shape, repeat):
    """reference tests/test_bloch.py:16-47."""
    bloch, _, operator, _, _, basis, metadata = _api()
    vectors = 2 * np.pi * np.eye(len(shape))
    meta = metadata.volume.spaced_volume_metadata_from_stacked_delta_x(
        tuple(vectors[i] for i in range(len(shape))), shape
    )
    potential = operator.build.cos_potential(meta, 1)
    mass = hbar**2

    sparse = bloch.build.kinetic_hamiltonian(potential, mass, repeat)
    repeat_potential = operator.build.repeat_potential(potential, repeat)
    full = operator.build.kinetic_hamiltonian(repeat_potential, mass)

    assert sparse.basis.metadata() == full.basis.metadata()
    transformed = basis.transformed_from_metadata(full.basis.metadata(), is_dual=full.basis.is_dual).upcast()
    np.testing.assert_allclose(
        sparse.with_basis(transformed).raw_data.reshape(transformed.inner.shape),
        full.with_basis(transformed).raw_data.reshape(transformed.inner.shape),
        atol=1e-14,
    )
    np.testing.assert_allclose(sparse.as_array(), full.as_array(), atol=1e-14)
    # and against the oracle's dense supercell matrix
    comps = o.cos_potential_components(shape, 1.0)
    want = o.supercell_hamiltonian(vectors, shape, comps, mass, repeat)
    np.testing.assert_allclose(
        sparse.with_basis(transformed).raw_data.reshape(transformed.inner.shape), want, atol=1e-14
    )


@pytest.mark.parametrize(("shape", "repeat"), GRID)
def test_build_momentum_bloch_operator(cuda, shape, repeat):
    """reference tests/test_bloch.py:50-86."""
    bloch, _, operator, _, _, basis, metadata = _api()
    vectors = 2 * np.pi * np.eye(len(shape))
    meta = metadata.volume.spaced_volume_metadata_from_stacked_delta_x(
        tuple(vectors[i] for i in range(len(shape))), shape
    )
    potential = operator.build.cos_potential(meta, 0)
    mass = hbar**2
    sparse = bloch.build.kinetic_hamiltonian(potential, mass, repeat)
    full = operator.build.kinetic_hamiltonian(operator.build.repeat_potential(potential, repeat), mass)
    assert sparse.basis.metadata() == full.basis.metadata()
    transformed = basis.transformed_from_metadata(full.basis.metadata(), is_dual=full.basis.is_dual).upcast()
    np.testing.assert_allclose(
        np.diag(full.with_basis(transformed).raw_data.reshape(transformed.inner.shape)),
        np.diag(sparse.with_basis(transformed).raw_data.reshape(transformed.inner.shape)),
        atol=1e-15,
    )
    np.testing.assert_allclose(sparse.as_array(), full.as_array(), atol=1e-14)


@pytest.mark.parametrize(("shape", "repeat"), GRID)
def test_build_potential_bloch_operator(cuda, shape, repeat):
    """reference tests/test_bloch.py:89-126."""
    bloch, _, operator, _, TupleBasis, basis, metadata = _api()
    from slate_quantum_b200.bloch.build import BlochFractionMetadata

    vectors = 2 * np.pi * np.eye(len(shape))
    meta = metadata.volume.spaced_volume_metadata_from_stacked_delta_x(
        tuple(vectors[i] for i in range(len(shape))), shape
    )
    potential = operator.build.cos_potential(meta, 1)
    fraction_basis = basis.from_metadata(BlochFractionMetadata.from_repeats(repeat)).upcast()
    operator_list = operator.OperatorList(
        TupleBasis((fraction_basis, potential.basis)).upcast(),
        np.tile(potential.raw_data, np.prod(repeat).item()),
    )
    sparse = bloch.build.bloch_operator_from_list(operator_list)
    full = operator.build.repeat_potential(potential, repeat)
    assert sparse.basis.metadata() == full.basis.metadata()
    transformed = basis.transformed_from_metadata(full.basis.metadata(), is_dual=full.basis.is_dual).upcast()
    np.testing.assert_allclose(
        sparse.with_basis(transformed).raw_data.reshape(transformed.inner.shape),
        full.with_basis(transformed).raw_data.reshape(transformed.inner.shape),
        atol=1e-15,
    )


def test_bloch_operator_from_list(cuda):
    """reference tests/test_bloch.py:129-151."""
    bloch, _, operator, _, TupleBasis, basis, metadata = _api()
    from slate_quantum_b200.bloch.build import BlochFractionMetadata
    from slate_quantum_b200.operator import Operator

    meta = metadata.volume.spaced_volume_metadata_from_stacked_delta_x((np.array([2 * np.pi]),), (3,))
    operator_basis = basis.transformed_from_metadata(meta).upcast()
    fraction_basis = basis.from_metadata(BlochFractionMetadata.from_repeats((2,))).upcast()
    operator_basis_tuple = TupleBasis((operator_basis, operator_basis.dual_basis())).upcast()
    operator_0 = Operator(operator_basis_tuple, np.ones((3, 3), dtype=complex))
    operator_1 = Operator(operator_basis_tuple, 2 * np.ones((3, 3), dtype=complex))
    operator_list = operator.OperatorList(
        TupleBasis((fraction_basis, operator_basis_tuple)).upcast(),
        np.array([operator_0.raw_data, operator_1.raw_data], dtype=complex),
    )
    sparse = bloch.build.bloch_operator_from_list(operator_list)
    basis.transformed_from_metadata(sparse.basis.metadata(), is_dual=sparse.basis.is_dual)
    assert sparse.raw_data.size == 2 * 9


def test_build_kinetic_operator(cuda):
    """reference tests/test_operator.py:22-28."""
    _, _, operator, _, _, _, metadata = _api()
    meta = metadata.volume.spaced_volume_metadata_from_stacked_delta_x((np.array([2 * np.pi]),), (5,))
    kinetic_operator = operator.build.kinetic_energy(meta, hbar**2)
    np.testing.assert_allclose(kinetic_operator.raw_data, [0.0, 0.5, 2.0, 2.0, 0.5])


def test_build_hamiltonian(cuda):
    """reference tests/test_operator.py:31-59."""
    _, _, operator, _, TupleBasis, basis, metadata = _api()
    meta = metadata.volume.spaced_volume_metadata_from_stacked_delta_x((np.array([2 * np.pi]),), (5,))
    potential = operator.build.cos_potential(meta, 0)
    hamiltonian = operator.build.kinetic_hamiltonian(potential, hbar**2)
    np.testing.assert_allclose(hamiltonian.raw_data, [0.0, 0.0, 0.0, 0.0, 0.5, 2.0, 2.0, 0.5])
    transformed_basis = basis.transformed_from_metadata(meta)
    transformed_operator = hamiltonian.with_basis(
        TupleBasis((transformed_basis, transformed_basis.dual_basis())).upcast()
    )
    expected = np.diag([0.0, 0.5, 2.0, 2.0, 0.5])
    np.testing.assert_allclose(transformed_operator.raw_data, np.ravel(expected).astype(np.complex128), atol=1e-15)
    kinetic_operator = operator.build.kinetic_energy(meta, hbar**2)
    np.testing.assert_allclose(kinetic_operator.as_array(), hamiltonian.as_array(), atol=1e-15)


def test_hamiltonian_eigenstates(cuda):
    """reference tests/test_operator.py:62-75."""
    _, _, operator, _, _, _, metadata = _api()
    meta = metadata.volume.spaced_volume_metadata_from_stacked_delta_x((np.array([2 * np.pi]),), (5,))
    potential = operator.build.cos_potential(meta, 0)
    hamiltonian = operator.build.kinetic_hamiltonian(potential, hbar**2)
    kinetic_operator = operator.build.kinetic_energy(meta, hbar**2)
    np.testing.assert_allclose(
        np.sort(operator.into_diagonal_hermitian(kinetic_operator).raw_data),
        np.sort(operator.into_diagonal_hermitian(hamiltonian).raw_data),
        atol=1e-15,
    )
    np.testing.assert_allclose(
        np.sort(operator.into_diagonal_hermitian(hamiltonian).raw_data.real), [0.0, 0.5, 0.5, 2.0, 2.0], atol=1e-14
    )


def test_build_cos_and_sin_operator(cuda):
    """reference tests/test_operator.py:305-320."""
    _, _, operator, _, _, _, metadata = _api()
    meta = metadata.volume.spaced_volume_metadata_from_stacked_delta_x((np.array([2 * np.pi, 0]),), (16,))
    actual = operator.build.cos_potential(meta, 2)
    expected = operator.build.potential_from_function(meta, lambda x: 1 + np.cos(x[0]))
    np.testing.assert_allclose(actual.as_array(), expected.as_array(), atol=1e-14)
    actual = operator.build.sin_potential(meta, 2)
    expected = operator.build.potential_from_function(meta, lambda x: 1 + np.sin(x[0]))
    np.testing.assert_allclose(actual.as_array(), expected.as_array(), atol=1e-14)


def test_build_fcc_operator(cuda):
    """reference tests/test_operator.py:323-353."""
    _, _, operator, _, _, _, metadata = _api()
    meta = metadata.volume.spaced_volume_metadata_from_stacked_delta_x(
        (np.array([np.sin(np.pi / 3), np.cos(np.pi / 3)]), np.array([0, 1])), (16, 16)
    )
    actual = operator.build.fcc_potential(meta, 1)
    eta = 2 * np.pi * (2 / np.sqrt(3))
    expected = operator.build.potential_from_function(
        meta,
        lambda x: (
            (1 / 3)
            + (2 / 9)
            * (
                np.cos(eta * x[0] + 0 * x[1])
                + np.cos(eta * np.cos(np.pi / 3) * x[0] + eta * np.sin(np.pi / 3) * x[1])
                + np.cos(-eta * np.cos(np.pi / 3) * x[0] + eta * np.sin(np.pi / 3) * x[1])
            )
        ),
    )
    np.testing.assert_allclose(actual.as_array(), expected.as_array(), atol=1e-13)


@pytest.mark.parametrize(("shape", "repeat"), [((8,), (6,)), ((4, 3), (3, 4)), ((6, 6), (4, 4))])
def test_bloch_eigenstates_and_schrodinger_evolution(cuda, shape, repeat):
    """examples/blochs_theorem.py + examples/schrodinger.py through the API, against the oracle and expm.

    (The reference has no test for dynamics/schrodinger.py; the oracle is cross-checked against
    scipy.linalg.expm of the dense supercell Hamiltonian.)
    """
    import scipy.linalg

    bloch, dynamics, operator, state, _, basis, metadata = _api()
    from slate_quantum_b200.core import FundamentalBasis
    from slate_quantum_b200.core.metadata import Domain
    from slate_quantum_b200.metadata import SpacedTimeMetadata, repeat_volume_metadata

    nd = len(shape)
    vectors = 2 * np.pi * np.eye(nd)
    meta = metadata.volume.spaced_volume_metadata_from_stacked_delta_x(tuple(vectors[i] for i in range(nd)), shape)
    potential = operator.build.sin_potential(meta, 3.0, offset=0.3)
    hamiltonian = bloch.build.kinetic_hamiltonian(potential, hbar**2, repeat)

    diagonal = bloch.into_diagonal(hamiltonian)
    comps = o.sin_potential_components(shape, 3.0, 0.3)
    h_super = o.supercell_hamiltonian(vectors, shape, comps, hbar**2, repeat)
    np.testing.assert_allclose(np.sort(diagonal.raw_data.real), np.linalg.eigvalsh(h_super), atol=1e-10)
    assert np.all(diagonal.raw_data.imag == 0)

    super_meta = repeat_volume_metadata(meta, repeat)
    big = tuple(n * r for n, r in zip(shape, repeat))
    initial_state = state.build.coherent(
        super_meta, tuple(np.pi * r for r in repeat), (0.0,) * nd, (1.5,) * nd
    )
    times = FundamentalBasis(SpacedTimeMetadata(12, domain=Domain(delta=3 * hbar)))
    evolution = dynamics.solve_schrodinger_equation_decomposition(initial_state, times, hamiltonian)
    t_values = times.metadata().values
    position_basis = basis.from_metadata(super_meta).upcast()
    in_position = evolution.with_state_basis(position_basis).raw_data.reshape(12, -1)
    h_x = o.momentum_matrix_to_position(h_super, big)
    ref = np.array([scipy.linalg.expm(-1j * h_x * t / hbar) @ initial_state.raw_data for t in t_values])
    np.testing.assert_allclose(in_position, ref, atol=1e-9)
    # the reference's return value itself: amplitudes in the eigenbasis, [T, M]
    amplitudes = evolution.raw_data
    assert amplitudes.shape == (12, int(np.prod(big)))
    np.testing.assert_allclose(np.abs(amplitudes), np.broadcast_to(np.abs(amplitudes[0]), amplitudes.shape), atol=1e-12)
    np.testing.assert_allclose(np.linalg.norm(amplitudes, axis=1), 1.0, atol=1e-12)


def test_get_eigenstates_hermitian_bloch(cuda):
    """examples/blochs_theorem.py:33: dense eigenstate list of a Bloch Hamiltonian."""
    bloch, _, operator, _, _, basis, metadata = _api()
    meta = metadata.volume.spaced_volume_metadata_from_stacked_delta_x((np.array([2 * np.pi]),), (6,))
    rng = np.random.default_rng(0)
    potential = operator.build.potential_from_function(
        meta, lambda x: 4 * rng.normal(size=x[0].size).astype(np.complex128)
    )
    hamiltonian = bloch.build.kinetic_hamiltonian(potential, hbar**2, (3,))
    eigenstates = operator.get_eigenstates_hermitian(hamiltonian)
    vecs = eigenstates.raw_data.reshape(18, 18)
    energies = eigenstates.basis.metadata().children[0].values.real
    # states are expressed in the Bloch state basis; convert to the supercell momentum basis and check H v = E v
    transformed = basis.transformed_from_metadata(hamiltonian.basis.metadata(), is_dual=hamiltonian.basis.is_dual).upcast()
    h_dense = hamiltonian.with_basis(transformed).raw_data.reshape(18, 18)
    momentum = basis.transformed_from_metadata(hamiltonian.basis.metadata().children[0]).upcast()
    vk = eigenstates.with_state_basis(momentum).raw_data.reshape(18, 18)
    del vecs
    res = h_dense @ vk.T - vk.T * energies[None, :]
    assert np.abs(res).max() < 1e-10


def test_measure_api_coherent_states(cuda):
    """operator.measure mirror on the reference's own coherent-state checks (tests/test_measure.py:18-66)."""
    from slate_quantum_b200 import operator, state
    from slate_quantum_b200.core.metadata import spaced_volume_metadata_from_stacked_delta_x

    n = 300
    metadata = spaced_volume_metadata_from_stacked_delta_x((np.array([2 * np.pi]),), (n,))
    dx = 2 * np.pi / n
    sigma_0 = (0.1 * np.pi,)
    centre = state.build.coherent(metadata, (150 * dx,), (0,), sigma_0)
    np.testing.assert_allclose(operator.measure.x(centre, axis=0), 150 * dx, rtol=1e-3)
    np.testing.assert_allclose(operator.measure.coherent_width(centre, axis=0), sigma_0[0], atol=1e-7 * dx)
    states = state.StateList.from_states(
        [state.build.coherent(metadata, (i * dx,), (0,), sigma_0) for i in range(0, n, 11)]
    )
    expected = np.arange(0, n, 11) * dx
    periodic = operator.measure.all_periodic_x(states, axis=0).raw_data
    difference = (periodic - expected + np.pi) % (2 * np.pi) - np.pi
    np.testing.assert_allclose(difference, 0, atol=1e-7 * dx)
    widths = operator.measure.all_coherent_width(states, axis=0).raw_data
    np.testing.assert_allclose(widths, sigma_0[0], atol=1e-7 * dx)
    for i, s in enumerate(states):
        wrapped = operator.measure.x(s, axis=0, offset=-expected[i], wrapped=True)
        np.testing.assert_allclose(wrapped, 0, atol=1e-7 * dx)


@pytest.mark.parametrize("shape", [(24,), (6, 5)])
def test_dense_hamiltonian_schrodinger_evolution_and_observables(cuda, shape):
    """the reference's examples/schrodinger.py on ONE cell (no Bloch decomposition):
    operator.build.kinetic_hamiltonian -> into_diagonal_hermitian (batch of one matrix) -> decomposition
    solver -> position basis -> operator.measure, against scipy.linalg.expm and the oracle's observables."""
    import scipy.linalg

    _, dynamics, operator, state, _, basis, metadata = _api()
    from slate_quantum_b200.core import FundamentalBasis
    from slate_quantum_b200.core.metadata import Domain
    from slate_quantum_b200.metadata import SpacedTimeMetadata

    nd = len(shape)
    vectors = 2 * np.pi * np.eye(nd)
    meta = metadata.volume.spaced_volume_metadata_from_stacked_delta_x(tuple(vectors[i] for i in range(nd)), shape)
    potential = operator.build.cos_potential(meta, 4.0)
    hamiltonian = operator.build.kinetic_hamiltonian(potential, hbar**2)
    comps = o.cos_potential_components(shape, 4.0)
    h_k = o.dense_hamiltonian(vectors, shape, o.pad_components(shape, comps), hbar**2)
    diagonal = operator.into_diagonal_hermitian(hamiltonian)
    np.testing.assert_allclose(np.sort(diagonal.raw_data.real), np.linalg.eigvalsh(h_k), atol=1e-10)

    initial_state = state.build.coherent(meta, (np.pi,) * nd, (1.0,) + (0.0,) * (nd - 1), (0.8,) * nd)
    times = FundamentalBasis(SpacedTimeMetadata(9, domain=Domain(delta=2 * hbar)))
    evolution = dynamics.solve_schrodinger_equation_decomposition(initial_state, times, hamiltonian)
    in_position_list = evolution.with_state_basis(basis.from_metadata(meta).upcast())
    in_position = in_position_list.raw_data.reshape(9, -1)
    h_x = o.momentum_matrix_to_position(h_k, shape)
    ref = np.array([scipy.linalg.expm(-1j * h_x * t / hbar) @ initial_state.raw_data for t in times.metadata().values])
    np.testing.assert_allclose(in_position, ref, atol=1e-9)
    for axis in range(nd):
        got = operator.measure.all_periodic_x(in_position_list, axis=axis).raw_data
        np.testing.assert_allclose(got, o.measure_all_periodic_x(ref, vectors, shape, axis), atol=1e-8)
        got_w = operator.measure.all_coherent_width(in_position_list, axis=axis).raw_data
        np.testing.assert_allclose(got_w, o.measure_all_coherent_width(ref, vectors, shape, axis), atol=1e-8)


def test_measure_api_momentum_observables(cuda):
    """operator.measure.{k, all_k, all_variance_k, all_uncertainty} (reference operator/_measure.py:307-420) against
    the oracle and the analytic values of a minimum-uncertainty wavepacket."""
    from slate_quantum_b200 import operator, state
    from slate_quantum_b200.core.metadata import spaced_volume_metadata_from_stacked_delta_x

    shape = (48, 40)
    lengths = (30.0, 24.0)
    metadata = spaced_volume_metadata_from_stacked_delta_x(
        (np.array([lengths[0], 0.0]), np.array([0.0, lengths[1]])), shape
    )
    delta_x = np.diag(lengths)
    k0 = (2 * np.pi * 5 / lengths[0], -2 * np.pi * 3 / lengths[1])
    sigmas = [(1.1, 1.4), (0.9, 1.2), (1.5, 1.0)]
    states = state.StateList.from_states(
        [state.build.coherent(metadata, (11.0, 9.0), k0, sg) for sg in sigmas]
    )
    raw = np.array([s.raw_data for s in states])
    for axis in (0, 1):
        got_k = operator.measure.all_k(states, axis=axis).raw_data
        np.testing.assert_allclose(got_k, o.measure_all_k(raw, delta_x, shape, axis), atol=1e-10)
        np.testing.assert_allclose(got_k, k0[axis], atol=1e-4)  # analytic value, up to the grid's aliasing
        got_v = operator.measure.all_variance_k(states, axis=axis).raw_data
        np.testing.assert_allclose(got_v, o.measure_all_variance_k(raw, delta_x, shape, axis), rtol=1e-9)
        np.testing.assert_allclose(got_v, [1 / (2 * sg[axis] ** 2) for sg in sigmas], rtol=1e-3)
        got_u = operator.measure.all_uncertainty(states, axis=axis).raw_data
        np.testing.assert_allclose(got_u, o.measure_all_uncertainty(raw, delta_x, shape, axis), rtol=1e-9)
        np.testing.assert_allclose(got_u, 0.5, rtol=1e-3)
    np.testing.assert_allclose(operator.measure.k(states[1], axis=0), k0[0], atol=1e-4)
    np.testing.assert_allclose(operator.measure.uncertainty(states[2], axis=1), 0.5, rtol=1e-3)
