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


@pytest.mark.parametrize("shape", [(24,), (6, 5), (37,), (3, 41)])
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


# ---------------------------------------------------------------------------------------- K7 operator algebra
def _meta_1d(length, n):
    _, _, _, _, _, _, metadata = _api()
    return metadata.volume.spaced_volume_metadata_from_stacked_delta_x((np.array([length]),), (n,))


def test_x_and_k_operators(cuda):
    """reference tests/test_operator.py:170-210."""
    _, _, operator, _, TupleBasis, basis, _ = _api()
    meta = _meta_1d(2 * np.pi, 5)
    x = operator.build.x(meta, axis=0)
    np.testing.assert_allclose(x.as_array(), np.diag(np.linspace(0, 2 * np.pi, 5, endpoint=False)), atol=1e-15)
    k = operator.build.k(meta, axis=0)
    basis_k = basis.transformed_from_metadata(meta)
    np.testing.assert_allclose(
        k.with_basis(TupleBasis((basis_k, basis_k.dual_basis())).upcast()).raw_data.reshape(5, 5),
        np.diag([0, 1, 2, -2, -1]), atol=1e-15,
    )
    np.testing.assert_allclose(k.as_array(), o.k_operator_array(np.array([[2 * np.pi]]), (5,), 0), atol=1e-15)
    np.testing.assert_allclose((k * hbar).as_array(), operator.build.p(meta, axis=0).as_array(), atol=1e-15)
    np.testing.assert_allclose(operator.build.k_squared(meta, axis=0).as_array(),
                               o.k_operator_array(np.array([[2 * np.pi]]), (5,), 0, power=2), atol=1e-14)


def test_x_k_commutator(cuda):
    """reference tests/test_operator.py:213-245."""
    _, _, operator, _, _, _, _ = _api()
    meta = _meta_1d(2 * np.pi, 5)
    x, k = operator.build.x(meta, axis=0), operator.build.k(meta, axis=0)
    matmul_0 = operator.matmul(x, k)
    np.testing.assert_allclose(matmul_0.as_array(), x.as_array() @ k.as_array(), atol=1e-15)
    matmul_1 = operator.matmul(k, x)
    np.testing.assert_allclose(matmul_1.as_array(), k.as_array() @ x.as_array(), atol=1e-15)
    manual = matmul_0 - matmul_1
    np.testing.assert_allclose(manual.as_array(), x.as_array() @ k.as_array() - k.as_array() @ x.as_array(), atol=1e-15)
    commutator = operator.commute(x, k)
    np.testing.assert_allclose(commutator.as_array(), manual.as_array(), atol=1e-15)
    np.testing.assert_allclose(commutator.as_array(),
                               o.operator_commutator(o.x_operator_array(np.array([[2 * np.pi]]), (5,), 0),
                                                     o.k_operator_array(np.array([[2 * np.pi]]), (5,), 0)), atol=1e-14)


def test_trivial_commutator(cuda):
    """reference tests/test_operator.py:248-259: exactly zero."""
    _, _, operator, _, _, _, _ = _api()
    meta = _meta_1d(2 * np.pi, 5)
    x, k = operator.build.x(meta, axis=0), operator.build.k(meta, axis=0)
    np.testing.assert_array_equal(operator.commute(x, x).as_array(), np.zeros((5, 5)))
    np.testing.assert_array_equal(operator.commute(k, k).as_array(), np.zeros((5, 5)))


def test_dagger(cuda):
    """reference tests/test_operator.py:479-499 (x and a non-hermitian operator)."""
    _, _, operator, _, _, _, _ = _api()
    meta = _meta_1d(2 * np.pi, 5)
    x = operator.build.x(meta, axis=0)
    np.testing.assert_allclose(operator.dagger(x).as_array(), x.as_array().T.conj())
    xk = operator.matmul(x, operator.build.k(meta, axis=0))  # not hermitian
    np.testing.assert_allclose(operator.dagger(xk).as_array(), xk.as_array().T.conj(), atol=1e-15)
    assert operator.dagger(xk).basis.is_dual == xk.basis.is_dual


def test_commute(cuda):
    """reference tests/test_operator.py:509-540 (300 points; tolerances are the reference's)."""
    _, _, operator, _, _, _, _ = _api()
    meta = _meta_1d(2 * np.pi * 10, 300)

    def commutator_numpy(a, b):
        return a @ b - b @ a

    x, k = operator.build.x(meta, axis=0), operator.build.k(meta, axis=0)
    np.testing.assert_allclose(operator.commute(x, k).as_array(), commutator_numpy(x.as_array(), k.as_array()),
                               atol=1e-13)
    mass = hbar**2
    kinetic = operator.build.kinetic_energy(meta, mass)
    np.testing.assert_allclose(operator.commute(kinetic, x).as_array(),
                               commutator_numpy(kinetic.as_array(), x.as_array()), atol=2e-12)
    hamiltonian = operator.build.kinetic_hamiltonian(operator.build.cos_potential(meta, 0), mass)
    np.testing.assert_allclose(operator.commute(hamiltonian, x).as_array(),
                               commutator_numpy(hamiltonian.as_array(), x.as_array()), atol=2e-12)
    # the canonical commutator on a smooth state (oracle: tests/test_algebra_oracle.py)
    _, _, _, state, _, basis, _ = _api()
    psi = o.coherent_state(np.array([[2 * np.pi * 10]]), (300,), (np.pi * 10,), (0.0,), (1.5,))
    s = state.State(basis.from_metadata(meta).upcast(), psi)
    got = operator.apply(operator.commute(x, k), s)
    np.testing.assert_allclose(got.with_basis(s.basis).raw_data, 1j * psi, atol=1e-9)
    np.testing.assert_allclose(operator.expectation(operator.commute(x, k), s), 1j * np.vdot(psi, psi), atol=1e-9)


def test_operator_list_products_and_commutators(cuda):
    """matmul_list_operator / matmul_operator_list / get_commutator_operator_list / dagger_each
    (reference operator/_linalg.py:150-215) against the oracle, on a list [x, k, x k]."""
    _, _, operator, _, _, _, _ = _api()
    n, length = 48, 7.0
    meta = _meta_1d(length, n)
    dxm = np.array([[length]])
    x, k = operator.build.x(meta, axis=0), operator.build.k(meta, axis=0)
    fund = x.with_basis(operator.matmul(x, k).basis)  # dense (fundamental, fundamental.dual) operator basis
    ops = operator.OperatorList.from_operators([fund, k.with_basis(fund.basis), operator.matmul(x, k)])
    xa, ka = o.x_operator_array(dxm, (n,), 0), o.k_operator_array(dxm, (n,), 0)
    ops_np = np.array([xa, ka, xa @ ka])
    h = operator.build.kinetic_hamiltonian(operator.build.cos_potential(meta, 3.0), 1.0, hbar=1.0)
    ha = h.as_array()
    scale = np.abs(ha).max() * np.abs(ops_np).max() * n

    def arr(lst):
        return np.array([op.as_array() for op in lst])

    np.testing.assert_allclose(arr(operator.matmul_operator_list(h, ops)), o.operator_matmul(ha, ops_np),
                               rtol=0, atol=1e-14 * scale)
    np.testing.assert_allclose(arr(operator.matmul_list_operator(ops, h)), o.operator_matmul(ops_np, ha),
                               rtol=0, atol=1e-14 * scale)
    np.testing.assert_allclose(arr(operator.get_commutator_operator_list(h, ops)), o.operator_commutator(ha, ops_np),
                               rtol=0, atol=1e-14 * scale)
    np.testing.assert_allclose(arr(operator.dagger_each(ops)), o.operator_dagger(ops_np), rtol=0,
                               atol=1e-14 * np.abs(ops_np).max())
    assert len(operator.get_commutator_operator_list(h, ops)) == 3


def test_state_occupations_and_inner_products(cuda):
    """reference tests/test_state.py:20-74 + the list forms (state/_state.py:208-308)."""
    _, _, _, state, TupleBasis, basis, _ = _api()
    meta = _meta_1d(2 * np.pi, 5)
    state_basis = basis.from_metadata(meta)
    data = np.zeros(5, dtype=np.complex128)
    data[0] = data[1] = (1 - 1j) / np.sqrt(2)
    s0 = state.State(state_basis, data)
    occ = state.get_occupations(s0)
    np.testing.assert_allclose(occ.raw_data, [1, 1, 0, 0, 0], atol=1e-15)
    assert occ.basis.metadata().basis == state_basis
    momentum_basis = basis.TransformedBasis(state_basis)
    occ = state.get_occupations(state.State(momentum_basis, data))
    np.testing.assert_allclose(occ.raw_data, [1, 1, 0, 0, 0], atol=1e-15)
    assert occ.basis.metadata().basis == momentum_basis
    data1 = data.copy()
    data1[1] = -data[1]
    s1 = state.State(state_basis, data1)
    np.testing.assert_allclose(state.inner_product(s1, s0), 0, atol=1e-15)
    np.testing.assert_allclose(state.normalization(s0), 2, atol=1e-15)
    np.testing.assert_allclose(state.normalization(state.normalize(s0)), 1, atol=1e-15)
    both = state.StateList.from_states([s0, s1])
    both_np = np.array([data, data1])
    np.testing.assert_allclose(state.inner_product_each(both, both).raw_data, o.inner_product_each(both_np, both_np),
                               atol=1e-15)
    np.testing.assert_allclose(state.all_inner_product(both, both).raw_data.reshape(2, 2),
                               o.all_inner_product(both_np, both_np), atol=1e-15)
    np.testing.assert_allclose(state.normalization_each(both).raw_data, [2, 2], atol=1e-15)
    np.testing.assert_allclose(state.normalize_each(both).raw_data.reshape(2, 5), o.normalize_each(both_np), atol=1e-15)
    np.testing.assert_allclose(state.get_all_occupations(both).raw_data.reshape(2, 5), o.occupations(both_np), atol=1e-15)
    # a list held in the momentum basis is compared in the first list's basis
    both_k = both.with_state_basis(momentum_basis)
    np.testing.assert_allclose(state.inner_product_each(both, both_k).raw_data, [2, 2], atol=1e-14)
    np.testing.assert_allclose(state.all_inner_product(both_k, both).raw_data.reshape(2, 2), [[2, 0], [0, 2]], atol=1e-14)


def test_expectation_of_each_on_evolved_states(cuda):
    """expectation_of_each (reference operator/_operator.py:196-207) of x, k and H over a Schrodinger-evolved list:
    equals the K6 observables for x and k, and <H> is conserved."""
    _, dynamics, operator, state, TupleBasis, basis, metadata = _api()
    n, length = 128, 16.0
    meta = _meta_1d(length, n)
    dxm = np.array([[length]])
    mass, hb = 1.0, 1.0
    h = operator.build.kinetic_hamiltonian(operator.build.cos_potential(meta, 2.0), mass, hbar=hb)
    psi = o.coherent_state(dxm, (n,), (6.0,), (1.2,), (0.9,))
    psi /= np.linalg.norm(psi)
    s = state.State(basis.from_metadata(meta).upcast(), psi)
    from slate_quantum_b200.core import FundamentalBasis
    from slate_quantum_b200.core.metadata import Domain
    from slate_quantum_b200.metadata import SpacedTimeMetadata

    times = FundamentalBasis(SpacedTimeMetadata(9, domain=Domain(delta=0.45)))
    evolved = dynamics.solve_schrodinger_equation_decomposition(s, times, h, hbar=hb)
    pos = evolved.with_state_basis(basis.from_metadata(meta).upcast()).raw_data.reshape(9, n)
    x, k = operator.build.x(meta, axis=0), operator.build.k(meta, axis=0)
    np.testing.assert_allclose(operator.expectation_of_each(x, evolved).raw_data,
                               o.expectation_of_each(o.x_operator_array(dxm, (n,), 0), pos), atol=1e-10)
    np.testing.assert_allclose(operator.expectation_of_each(x, evolved).raw_data.real,
                               operator.measure.all_x(evolved, axis=0).raw_data, atol=1e-10)
    np.testing.assert_allclose(operator.expectation_of_each(k, evolved).raw_data,
                               o.expectation_of_each(o.k_operator_array(dxm, (n,), 0), pos), atol=1e-10)
    np.testing.assert_allclose(operator.expectation_of_each(k, evolved).raw_data.real,
                               operator.measure.all_k(evolved, axis=0).raw_data, atol=1e-10)
    energy = operator.expectation_of_each(h, evolved).raw_data
    np.testing.assert_allclose(energy, energy[0], atol=1e-10)
    np.testing.assert_allclose(energy.imag, 0, atol=1e-12)
    # the same three through the dense-operator GEMM path (the diagonal forms above take the one-pass reduction)
    fund = basis.from_metadata(meta)
    dense_basis = TupleBasis((fund, fund.dual_basis()))
    for op in (x, k, h):
        dense = op.with_basis(dense_basis)
        np.testing.assert_allclose(operator.expectation_of_each(dense, evolved).raw_data,
                                   operator.expectation_of_each(op, evolved).raw_data, rtol=0,
                                   atol=1e-11 * max(1.0, np.abs(dense.raw_data).max()))
    # a diagonalised Hamiltonian is diagonal in its own eigenbasis: no basis change of the evolved list at all
    diagonal = operator.into_diagonal_hermitian(h)
    np.testing.assert_allclose(operator.expectation_of_each(diagonal, evolved).raw_data, energy, atol=1e-10)
    occ = state.get_all_occupations(evolved)  # eigenbasis occupations are constants of the motion
    occ = occ.raw_data.reshape(9, n)
    np.testing.assert_allclose(occ, np.broadcast_to(occ[0], occ.shape), atol=1e-13)


def test_build_cl_operator(cuda):
    """reference tests/test_operator.py:451-476, plus parity of the temperature correction and the Hamiltonian
    shift (noise/build/_build.py:141-172) with the oracle."""
    _, _, operator, _, _, _, _ = _api()
    from slate_quantum_b200 import noise

    n = 160
    meta = _meta_1d(2 * np.pi, n)
    dxm = np.array([[2 * np.pi]])
    temperature, mass = 1, 1
    operators = noise.build.caldeira_leggett_operators(meta)
    assert len(operators) == 1  # one operator for a 1-d system
    x_operator = operator.build.x(meta, axis=0)
    np.testing.assert_allclose(operators[0, :].as_array(), x_operator.as_array(), atol=1e-15)
    hamiltonian = operator.build.kinetic_energy(meta, mass)
    corrected = noise.build.temperature_corrected_operators(hamiltonian, operators, temperature)
    assert len(corrected) == 1
    commutator = operator.commute(operator.build.x(meta, axis=0), operator.build.p(meta, axis=0))
    np.testing.assert_allclose(commutator.as_array(), (1j * hbar), atol=1e-15)
    # parity with the oracle on the same dense arrays
    h_np, x_np = hamiltonian.as_array(), o.x_operator_array(dxm, (n,), 0)
    want = o.temperature_corrected_operators(h_np, x_np[np.newaxis], temperature)
    got = corrected[0].as_array()
    np.testing.assert_allclose(got, want[0], rtol=0, atol=1e-12 * np.abs(want).max())
    eta = 0.7
    shift = noise.build.hamiltonian_shift(hamiltonian, operators, eta)
    want_shift = o.hamiltonian_shift(h_np, x_np[np.newaxis], eta)
    np.testing.assert_allclose(shift.as_array(), want_shift, rtol=0, atol=1e-12 * np.abs(want_shift).max())
    # a longer list: [x, k, x k + k x]
    k_op = operator.build.k(meta, axis=0)
    xk = operator.matmul(x_operator, k_op)
    sym = xk + operator.dagger(xk)
    ops = operator.OperatorList.from_operators(
        [x_operator.with_basis(xk.basis), k_op.with_basis(xk.basis), sym.with_basis(xk.basis)]
    )
    ops_np = np.array([op.as_array() for op in ops])
    got = noise.build.temperature_corrected_operators(hamiltonian, ops, 300.0, eta)
    want = o.temperature_corrected_operators(h_np, ops_np, 300.0, eta)
    for m in range(3):
        np.testing.assert_allclose(got[m].as_array(), want[m], rtol=0, atol=1e-12 * np.abs(want[m]).max())
    got_shift = noise.build.hamiltonian_shift(hamiltonian, ops, eta).as_array()
    want_shift = o.hamiltonian_shift(h_np, ops_np, eta)
    np.testing.assert_allclose(got_shift, want_shift, rtol=0, atol=1e-11 * np.abs(want_shift).max())


def test_expectation_of_each_bloch_hamiltonian(cuda):
    """<H>, <V>, <T> of Bloch-evolved states through the diagonal one-pass path (no dense supercell operator):
    the sparse Bloch Hamiltonian's expectation equals the oracle's dense supercell one and is conserved."""
    bloch, dynamics, operator, state, TupleBasis, basis, metadata = _api()
    from slate_quantum_b200.core import FundamentalBasis
    from slate_quantum_b200.core.metadata import Domain
    from slate_quantum_b200.metadata import SpacedTimeMetadata, repeat_volume_metadata

    shape, repeat = (6, 4), (3, 4)
    vectors = 2 * np.pi * np.eye(2)
    meta = metadata.volume.spaced_volume_metadata_from_stacked_delta_x(tuple(vectors), shape)
    potential = operator.build.cos_potential(meta, 3.0)
    mass = hbar**2
    full_meta = repeat_volume_metadata(meta, repeat)
    big = tuple(a * b for a, b in zip(shape, repeat))
    rng = np.random.default_rng(21)
    psi = rng.normal(size=int(np.prod(big))) + 1j * rng.normal(size=int(np.prod(big)))
    psi /= np.linalg.norm(psi)
    initial = state.State(basis.from_metadata(full_meta).upcast(), psi)
    times = FundamentalBasis(SpacedTimeMetadata(7, domain=Domain(delta=3 * hbar)))
    hamiltonian = bloch.build.kinetic_hamiltonian(potential, mass, repeat)
    evolved = dynamics.solve_schrodinger_equation_decomposition(initial, times, hamiltonian)
    pos = evolved.with_state_basis(basis.from_metadata(full_meta).upcast()).raw_data.reshape(7, -1)
    comps = o.cos_potential_components(shape, 3.0)
    h_x = o.momentum_matrix_to_position(o.supercell_hamiltonian(vectors, shape, comps, mass, repeat), big)
    want = o.expectation_of_each(h_x, pos)
    np.testing.assert_allclose(want, want[0], atol=1e-10 * abs(want[0]))
    # the repeated potential + kinetic split Hamiltonian on the supercell: two diagonal passes
    split = operator.build.kinetic_hamiltonian(operator.build.repeat_potential(potential, repeat), mass)
    got = operator.expectation_of_each(split, evolved).raw_data
    np.testing.assert_allclose(got, want, rtol=0, atol=1e-10 * abs(want[0]))
    # the diagonalised Bloch Hamiltonian is diagonal in the evolved list's own (eigen)basis: one pass, no transform
    from slate_quantum_b200.operator._operator import _diagonal_form

    diagonal = bloch.into_diagonal(hamiltonian)
    assert _diagonal_form(diagonal) is not None
    np.testing.assert_allclose(operator.expectation_of_each(diagonal, evolved).raw_data, want, rtol=0,
                               atol=1e-10 * abs(want[0]))
    v_part = operator.expectation_of_each(operator.build.repeat_potential(potential, repeat), evolved).raw_data
    t_part = operator.expectation_of_each(operator.build.kinetic_energy(full_meta, mass), evolved).raw_data
    np.testing.assert_allclose(v_part + t_part, want, rtol=0, atol=1e-10 * abs(want[0]))
    # in the position basis the kinetic term has a constant diagonal (its mean), the rest of diag(H) is V(x)
    v_x = np.diag(h_x).real - o.kinetic_energy(vectors * np.array(repeat)[:, None], big, mass).real.mean()
    np.testing.assert_allclose(v_part.real, (np.abs(pos) ** 2) @ v_x, rtol=0, atol=1e-10 * abs(want[0]))


def test_measure_from_function_and_apply_to_each(cuda):
    """operator.measure.{potential_from_function, all_potential_from_function, momentum_from_function}
    (reference operator/_measure.py:19-83) and operator.apply_to_each (operator/_operator.py:219-226), for diagonal
    and dense operators, against the oracle."""
    _, _, operator, state, TupleBasis, basis, metadata = _api()
    shape = (24, 18)
    dxm = np.diag([2 * np.pi, 5.0])
    meta = metadata.volume.spaced_volume_metadata_from_stacked_delta_x(tuple(dxm), shape)
    s = np.outer(o.coherent_state(dxm[:1, :1], (24,), (2.0,), (1.0,), (0.5,)),
                 o.coherent_state(dxm[1:, 1:], (18,), (1.7,), (-2 * np.pi / 5,), (0.6,))).ravel()
    states_np = np.array([s, 3.0j * np.roll(s, 5), s + 0.5 * np.roll(s, 40)])
    fund = basis.from_metadata(meta)
    states = state.StateList.from_states([state.State(fund.upcast(), row) for row in states_np])

    def f_x(x):
        return np.cos(x[0]) + 0.1 * x[1] ** 2

    def f_k(k):
        return k[0] ** 2 + 0.5 * k[1]

    np.testing.assert_allclose(operator.measure.all_potential_from_function(states, f_x).raw_data,
                               o.measure_all_potential_from_function(states_np, dxm, shape, f_x), atol=1e-12)
    np.testing.assert_allclose(
        operator.measure.all_potential_from_function(states, f_x, wrapped=True, offset=(0.3, -1.0)).raw_data,
        o.measure_all_potential_from_function(states_np, dxm, shape, f_x, (0.3, -1.0), True), atol=1e-12)
    np.testing.assert_allclose(operator.measure.all_momentum_from_function(states, f_k).raw_data,
                               o.measure_all_momentum_from_function(states_np, dxm, shape, f_k), atol=1e-10)
    np.testing.assert_allclose(operator.measure.potential_from_function(states[1], f_x),
                               o.measure_all_potential_from_function(states_np[1], dxm, shape, f_x)[0], atol=1e-12)
    np.testing.assert_allclose(operator.measure.momentum_from_function(states[2], f_k),
                               o.measure_all_momentum_from_function(states_np[2], dxm, shape, f_k)[0], atol=1e-10)
    # f = the coordinate itself reproduces the named observables
    np.testing.assert_allclose(operator.measure.all_potential_from_function(states, lambda x: x[1]).raw_data,
                               operator.measure.all_x(states, axis=1).raw_data, atol=1e-12)
    np.testing.assert_allclose(operator.measure.all_momentum_from_function(states, lambda k: k[0]).raw_data,
                               operator.measure.all_k(states, axis=0).raw_data, atol=1e-10)

    # apply_to_each: diagonal (column scaling) and dense (GEMM) forms agree with O @ psi
    v_op = operator.build.potential_from_function(meta, f_x)
    k_op = operator.build.momentum_from_function(meta, f_k)
    dense_basis = TupleBasis((fund, fund.dual_basis()))
    for op in (v_op, k_op):
        o_np = op.as_array()
        want = states_np @ o_np.T
        scale = np.abs(want).max()
        got = operator.apply_to_each(op, states).with_state_basis(fund.upcast()).raw_data.reshape(3, -1)
        np.testing.assert_allclose(got, want, rtol=0, atol=1e-12 * scale)
        got_dense = operator.apply_to_each(op.with_basis(dense_basis), states)
        np.testing.assert_allclose(got_dense.with_state_basis(fund.upcast()).raw_data.reshape(3, -1), want, rtol=0,
                                   atol=1e-12 * scale)
        np.testing.assert_allclose(operator.apply(op, states[0]).with_basis(fund.upcast()).raw_data, want[0], rtol=0,
                                   atol=1e-12 * scale)


# ---------------------------------------------------------------------------------------- round 2 additions
def test_build_square_operator(cuda):
    """operator.build.square_potential (reference operator/_build/_potential.py:134-164): with every Fourier term
    kept the potential is 0.5 * height * prod_a sign(cos(2 pi j_a / N_a)) on the grid (lanczos_factor = 0), and
    cropping to n_terms keeps the lowest |frequency| components of that square wave."""
    _, _, operator, _, _, _, metadata = _api()
    shape = (12, 10)
    meta = metadata.volume.spaced_volume_metadata_from_stacked_delta_x(
        (np.array([2 * np.pi, 0]), np.array([0, 2 * np.pi])), shape
    )
    wave = [np.sign(np.cos(np.linspace(0, 2 * np.pi, n, endpoint=False))) for n in shape]
    expected = 0.5 * 3.0 * np.multiply.outer(wave[0], wave[1]).ravel()
    actual = operator.build.square_potential(meta, 3.0)
    np.testing.assert_allclose(actual.as_array(), np.diag(expected), atol=1e-13)
    # cropped, with a Lanczos factor: the analytic truncated Fourier series
    n_terms = (5, 3)
    cropped = operator.build.square_potential(meta, 3.0, n_terms=n_terms, lanczos_factor=1.0)
    series = []
    for n_full, n_t in zip(shape, n_terms):
        coeff = np.sinc(2 * np.fft.fftfreq(n_t)) ** 1.0 * np.fft.ifft(
            np.sign(np.cos(np.linspace(0, 2 * np.pi, n_t, endpoint=False)))
        )
        freqs = np.fft.ifftshift(np.arange(-(n_t // 2), n_t - (n_t // 2)))
        j = np.arange(n_full)
        # V(x) = fft_ortho(V~) on the dual index: sum_k c_k exp(-2 pi i k j / N)
        series.append(sum(c * np.exp(-2j * np.pi * k * j / n_full) for c, k in zip(coeff, freqs)))
    expected = 0.5 * 3.0 * np.multiply.outer(series[0], series[1]).ravel()
    np.testing.assert_allclose(np.diag(cropped.as_array()), expected, atol=1e-13)


def test_diagonalised_operator_back_to_dense(cuda):
    """A diagonalised Hamiltonian converted back to a dense basis equals the Hamiltonian, and products with it
    equal products with the dense operator (the dual explicit basis conjugates MATERIALISED data)."""
    bloch, _, operator, _, TupleBasis, basis, metadata = _api()
    meta = metadata.volume.spaced_volume_metadata_from_stacked_delta_x((np.array([2 * np.pi, 0]), np.array([0, 2 * np.pi])), (4, 3))
    potential = operator.build.sin_potential(meta, 2.0, offset=0.3)
    hamiltonian = operator.build.kinetic_hamiltonian(potential, hbar**2)
    diagonal = operator.into_diagonal_hermitian(hamiltonian)
    for dense_basis in (
        basis.transformed_from_metadata(hamiltonian.basis.metadata(), is_dual=(False, True)).upcast(),
        basis.from_metadata(hamiltonian.basis.metadata(), is_dual=(False, True)).upcast(),
    ):
        want = hamiltonian.with_basis(dense_basis).raw_data
        got = diagonal.with_basis(dense_basis).raw_data
        np.testing.assert_allclose(got, want, atol=1e-12)
    np.testing.assert_allclose(diagonal.as_array(), hamiltonian.as_array(), atol=1e-12)
    x = operator.build.x(meta, axis=0)
    np.testing.assert_allclose(
        operator.matmul(diagonal, x).as_array(), operator.matmul(hamiltonian, x).as_array(), atol=1e-12
    )
    np.testing.assert_allclose(
        operator.commute(diagonal, x).as_array(), operator.commute(hamiltonian, x).as_array(), atol=1e-12
    )
    # the same through a Bloch (block diagonal) Hamiltonian
    sparse = bloch.build.kinetic_hamiltonian(potential, hbar**2, (3, 2))
    np.testing.assert_allclose(bloch.into_diagonal(sparse).as_array(), sparse.as_array(), atol=1e-12)


@pytest.mark.parametrize(("shape", "repeat"), [((5,), (6,)), ((4, 3), (3, 4)), ((3, 2, 2), (2, 3, 2))])
def test_bloch_basis_list_conversion_is_factorised_and_exact(cuda, shape, repeat):
    """BlochTransposedBasis.__into_fundamental__ / __from_fundamental__ on state lists (the chain of reference
    bloch/_transposed_basis.py:102-167 + _shifted_basis.py:62-111 + TransformedBasis): the factorised route
    (in-cell FFT + twiddle + across-block FFT) equals the oracle's permute + supercell FFT, kets and bras."""
    import torch

    bloch, _, _, _, _, basis, metadata = _api()
    from slate_quantum_b200.metadata import repeat_volume_metadata

    nd = len(shape)
    vectors = 2 * np.pi * np.eye(nd)
    meta = metadata.volume.spaced_volume_metadata_from_stacked_delta_x(tuple(vectors[i] for i in range(nd)), shape)
    super_meta = repeat_volume_metadata(meta, repeat)
    bloch_basis = bloch.build.basis_from_metadata(super_meta)
    big = tuple(n * r for n, r in zip(shape, repeat))
    rng = np.random.default_rng(5)
    m = int(np.prod(big))
    vecs = rng.normal(size=(7, m)) + 1j * rng.normal(size=(7, m))
    want = o.momentum_to_position(o.bloch_into_supercell(vecs, shape, repeat), big)
    got = bloch_basis.__into_fundamental__(vecs, axis=-1).ok().cpu().numpy()
    np.testing.assert_allclose(got, want, atol=1e-12)
    back = bloch_basis.__from_fundamental__(got, axis=-1).ok().cpu().numpy()
    np.testing.assert_allclose(back, vecs, atol=1e-12)
    # along axis 0 (vectors stored as columns) and on the dual basis (bra indices: the conjugate transform)
    got0 = bloch_basis.__into_fundamental__(vecs.T.copy(), axis=0).ok().cpu().numpy()
    np.testing.assert_allclose(got0, want.T, atol=1e-12)
    dual = bloch_basis.dual_basis()
    got_dual = dual.__into_fundamental__(vecs.conj(), axis=-1).ok().cpu().numpy()
    np.testing.assert_allclose(got_dual, want.conj(), atol=1e-12)
    back_dual = dual.__from_fundamental__(got_dual, axis=-1).ok().cpu().numpy()
    np.testing.assert_allclose(back_dual, vecs.conj(), atol=1e-12)
    del torch


def _bloch_problem(cfg_shape, cfg_repeat, height, n_times, t_max):
    """(hamiltonian, initial state, times basis, position basis, oracle pieces) of a square-lattice cos problem."""
    bloch, _, operator, state, _, basis, metadata = _api()
    from slate_quantum_b200.core import FundamentalBasis
    from slate_quantum_b200.core.metadata import Domain
    from slate_quantum_b200.metadata import SpacedTimeMetadata, repeat_volume_metadata

    nd = len(cfg_shape)
    vectors = 2 * np.pi * np.eye(nd)
    meta = metadata.volume.spaced_volume_metadata_from_stacked_delta_x(tuple(vectors[i] for i in range(nd)), cfg_shape)
    potential = operator.build.cos_potential(meta, height)
    hamiltonian = bloch.build.kinetic_hamiltonian(potential, hbar**2, cfg_repeat)
    super_meta = repeat_volume_metadata(meta, cfg_repeat)
    psi0 = state.build.coherent(super_meta, tuple(np.pi * r for r in cfg_repeat), (0.0,) * nd, (np.pi,) * nd)
    times = FundamentalBasis(SpacedTimeMetadata(n_times, domain=Domain(delta=t_max), interpolation="Fourier"))
    position_basis = basis.from_metadata(super_meta).upcast()
    return hamiltonian, psi0, times, position_basis, vectors


def test_lazy_tiled_evolution_access_paths(cuda):
    """EvolvedStateList.with_state_basis returns a lazy, time-tiled list: tiles(), host_tiles(), copy_to_host(),
    device_data, raw_data, indexing and iteration all give the oracle's states, for the position, momentum and
    Bloch bases; operator.measure on the lazy list equals operator.measure on the materialised one."""
    import torch

    _, dynamics, operator, _, _, basis, _ = _api()
    from slate_quantum_b200.dynamics.schrodinger import TiledEvolvedStates

    shape, repeat, n_t = (4, 4), (4, 3), 50
    hamiltonian, psi0, times, position_basis, vectors = _bloch_problem(shape, repeat, 4.0, n_t, 6 * hbar)
    evolution = dynamics.solve_schrodinger_equation_decomposition(psi0, times, hamiltonian)
    comps = o.cos_potential_components(shape, 4.0)
    w_ref, t_ref = o.eigh_blocks(o.bloch_blocks(vectors, shape, comps, hbar**2, repeat))
    t_values = np.asarray(times.metadata().values)
    want = o.evolve_pipeline(psi0.raw_data, shape, repeat, w_ref, t_ref, t_values, hbar)

    lazy = evolution.with_state_basis(position_basis)
    assert isinstance(lazy, TiledEvolvedStates) and lazy._kind == "position"  # noqa: SLF001
    assert len(lazy) == n_t and lazy.basis.size == n_t * want["position"].shape[1]
    got = np.concatenate([tile.cpu().numpy() for _, tile in lazy.tiles(rows=16)])  # ragged last tile
    np.testing.assert_allclose(got, want["position"], atol=1e-9)
    starts = [t0 for t0, _ in lazy.tiles(rows=16)]
    assert starts == [0, 16, 32, 48]
    host = np.concatenate([view.numpy().copy() for _, view in lazy.host_tiles(rows=16, chunk_bytes=7 * lazy.row_bytes)])
    np.testing.assert_allclose(host, want["position"], atol=1e-9)
    pageable = np.empty((n_t, lazy.state_size), dtype=np.complex128)
    lazy.copy_to_host(pageable, rows=20)
    np.testing.assert_allclose(pageable, want["position"], atol=1e-9)
    pinned = torch.empty((n_t, lazy.state_size), dtype=torch.complex128).pin_memory()
    lazy.copy_to_host(pinned, rows=20)
    np.testing.assert_allclose(pinned.numpy(), want["position"], atol=1e-9)
    np.testing.assert_allclose(lazy[7].raw_data, want["position"][7], atol=1e-9)
    np.testing.assert_allclose(lazy[n_t - 1, :].raw_data, want["position"][-1], atol=1e-9)
    # observables on the LAZY list (tile by tile) == on the materialised list == oracle; and the same through K6 fused
    # into the last cell-FFT pass (SQB_MEASURE_FUSED=1: the position-basis states are never written)
    import os

    os.environ["SQB_MEASURE_FUSED"] = "1"
    try:
        fused_evolution = dynamics.solve_schrodinger_equation_decomposition(psi0, times, hamiltonian)
        x_fused = operator.measure.all_x(fused_evolution, axis=0).raw_data
        w_fused = operator.measure.all_coherent_width(fused_evolution, axis=1).raw_data
    finally:
        os.environ["SQB_MEASURE_FUSED"] = "0"
    x_lazy = operator.measure.all_x(evolution, axis=0).raw_data
    np.testing.assert_allclose(x_fused, x_lazy, atol=1e-10)
    w_lazy = operator.measure.all_coherent_width(evolution, axis=1).raw_data
    big = tuple(n * r for n, r in zip(shape, repeat))
    super_dx = vectors * np.array(repeat)[:, None]
    np.testing.assert_allclose(x_lazy, o.measure_all_x(want["position"], super_dx, big, 0), atol=1e-9)
    np.testing.assert_allclose(w_lazy, o.measure_all_coherent_width(want["position"], super_dx, big, 1), atol=1e-8)
    np.testing.assert_allclose(w_fused, w_lazy, atol=1e-9)
    np.testing.assert_allclose(operator.measure.all_variance_x(evolution, axis=0).raw_data,
                               o.measure_all_variance_x(want["position"], super_dx, big, 0), atol=1e-8)
    np.testing.assert_allclose(operator.measure.all_periodic_x(evolution, axis=1).raw_data,
                               o.measure_all_periodic_x(want["position"], super_dx, big, 1), atol=1e-9)
    np.testing.assert_allclose(lazy.raw_data.reshape(n_t, -1), want["position"], atol=1e-9)  # materialises
    # a materialised list takes the plain read passes: same numbers as the fused (never written) route above
    np.testing.assert_allclose(operator.measure.all_x(lazy, axis=0).raw_data, x_lazy, atol=1e-10)
    np.testing.assert_allclose(operator.measure.all_coherent_width(lazy, axis=1).raw_data, w_lazy, atol=1e-9)
    # momentum and Bloch bases
    momentum_basis = basis.transformed_from_metadata(position_basis.metadata()).upcast()
    in_k = evolution.with_state_basis(momentum_basis)
    assert in_k._kind == "momentum"  # noqa: SLF001
    np.testing.assert_allclose(in_k.raw_data.reshape(n_t, -1), want["momentum"], atol=1e-9)
    bloch_basis = evolution._eigenbasis().inner  # noqa: SLF001  (the eigenbasis' inner = Bloch-ordered momentum basis)
    in_b = evolution.with_state_basis(bloch_basis)
    assert in_b._kind == "bloch"  # noqa: SLF001
    np.testing.assert_allclose(in_b.raw_data.reshape(n_t, -1), want["bloch"], atol=1e-9)
    states = list(in_b)
    assert len(states) == n_t
    np.testing.assert_allclose(states[3].raw_data, want["bloch"][3], atol=1e-9)
    # <H> through the diagonal-form expectation, tile by tile on the lazy list
    energy = operator.expectation_of_each(hamiltonian, evolution).raw_data
    assert np.abs(energy - energy[0]).max() < 1e-9 * abs(energy[0])


def test_cfg1_end_to_end_through_the_api(cuda):
    """BASELINE.json configs[0] (examples/blochs_theorem.py-style: 1-D cosine, 3 Fourier components, 100 k-points,
    64 states per block) through the drop-in API, against the oracle: blocks, every eigenvalue, degenerate-subspace
    projectors, evolved states in the eigen- and position bases."""
    _, dynamics, operator, _, _, _, _ = _api()
    from slate_quantum_b200 import bloch
    from slate_quantum_b200.configs import CONFIGS

    cfg = CONFIGS["cfg1"]
    hamiltonian, psi0, times, position_basis, vectors = _bloch_problem(cfg.shape, cfg.repeat, cfg.height, cfg.n_times,
                                                                       cfg.t_max_over_hbar * hbar)
    comps = cfg.potential_components()
    blocks = o.bloch_blocks(vectors, cfg.shape, comps, hbar**2, cfg.repeat)
    np.testing.assert_allclose(hamiltonian.raw_data.reshape(blocks.shape), blocks, atol=1e-13)
    diagonal = bloch.into_diagonal(hamiltonian)
    w_ref, t_ref = o.eigh_blocks(blocks)
    w = diagonal.raw_data.real.reshape(w_ref.shape)
    np.testing.assert_allclose(w, w_ref, rtol=0, atol=1e-10 * np.abs(w_ref).max())
    transform = diagonal.basis.inner.inner.children[0].inner.transform().raw_data.reshape(t_ref.shape)
    res, orth = o.eigh_residuals(blocks, w, transform)
    assert res < 1e-10 and orth < 1e-10
    for b in range(cfg.n_k):
        assert o.subspace_projector_error(w_ref[b], t_ref[b], transform[b], float(np.linalg.norm(blocks[b]))) < 1e-6
    evolution = dynamics.solve_schrodinger_equation_decomposition(psi0, times, hamiltonian)
    t_values = np.asarray(times.metadata().values)
    want = o.evolve_pipeline(psi0.raw_data, cfg.shape, cfg.repeat, w_ref, t_ref, t_values, hbar)
    # eigenvector phases (and rotations inside degenerate pairs) are a gauge choice: compare each block's weight
    got_w = np.linalg.norm(evolution.raw_data.reshape(cfg.n_times, cfg.n_k, cfg.n), axis=2)
    np.testing.assert_allclose(got_w, np.linalg.norm(want["eigen"].reshape(cfg.n_times, cfg.n_k, cfg.n), axis=2), atol=1e-9)
    in_position = evolution.with_state_basis(position_basis).raw_data.reshape(cfg.n_times, -1)
    np.testing.assert_allclose(in_position, want["position"], atol=1e-9)
    del operator


def test_cfg2_full_size_through_the_api_against_the_oracle(cuda):
    """BASELINE.json configs[1] at FULL size (1024 k-points, N = 128, 1024 times) through
    bloch.build.kinetic_hamiltonian -> solve_schrodinger_equation_decomposition -> with_state_basis(position):
    sampled k-blocks of the first, a middle and the last time row against the CPU oracle at 1e-9, eigenvalues of the
    sampled blocks at rel 1e-10."""
    _, dynamics, _, _, _, _, _ = _api()
    from slate_quantum_b200.configs import CONFIGS

    cfg = CONFIGS["cfg2"]
    hamiltonian, psi0, times, position_basis, vectors = _bloch_problem(cfg.shape, cfg.repeat, cfg.height, cfg.n_times,
                                                                       cfg.t_max_over_hbar * hbar)
    evolution = dynamics.solve_schrodinger_equation_decomposition(psi0, times, hamiltonian)
    lazy = evolution.with_state_basis(position_basis)
    rows_wanted = [0, cfg.n_times // 2 - 1, cfg.n_times - 1]
    kept = {}
    for t0, tile in lazy.tiles(rows=512):
        for r in rows_wanted:
            if t0 <= r < t0 + tile.shape[0]:
                kept[r] = tile[r - t0].cpu().numpy()
    t_values = np.asarray(times.metadata().values)
    np.testing.assert_allclose(t_values, cfg.times(), rtol=1e-15)
    blocks = [0, 341, 513, 1023]
    err = o.sampled_block_parity(
        vectors, cfg.shape, cfg.potential_components(), hbar**2, cfg.repeat, psi0.raw_data, t_values[rows_wanted],
        np.stack([kept[r] for r in rows_wanted]), blocks, hbar,
    )
    assert err < 1e-9, err
    w = evolution._eigenvalues.cpu().numpy()  # noqa: SLF001
    for b in blocks:
        w_ref = np.linalg.eigvalsh(o.bloch_blocks(vectors, cfg.shape, cfg.potential_components(), hbar**2, cfg.repeat,
                                                  block_begin=b, block_count=1)[0])
        np.testing.assert_allclose(w[b], w_ref, rtol=0, atol=1e-10 * np.abs(w_ref).max())


def test_dense_single_cell_eigenstates_above_512_states(cuda):
    """examples/eigenstates.py:16-35 on a single large cell (N = 1536 > the 512-state single-SM limit):
    operator.build.kinetic_hamiltonian + into_diagonal_hermitian against np.linalg.eigh at 1e-10, and a short
    eigen-decomposition evolution of a coherent state against the oracle's dense propagation."""
    _, dynamics, operator, state, _, basis, metadata = _api()
    from slate_quantum_b200.core import FundamentalBasis
    from slate_quantum_b200.core.metadata import Domain
    from slate_quantum_b200.metadata import SpacedTimeMetadata

    n = 1536
    meta = metadata.volume.spaced_volume_metadata_from_stacked_delta_x((np.array([2 * np.pi]),), (n,))
    potential = operator.build.cos_potential(meta, 40.0)
    hamiltonian = operator.build.kinetic_hamiltonian(potential, hbar**2)
    dense = o.dense_hamiltonian(2 * np.pi * np.eye(1), (n,), o.pad_components((n,), o.cos_potential_components((n,), 40.0)),
                                hbar**2, None)
    w_ref, v_ref = np.linalg.eigh(dense)
    diagonal = operator.into_diagonal_hermitian(hamiltonian)
    w = diagonal.raw_data.real
    np.testing.assert_allclose(w, w_ref, rtol=0, atol=1e-10 * np.abs(w_ref).max())
    transform = diagonal.basis.inner.inner.children[0].inner.transform().raw_data.reshape(n, n)
    res, orth = o.eigh_residuals(dense[None], w[None], transform[None])
    assert res < 1e-10 and orth < 1e-10
    initial = state.build.coherent(meta, (np.pi,), (0.0,), (0.4,))
    times = FundamentalBasis(SpacedTimeMetadata(5, domain=Domain(delta=2 * hbar), interpolation="Fourier"))
    evolution = dynamics.solve_schrodinger_equation_decomposition(initial, times, hamiltonian)
    momentum = basis.transformed_from_metadata(meta).upcast()
    got = evolution.with_state_basis(momentum).raw_data.reshape(5, n)
    psi_k = np.fft.fft(initial.raw_data, norm="ortho")
    c = v_ref.conj().T @ psi_k
    t_values = np.asarray(times.metadata().values)
    want = np.array([v_ref @ (c * np.exp(-1j * w_ref * t / hbar)) for t in t_values])
    np.testing.assert_allclose(got, want, atol=1e-9)


def test_noise_kernel_eigenvalue_diagonalisation(cuda):
    """noise._kernel on the GPU (K7 contractions + K2 eigensolve) against the oracle restatement of
    noise/_kernel.py:38-177 and noise/diagonalize/_eigenvalue.py:22-131: kernels at 1e-12, eigenvalues at rel
    1e-10, and the returned operators rebuild the kernel (gauge-free comparison of the eigenvectors)."""
    _, _, operator, _, TupleBasis, basis, metadata = _api()
    from slate_quantum_b200 import noise
    from slate_quantum_b200.core.basis import DiagonalBasis
    from slate_quantum_b200.metadata import eigenvalue_basis
    from slate_quantum_b200.operator import OperatorList

    rng = np.random.default_rng(12)
    n, n_ops = 12, 4
    meta = metadata.volume.spaced_volume_metadata_from_stacked_delta_x((np.array([2 * np.pi]),), (n,))
    values = np.array([0.4, 0.9, 1.6, 2.5])
    state_basis = basis.from_metadata(meta)
    # diagonal operators
    diag = rng.normal(size=(n_ops, n)) + 1j * rng.normal(size=(n_ops, n))
    diag_basis = DiagonalBasis(TupleBasis((state_basis, state_basis.dual_basis())))
    ops = OperatorList(TupleBasis((eigenvalue_basis(values), diag_basis)).upcast(), diag.ravel())
    kernel = noise.diagonal_kernel_from_operators(ops)
    want = o.diagonal_noise_kernel_from_operators(values, diag)
    np.testing.assert_allclose(kernel.raw_data.reshape(n, n), want, atol=1e-12)
    found = noise.get_periodic_noise_operators_diagonal_eigenvalue(kernel)
    w_ref, _ = o.noise_operators_diagonal_eigenvalue(want)
    w = np.asarray(found.basis.metadata().children[0].values)
    np.testing.assert_allclose(w, w_ref, rtol=0, atol=1e-10 * np.abs(w_ref).max())
    rebuilt = o.diagonal_noise_kernel_from_operators(w, found.raw_data.reshape(n, n))
    np.testing.assert_allclose(rebuilt, want, atol=1e-10)
    np.testing.assert_allclose(noise.diagonal_kernel_from_operators(found).raw_data.reshape(n, n), want, atol=1e-10)
    # full operators: n^2 = 144 x 144 kernel matrix
    full = rng.normal(size=(n_ops, n, n)) + 1j * rng.normal(size=(n_ops, n, n))
    dense_basis = basis.from_metadata(ops.basis.metadata().children[1], is_dual=(False, True))
    ops_full = OperatorList(TupleBasis((eigenvalue_basis(values), dense_basis)).upcast(), full.ravel())
    kernel_full = noise.noise_kernel_from_operators(ops_full)
    want_full = o.noise_kernel_from_operators(values, full)
    np.testing.assert_allclose(kernel_full.raw_data.reshape(n, n, n, n), want_full, atol=1e-12)
    found_full = noise.get_periodic_noise_operators_eigenvalue(kernel_full)
    w_full = np.asarray(found_full.basis.metadata().children[0].values)
    w_full_ref, _ = o.noise_operators_eigenvalue(want_full)
    np.testing.assert_allclose(w_full, w_full_ref, rtol=0, atol=1e-10 * np.abs(w_full_ref).max())
    rebuilt_full = o.noise_kernel_from_operators(w_full, found_full.raw_data.reshape(n * n, n, n))
    np.testing.assert_allclose(rebuilt_full, want_full, atol=1e-9)
    del operator


def test_measure_on_a_non_orthogonal_cell(cuda):
    """operator.measure.{all_x, all_periodic_x, all_variance_x, all_coherent_width} on the hexagonal cell of
    reference tests/test_operator.py:324-330 (cfg4's lattice): the Cartesian x mixes both lattice axes; against the
    oracle restatement of operator/_measure.py:105-305 with slate_core's stacked x points."""
    _, _, operator, state, _, basis, metadata = _api()
    from slate_quantum_b200.state import StateList

    vectors = 2 * np.pi * np.array([[np.sin(np.pi / 3), np.cos(np.pi / 3)], [0.0, 1.0]])
    shape = (12, 10)
    meta = metadata.volume.spaced_volume_metadata_from_stacked_delta_x((vectors[0], vectors[1]), shape)
    rng = np.random.default_rng(4)
    states_np = rng.normal(size=(6, 120)) + 1j * rng.normal(size=(6, 120))
    states_np[0] = o.coherent_state(vectors, shape, (2.0, 3.5), (0.0, 0.0), (0.8, 0.8))
    list_basis = basis.from_metadata(metadata.SimpleMetadata(6)) if hasattr(metadata, "SimpleMetadata") else None
    from slate_quantum_b200.core import FundamentalBasis, TupleBasis
    from slate_quantum_b200.core.metadata import SimpleMetadata

    del list_basis
    states = StateList(TupleBasis((FundamentalBasis(SimpleMetadata(6)), basis.from_metadata(meta))).upcast(), states_np.ravel())
    for axis in (0, 1):
        np.testing.assert_allclose(operator.measure.all_x(states, axis=axis).raw_data,
                                   o.measure_all_x(states_np, vectors, shape, axis), atol=1e-10)
        np.testing.assert_allclose(operator.measure.all_x(states, axis=axis, offset=0.7, wrapped=True).raw_data,
                                   o.measure_all_x(states_np, vectors, shape, axis, 0.7, True), atol=1e-10)
        np.testing.assert_allclose(operator.measure.all_periodic_x(states, axis=axis).raw_data,
                                   o.measure_all_periodic_x(states_np, vectors, shape, axis), atol=1e-10)
        np.testing.assert_allclose(operator.measure.all_variance_x(states, axis=axis).raw_data,
                                   o.measure_all_variance_x(states_np, vectors, shape, axis), atol=1e-9)
        np.testing.assert_allclose(operator.measure.all_coherent_width(states, axis=axis).raw_data,
                                   o.measure_all_coherent_width(states_np, vectors, shape, axis), atol=1e-9)
    del state
