"""CPU: the oracle's operator / state algebra (K7) pinned by the reference's own tests
(tests/test_operator.py:180-260, :455-540; tests/test_state.py:20-75)."""

from __future__ import annotations

import numpy as np

from oracle import bloch_oracle as o

DX = np.array([[2 * np.pi]])
HBAR = o.HBAR_SI


def test_k_operator_literal():
    """tests/test_operator.py:184-210: diag([0,1,2,-2,-1]) in the momentum basis, ifft on kets / fft on bras."""
    want = np.fft.fft(np.fft.ifft(np.diag([0, 1, 2, -2, -1]), norm="ortho", axis=0), norm="ortho", axis=1)
    np.testing.assert_allclose(o.k_operator_array(DX, (5,), 0), want, atol=1e-15)
    assert np.allclose(o.k_operator_array(DX, (5,), 0), o.operator_dagger(o.k_operator_array(DX, (5,), 0)))


def test_x_operator_literal():
    """tests/test_operator.py:170-181: diag(linspace(0, 2 pi, 5, endpoint=False))."""
    np.testing.assert_allclose(o.x_operator_array(DX, (5,), 0), np.diag(np.linspace(0, 2 * np.pi, 5, endpoint=False)))


def test_x_k_commutator():
    """tests/test_operator.py:213-245."""
    x, k = o.x_operator_array(DX, (5,), 0), o.k_operator_array(DX, (5,), 0)
    np.testing.assert_allclose(o.operator_matmul(x, k), x @ k, atol=1e-15)
    np.testing.assert_allclose(o.operator_commutator(x, k), x @ k - k @ x, atol=1e-15)
    # tests/test_operator.py:248-259: an operator commutes with itself, exactly
    np.testing.assert_array_equal(o.operator_commutator(x, x), np.zeros((5, 5)))
    np.testing.assert_array_equal(o.operator_commutator(k, k), np.zeros((5, 5)))


def test_canonical_commutator_on_a_smooth_state():
    """[x, k] psi = i psi for a wavepacket well inside the cell (the continuum identity the reference's
    tests/test_operator.py:471-476 quotes as [x, p] = i hbar)."""
    n, length = 300, 2 * np.pi * 10
    dxm = np.array([[length]])
    psi = o.coherent_state(dxm, (n,), (length / 2,), (0.0,), (1.5,))
    comm = o.operator_commutator(o.x_operator_array(dxm, (n,), 0), o.k_operator_array(dxm, (n,), 0))
    np.testing.assert_allclose(o.operator_apply(comm, psi), 1j * psi, atol=1e-9)
    np.testing.assert_allclose(o.expectation_of_each(comm, psi), [1j], atol=1e-9)


def test_commutator_list_and_dagger():
    rng = np.random.default_rng(0)
    a = rng.normal(size=(6, 6)) + 1j * rng.normal(size=(6, 6))
    ops = rng.normal(size=(3, 6, 6)) + 1j * rng.normal(size=(3, 6, 6))
    got = o.operator_commutator(a, ops)
    for m in range(3):
        np.testing.assert_allclose(got[m], a @ ops[m] - ops[m] @ a, atol=1e-14)
    np.testing.assert_allclose(o.operator_dagger(ops)[1], ops[1].conj().T)
    # (AB)^dagger = B^dagger A^dagger
    np.testing.assert_allclose(o.operator_dagger(o.operator_matmul(a, ops[0])),
                               o.operator_matmul(o.operator_dagger(ops[0]), o.operator_dagger(a)), atol=1e-14)


def test_occupations_literal():
    """tests/test_state.py:50-74: two amplitudes (1-1j)/sqrt(2) -> occupations [1, 1, 0, 0, 0]."""
    data = np.zeros(5, dtype=np.complex128)
    data[:2] = (1 - 1j) / np.sqrt(2)
    np.testing.assert_allclose(o.occupations(data), [1, 1, 0, 0, 0], atol=1e-15)


def test_inner_products_literal():
    """tests/test_state.py:20-47: orthogonal and parallel position states."""
    s0 = np.zeros(5, dtype=np.complex128)
    s0[:2] = (1 - 1j) / np.sqrt(2)
    s1 = np.zeros(5, dtype=np.complex128)
    s1[0], s1[1] = (1 - 1j) / np.sqrt(2), -(1 - 1j) / np.sqrt(2)
    both = np.array([s0, s1])
    np.testing.assert_allclose(o.inner_product_each(both, both), [2, 2], atol=1e-15)
    np.testing.assert_allclose(o.all_inner_product(both, both), [[2, 0], [0, 2]], atol=1e-15)
    normed = o.normalize_each(both)
    np.testing.assert_allclose(o.inner_product_each(normed, normed), [1, 1], atol=1e-15)


def test_expectation_matches_measure():
    """<x> through the dense operator equals the K6 observable on a normalised state."""
    n, length = 200, 30.0
    dxm = np.array([[length]])
    psi = o.coherent_state(dxm, (n,), (11.0,), (0.8,), (1.1,))
    psi = o.normalize_each(psi)[0]
    ex = o.expectation_of_each(o.x_operator_array(dxm, (n,), 0), psi)
    np.testing.assert_allclose(ex.real, o.measure_all_x(psi, dxm, (n,), 0), atol=1e-12)
    ek = o.expectation_of_each(o.k_operator_array(dxm, (n,), 0), psi)
    np.testing.assert_allclose(ek.real, o.measure_all_k(psi, dxm, (n,), 0), atol=1e-10)


def test_temperature_corrected_operators_and_shift():
    """noise/build/_build.py:141-172 on dense arrays: hermitian L and H give an anti-hermitian correction term,
    and the shift vanishes when sum L^dagger L commutes with H (e.g. a single unitary L)."""
    rng = np.random.default_rng(1)
    n = 7
    g = rng.normal(size=(n, n)) + 1j * rng.normal(size=(n, n))
    h = g + g.conj().T
    ops = np.array([np.diag(rng.normal(size=n)).astype(complex), o.k_operator_array(DX, (n,), 0)])
    t, eta = 150.0, 0.3
    got = o.temperature_corrected_operators(h, ops, t, eta, hbar=1.0)
    kb_t = o.BOLTZMANN_SI * t
    for m in range(2):
        want = np.sqrt(2 * eta * kb_t) * ops[m] - np.sqrt(eta / (8 * kb_t)) * (h @ ops[m] - ops[m] @ h)
        np.testing.assert_allclose(got[m], want, rtol=1e-13)
    q, _ = np.linalg.qr(g)
    np.testing.assert_allclose(o.hamiltonian_shift(h, q[np.newaxis], eta, hbar=1.0), np.zeros((n, n)), atol=1e-13)
    shift = o.hamiltonian_shift(h, ops, eta, hbar=1.0)
    np.testing.assert_allclose(shift, o.operator_dagger(shift), atol=1e-12)  # i [H, hermitian] is hermitian


def test_measure_from_function_reduces_to_named_observables():
    """f(x) = x_axis and f(k) = k_axis reproduce all_x / all_k (pinned in tests/test_measure_oracle.py), f = x^2 the
    second moment; an unnormalised state gives the same values (the reference normalises first)."""
    shape = (24, 18)
    dxm = np.diag([2 * np.pi, 5.0])
    s = np.outer(o.coherent_state(dxm[:1, :1], (24,), (2.0,), (1.0,), (0.5,)),
                 o.coherent_state(dxm[1:, 1:], (18,), (1.7,), (-2 * np.pi / 5, ), (0.6,))).ravel()
    states = np.array([s, 3.0j * np.roll(s, 5)])
    for axis in range(2):
        np.testing.assert_allclose(o.measure_all_potential_from_function(states, dxm, shape, lambda x, a=axis: x[a]),
                                   o.measure_all_x(states, dxm, shape, axis), atol=1e-12)
        np.testing.assert_allclose(o.measure_all_momentum_from_function(states, dxm, shape, lambda k, a=axis: k[a]),
                                   o.measure_all_k(states, dxm, shape, axis), atol=1e-12)
    np.testing.assert_allclose(
        o.measure_all_potential_from_function(states, dxm, shape, lambda x: x[0] ** 2 + 0 * x[1]),
        o._normalised_density(states) @ (o.stacked_x_points(dxm, shape)[0] ** 2), atol=1e-12)


def test_oracle_against_the_golden_literals_file():
    """tests/golden/reference_literals.json carries the K6 / K7 literals transcribed from the reference's tests."""
    import json
    from pathlib import Path

    lit = json.loads((Path(__file__).parent / "golden" / "reference_literals.json").read_text())
    np.testing.assert_allclose(np.diag(o.x_operator_array(DX, (5,), 0)).real,
                               lit["x_operator_diagonal_2pi_cell_5_points"]["value"], atol=1e-15)
    np.testing.assert_allclose(o.stacked_k_points(DX, (5,))[0], lit["k_operator_fft_order_5"]["value"], atol=1e-15)
    occ = lit["occupations_two_equal_amplitudes"]
    amp = np.array(occ["amplitudes_re_im"]) @ np.array([1, 1j])
    np.testing.assert_allclose(o.occupations(amp), occ["value"], atol=1e-15)
    ip = lit["inner_product_orthogonal_position_states"]
    s0 = np.array(ip["state_0_re_im"]) @ np.array([1, 1j])
    s1 = np.array(ip["state_1_re_im"]) @ np.array([1, 1j])
    np.testing.assert_allclose(o.inner_product_each(s1[None], s0[None]), [ip["value"]], atol=1e-15)


def test_noise_kernel_eigenvalue_diagonalisation_round_trips():
    """Oracle restatement of noise/_kernel.py:38-177 and noise/diagonalize/_eigenvalue.py:22-131: the operators
    returned by the eigenvalue method rebuild the kernel they came from (the reference asserts exactly this
    reconstruction at :55-64 and :105-115)."""
    from oracle import bloch_oracle as o

    rng = np.random.default_rng(8)
    n, n_ops = 5, 3
    values = np.array([0.7, 1.3, 2.0])
    diag = rng.normal(size=(n_ops, n)) + 1j * rng.normal(size=(n_ops, n))
    beta = o.diagonal_noise_kernel_from_operators(values, diag)
    np.testing.assert_allclose(beta, beta.conj().T, atol=1e-14)
    w, ops = o.noise_operators_diagonal_eigenvalue(beta)
    assert np.all(w[:-n_ops] < 1e-12 * w[-1]) and np.all(w[-n_ops:] > 0)  # rank = number of independent operators
    np.testing.assert_allclose(o.diagonal_noise_kernel_from_operators(w, ops), beta, atol=1e-12)
    full_ops = rng.normal(size=(n_ops, n, n)) + 1j * rng.normal(size=(n_ops, n, n))
    kernel = o.noise_kernel_from_operators(values, full_ops)
    w2, ops2 = o.noise_operators_eigenvalue(kernel)
    np.testing.assert_allclose(o.noise_kernel_from_operators(w2, ops2), kernel, atol=1e-11)
    # a kernel built from ONE operator has one non-zero eigenvalue = lambda ||L||^2, operator parallel to L
    single = o.noise_kernel_from_operators(np.array([2.0]), full_ops[:1])
    w3, ops3 = o.noise_operators_eigenvalue(single)
    np.testing.assert_allclose(w3[-1], 2.0 * np.linalg.norm(full_ops[0]) ** 2, rtol=1e-12)
    overlap = abs(np.vdot(ops3[-1], full_ops[0])) / np.linalg.norm(full_ops[0])
    np.testing.assert_allclose(overlap, 1.0, atol=1e-12)
