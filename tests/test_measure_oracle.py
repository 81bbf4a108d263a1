"""The oracle's observables against the invariants of the reference's own tests/test_measure.py (CPU)."""
from __future__ import annotations

import numpy as np

from oracle import bloch_oracle as o

DX = np.array([[2 * np.pi]])


def test_measure_x_position_state():
    """reference tests/test_measure.py:8-15: x of the i-th position eigenstate is exactly i dx."""
    n = 50
    dx = 2 * np.pi / n
    states = np.eye(n, dtype=np.complex128)
    x = o.measure_all_x(states, DX, (n,), 0)
    np.testing.assert_array_equal(x, np.arange(n) * dx)


def test_measure_x_coherent_state():
    """reference tests/test_measure.py:18-50 (x, periodic x with its sign, wrapped x around the centre)."""
    n = 300
    dx = 2 * np.pi / n
    sigma_0 = (0.1 * np.pi,)
    centre = o.coherent_state(DX, (n,), (150 * dx,), (0,), sigma_0)
    np.testing.assert_allclose(o.measure_all_x(centre, DX, (n,), 0)[0], 150 * dx, rtol=1e-3)
    states = np.array([o.coherent_state(DX, (n,), (i * dx,), (0,), sigma_0) for i in range(n)])
    expected = np.arange(n) * dx
    periodic = o.measure_all_periodic_x(states, DX, (n,), 0)
    difference = (periodic - expected + np.pi) % (2 * np.pi) - np.pi
    np.testing.assert_allclose(difference, 0, atol=1e-7 * dx)
    for i in range(0, n, 7):
        wrapped = o.measure_all_x(states[i], DX, (n,), 0, offset=-expected[i], wrapped=True)
        np.testing.assert_allclose(wrapped, 0, atol=1e-7 * dx)


def test_measure_width_coherent_state():
    """reference tests/test_measure.py:53-66 and :69-86."""
    n = 300
    dx = 2 * np.pi / n
    for sigma_0 in (0.1 * np.pi, 0.01 * np.pi):
        s = o.coherent_state(DX, (n,), (150 * dx,), (0,), (sigma_0,))
        width = o.measure_all_coherent_width(s, DX, (n,), 0)
        np.testing.assert_allclose(width, sigma_0, atol=1e-7 * dx)
    s = o.coherent_state(DX, (n,), (150 * dx,), (0,), (0.1 * np.pi,))
    np.testing.assert_allclose(o.measure_all_variance_x(s, DX, (n,), 0), (0.1 * np.pi) ** 2 / 2, rtol=1e-3)


def test_measure_two_axes_are_independent():
    """2-D product state: each axis' observables equal the 1-D ones of its factor."""
    shape = (24, 36)
    dxm = np.diag([2 * np.pi, 3.0])
    s0 = o.coherent_state(dxm[:1, :1], (shape[0],), (1.3,), (0.7,), (0.4,))
    s1 = o.coherent_state(dxm[1:, 1:], (shape[1],), (2.1,), (0.0,), (0.3,))
    prod = np.outer(s0, s1).ravel()
    np.testing.assert_allclose(o.measure_all_periodic_x(prod, dxm, shape, 0),
                               o.measure_all_periodic_x(s0, dxm[:1, :1], shape[:1], 0), atol=1e-12)
    np.testing.assert_allclose(o.measure_all_variance_x(prod, dxm, shape, 1),
                               o.measure_all_variance_x(s1, dxm[1:, 1:], shape[1:], 0), atol=1e-12)


def test_measure_k_and_uncertainty_of_a_gaussian():
    """Momentum observables (reference operator/_measure.py:307-420) on a wavepacket with known answers:
    <k> = k0, var_k = 1 / (2 sigma^2), sigma_x sigma_k = 1/2 (minimum-uncertainty state)."""
    n = 400
    length = 40.0
    dxm = np.array([[length]])
    sigma, k0 = 1.3, 2 * np.pi * 12 / length
    s = o.coherent_state(dxm, (n,), (17.0,), (k0,), (sigma,))
    np.testing.assert_allclose(o.measure_all_k(s, dxm, (n,), 0), k0, atol=1e-9)
    np.testing.assert_allclose(o.measure_all_variance_k(s, dxm, (n,), 0), 1 / (2 * sigma**2), rtol=1e-8)
    np.testing.assert_allclose(o.measure_all_uncertainty(s, dxm, (n,), 0), 0.5, rtol=1e-8)
    # the k points are the FFT-ordered literal of tests/test_operator.py:184-196
    np.testing.assert_allclose(o.stacked_k_points(np.array([[2 * np.pi]]), (5,))[0], [0, 1, 2, -2, -1])
