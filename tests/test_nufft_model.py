"""CPU checks of the numpy model of the uniform-grid NUFFT evolve kernel (tests/nufft_model.py) against the direct sum
the reference evaluates (dynamics/schrodinger.py:51-53 + the eigenvector back-transformation)."""
from __future__ import annotations

import numpy as np
import pytest

import nufft_model as nm

HBAR = 1.0545718e-34


def _direct(w, c, v, t0, dt, t_count):
    t = t0 + dt * np.arange(t_count)
    return (c * np.exp(-1j * np.outer(t, w) / HBAR)) @ v


@pytest.mark.parametrize(("n", "t_count", "scale"), [(40, 300, 5.0), (64, 700, 40.0), (17, 512, 900.0)])
def test_model_matches_direct_sum(n, t_count, scale):
    rng = np.random.default_rng(n)
    w = rng.uniform(-0.1 * scale, scale, n) * HBAR * 1e12
    w[3] = w[4] = w[5]  # a degenerate cluster: all three sources share one window
    c = (rng.normal(size=n) + 1j * rng.normal(size=n)) / np.sqrt(n)
    v = (rng.normal(size=(n, 8)) + 1j * rng.normal(size=(n, 8))) / np.sqrt(n)
    t0, dt = 0.3e-12, 0.37e-12
    got = nm.evolve_uniform(w, c, v, t0, dt, t_count, HBAR)
    want = _direct(w, c, v, t0, dt, t_count)
    # aliasing error of the W = 14 window: ~1.2e-12 per unit source amplitude, plus the rounding of the phases themselves
    amp = (np.abs(c)[:, None] * np.abs(v)).sum(axis=0).max()
    assert np.abs(got - want).max() < 3e-12 * amp * max(1.0, scale / 40.0)


@pytest.mark.parametrize(("t_count", "want_k"), [(2048, 4), (1024, 2), (1536, 1), (4096 + 100, 1), (4096, 8)])
def test_interleaved_tiles_spread_a_narrow_band(t_count, want_k):
    """A time step that is small against the spectrum's bandwidth (the physical case): every source falls on a few grid
    points unless a tile takes every K-th row of a super-tile.  Ragged and odd tile counts fall back to smaller K."""
    rng = np.random.default_rng(t_count)
    n = 24
    dt = 1e-12
    w = rng.uniform(-0.05, 0.4, n) * HBAR / dt  # angles within 0.45 rad, as at cfg3
    c = (rng.normal(size=n) + 1j * rng.normal(size=n)) / np.sqrt(n)
    v = (rng.normal(size=(n, 3)) + 1j * rng.normal(size=(n, 3))) / np.sqrt(n)
    assert nm.interleave(w, dt, HBAR, -(-t_count // nm.TT)) == want_k
    got = nm.evolve_uniform(w, c, v, 0.2e-12, dt, t_count, HBAR)
    amp = (np.abs(c)[:, None] * np.abs(v)).sum(axis=0).max()
    assert np.abs(got - _direct(w, c, v, 0.2e-12, dt, t_count)).max() < 3e-12 * amp


def test_sources_at_the_grid_seam_wrap_around():
    n = 6
    theta = np.array([1e-9, 2 * np.pi - 1e-9, 0.0, np.pi, 2 * np.pi * 3.5 / nm.NG, 2 * np.pi * (nm.NG - 6.5) / nm.NG])
    dt = 1e-12
    w = theta * HBAR / dt
    c = np.ones(n, dtype=complex)
    v = np.eye(n, 8, dtype=complex) + 0.5j
    got = nm.evolve_uniform(w, c, v, 0.0, dt, 512, HBAR)
    assert np.abs(got - _direct(w, c, v, 0.0, dt, 512)).max() < 3e-12 * n


def test_window_transform_by_the_midpoint_rule_matches_gauss_legendre():
    width, beta = 14, 2.3 * 14
    x, wq = np.polynomial.legendre.leggauss(300)
    mp = np.arange(nm.TT) - nm.TT // 2
    ft = (width / 2) * np.array(
        [(wq * nm.es_window(x, beta) * np.cos(m * (2 * np.pi / nm.NG) * width / 2 * x)).sum() for m in mp]
    )
    assert np.abs(nm.deconvolution(width, beta) * ft - 1).max() < 1e-13


def test_dif_stages_are_a_1024_point_dft():
    rng = np.random.default_rng(2)
    g = rng.normal(size=(nm.NG, 2)) + 1j * rng.normal(size=(nm.NG, 2))
    full = np.fft.fft(g, axis=0)
    mp = np.arange(nm.TT) - nm.TT // 2
    assert np.abs(nm.fft_dif(g) - full[mp % nm.NG]).max() < 1e-11
