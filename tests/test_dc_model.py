"""CPU checks of the numpy model of the divide-and-conquer tridiagonal eigensolver (tests/dc_model.py).

The model is the specification the CUDA kernels of slate_quantum_b200/csrc/eigh_dc.cuh are written against;
here it is pinned against numpy's LAPACK on the hard cases of the literature (Wilkinson, glued Wilkinson,
clustered, graded, zero couplings) and on tridiagonalised Bloch blocks.
"""
from __future__ import annotations

import numpy as np
import pytest

import dc_model as dc
import eigh_model as em
from oracle import bloch_oracle as o


def tridiagonal_cases():
    rng = np.random.default_rng(7)
    cases = {}
    for n in (1, 2, 3, 31, 32, 33, 65, 100, 129):
        cases[f"random{n}"] = (rng.normal(size=n), rng.normal(size=n))
    for n in (21, 65, 129):
        cases[f"wilkinson{n}"] = (np.abs(np.arange(n) - (n - 1) // 2).astype(float), np.ones(n))
    for glue in (1e-3, 1e-8, 1e-14):
        d, e = [], []
        for _ in range(6):
            d += list(np.abs(np.arange(21) - 10.0))
            e += [1.0] * 20 + [glue]
        cases[f"glued{glue}"] = (np.array(d), np.array(e))
    cases["equal_tiny_e"] = (np.ones(100), np.full(100, 1e-10))
    cases["toeplitz"] = (np.full(128, 2.0), np.full(128, -1.0))
    cases["zeros"] = (np.zeros(70), np.zeros(70))
    cases["identity"] = (np.ones(70), np.zeros(70))
    e = rng.normal(size=130)
    e[::7] = 0.0
    cases["zero_couplings"] = (rng.normal(size=130), e)
    cases["graded"] = (10.0 ** (-np.arange(96) / 8.0), 0.3 * 10.0 ** (-np.arange(96) / 8.0))
    cases["huge"] = (1e150 * rng.normal(size=80), 1e150 * rng.normal(size=80))
    cases["tiny"] = (1e-150 * rng.normal(size=80), 1e-150 * rng.normal(size=80))
    return cases


CASES = tridiagonal_cases()


def check_tridiagonal(d, e, w, q, tol_scale=1.0):
    n = len(d)
    t = np.diag(d) + np.diag(e[: n - 1], 1) + np.diag(e[: n - 1], -1)
    norm = max(np.linalg.norm(t), np.finfo(float).tiny)
    w_ref = np.linalg.eigvalsh(t)
    assert np.all(np.diff(w) >= 0)
    assert np.abs(w - w_ref).max() <= 1e-13 * tol_scale * max(np.abs(t).max(), np.finfo(float).tiny) * max(1, n / 32)
    assert np.linalg.norm(t @ q.T - q.T * w[None, :]) / norm < 1e-13 * tol_scale
    assert np.linalg.norm(q @ q.T - np.eye(n)) < 1e-12 * tol_scale * max(1, n / 64)


@pytest.mark.parametrize("name", sorted(CASES))
def test_dc_model_hard_cases(name):
    d, e = CASES[name]
    w, q = dc.dc_eigh_tridiagonal(d, e)
    check_tridiagonal(d, e, w, q)


def test_dc_model_tree_covers_every_row():
    for n in (1, 31, 32, 33, 100, 256, 343, 512):
        depth = dc.tree_depth(n)
        bounds = [dc.node_start(n, depth, i) for i in range((1 << depth) + 1)]
        assert bounds[0] == 0 and bounds[-1] == n
        sizes = np.diff(bounds)
        assert sizes.min() >= 1 and sizes.max() <= dc.LEAF
        for level in range(depth):
            for i in range(1 << level):  # children tile their parent
                assert dc.node_start(n, level, i) == dc.node_start(n, level + 1, 2 * i)


def test_dc_model_on_bloch_block():
    hbar = o.HBAR_SI
    shape, repeat = (8, 8), (4, 4)
    comps = o.cos_potential_components(shape, 10.0)
    for b in (0, 5):  # b = 0: the Gamma point, exactly degenerate levels
        h = o.bloch_blocks(2 * np.pi * np.eye(2), shape, comps, hbar**2, repeat, block_begin=b, block_count=1)[0]
        d, e, _, _ = em.tridiagonalise(h)
        w, q = dc.dc_eigh_tridiagonal(d, e)
        check_tridiagonal(d, e, w, q)
        assert np.abs(w - np.linalg.eigvalsh(h)).max() < 1e-12 * np.abs(h).max()


def test_secular_roots_interlace():
    rng = np.random.default_rng(3)
    k = 40
    dk = np.sort(rng.normal(size=k))
    z = rng.normal(size=k)
    z /= np.linalg.norm(z)
    rho = 0.7
    lam = []
    for j in range(k):
        origin, tau, its = dc.secular_root(j, dk, z * z, rho)
        assert its < dc.MAX_SECULAR_ITER
        lam.append(dk[origin] + tau)
    lam = np.array(lam)
    assert np.all(lam[:-1] > dk[:-1]) and np.all(lam[:-1] < dk[1:]) and lam[-1] > dk[-1]
    ref = np.linalg.eigvalsh(np.diag(dk) + rho * np.outer(z, z))
    assert np.abs(lam - ref).max() < 1e-13


def test_blocked_back_transform_model_equals_reflector_by_reflector():
    """The compact-WY blocking of csrc/eigh_wy.cuh reproduces the reflector-by-reflector back-transformation."""
    rng = np.random.default_rng(12)
    for n in (5, 33, 100, 130):
        g = rng.normal(size=(n, n)) + 1j * rng.normal(size=(n, n))
        d, e, tau, y = em.tridiagonalise(g + g.conj().T)
        z = rng.normal(size=(n, n))
        np.testing.assert_allclose(em.back_transform_wy(z, tau, y), em.back_transform(z, tau, y), atol=1e-12)


@pytest.mark.parametrize(("n", "nb"), [(5, 2), (33, 8), (64, 16), (130, 16), (96, 32), (40, 64)])
def test_blocked_tridiagonalisation_model_equals_unblocked(n, nb):
    """zlatrd-style panels (trailing matrix read-only inside a panel, one rank-2nb update per panel) give the same
    tridiagonal, reflectors and tau as the deferred rank-2 version the GPU kernel runs today."""
    from eigh_model import tridiagonalise, tridiagonalise_blocked

    rng = np.random.default_rng(n * 100 + nb)
    g = rng.normal(size=(n, n)) + 1j * rng.normal(size=(n, n))
    h = g + g.conj().T
    d0, e0, tau0, y0 = tridiagonalise(h)
    d1, e1, tau1, y1, (reads, writes) = tridiagonalise_blocked(h, nb)
    scale = np.abs(h).max() * n
    np.testing.assert_allclose(d1, d0, rtol=0, atol=1e-13 * scale)
    np.testing.assert_allclose(e1, e0, rtol=0, atol=1e-13 * scale)
    np.testing.assert_allclose(tau1, tau0, rtol=0, atol=1e-12)
    np.testing.assert_allclose(y1, y0, rtol=0, atol=1e-11)
    # the tridiagonal has the spectrum of H
    t = np.diag(d1) + np.diag(e1[: n - 1], 1) + np.diag(e1[: n - 1], -1)
    np.testing.assert_allclose(np.linalg.eigvalsh(t), np.linalg.eigvalsh(h), rtol=0, atol=1e-12 * scale)
    # one read sweep per column, one read+write sweep per PANEL (the unblocked kernel writes once per column)
    assert reads == n - 1 and writes == -(-(n - 1) // nb)
