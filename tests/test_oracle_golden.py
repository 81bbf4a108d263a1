"""CPU: the oracle against the reference's literal vectors, its test invariants and the frozen fixtures."""

import itertools
import json
from pathlib import Path

import numpy as np
import pytest
import scipy.linalg

from oracle import bloch_oracle as o

HB = o.HBAR_SI
GOLD = Path(__file__).parent / "golden"
LIT = json.loads((GOLD / "reference_literals.json").read_text())
SHAPES_1D = [tuple(s) for s in LIT["bloch_test_grid"]["shapes_1d"]]
SHAPES_2D = [tuple(s) for s in LIT["bloch_test_grid"]["shapes_2d"]]
GRID = [*itertools.product(SHAPES_1D, SHAPES_1D), *itertools.product(SHAPES_2D, SHAPES_2D)]


def test_kinetic_literal():
    got = o.kinetic_energy(np.array([[2 * np.pi]]), (5,), HB**2)
    np.testing.assert_allclose(got, LIT["kinetic_energy_2pi_cell_5_points_mass_hbar2"]["value"])
    assert got.dtype == np.complex128 and np.all(got.imag == 0)


def test_split_raw_layout_and_dense_diagonal():
    comps = o.cos_potential_components((5,), 0.0)
    kin = o.kinetic_energy(np.array([[2 * np.pi]]), (5,), HB**2)
    np.testing.assert_allclose(o.split_raw_data(comps, kin), LIT["split_hamiltonian_raw_cos0"]["value"])
    dense = o.dense_hamiltonian(np.array([[2 * np.pi]]), (5,), o.pad_components((5,), comps), HB**2)
    np.testing.assert_allclose(dense, np.diag(LIT["split_hamiltonian_dense_diagonal"]["value"]), atol=1e-15)


def test_fft_order_and_fractions():
    np.testing.assert_array_equal(o.nk_points(5), LIT["k_operator_fft_order_5"]["value"])
    np.testing.assert_array_equal(o.sample_fractions((4,))[0], LIT["bloch_fraction_values_R4"]["value"])


def test_fft_direction_convention():
    """tests/test_operator.py:197-204: as_array == fft(ifft(diag, axis=0, ortho), axis=1, ortho)."""
    diag = np.diag(np.array(LIT["k_operator_fft_order_5"]["value"], dtype=complex))
    want = np.fft.fft(np.fft.ifft(diag, norm="ortho", axis=0), norm="ortho", axis=1)
    np.testing.assert_allclose(o.momentum_matrix_to_position(diag, (5,)), want, atol=1e-15)


def test_eigenvalue_multiset_through_eigh():
    """tests/test_operator.py:62-75."""
    dx = np.array([[2 * np.pi]])
    kin = o.kinetic_energy(dx, (5,), HB**2).real
    dense = o.dense_hamiltonian(dx, (5,), o.pad_components((5,), o.cos_potential_components((5,), 0.0)), HB**2)
    w, _ = o.eigh_blocks(dense[None])
    np.testing.assert_allclose(np.sort(w[0]), np.sort(kin), atol=1e-15)


def test_cos_sin_fcc_potentials_are_the_analytic_functions():
    """tests/test_operator.py:305-320 and :323-353 (sin pins the fft-vs-ifft direction of the diagonal)."""
    x = np.arange(16) * 2 * np.pi / 16
    np.testing.assert_allclose(o.potential_position_values((16,), o.cos_potential_components((16,), 2.0)), 1 + np.cos(x), atol=1e-14)
    np.testing.assert_allclose(o.potential_position_values((16,), o.sin_potential_components((16,), 2.0)), 1 + np.sin(x), atol=1e-14)
    dx = np.array([[np.sin(np.pi / 3), np.cos(np.pi / 3)], [0, 1.0]])
    f = np.array(np.meshgrid(np.arange(16) / 16, np.arange(16) / 16, indexing="ij")).reshape(2, -1)
    xy = np.einsum("ij,ik->jk", dx, f)
    eta = 2 * np.pi * 2 / np.sqrt(3)
    want = 1 / 3 + 2 / 9 * (
        np.cos(eta * xy[0])
        + np.cos(eta * np.cos(np.pi / 3) * xy[0] + eta * np.sin(np.pi / 3) * xy[1])
        + np.cos(-eta * np.cos(np.pi / 3) * xy[0] + eta * np.sin(np.pi / 3) * xy[1])
    )
    got = o.potential_position_values((16, 16), o.fcc_potential_components((16, 16), 1.0)).ravel()
    np.testing.assert_allclose(got, want, atol=1e-14)


@pytest.mark.parametrize(("shape", "repeat"), GRID)
@pytest.mark.parametrize("height", [1.0, 0.0])
def test_bloch_blocks_equal_supercell(shape, repeat, height):
    """tests/test_bloch.py:16-47 (height 1) and :50-86 (height 0): atol 1e-14 / 1e-15."""
    dx = 2 * np.pi * np.eye(len(shape))
    comps = o.cos_potential_components(shape, height)
    blocks = o.bloch_blocks(dx, shape, comps, HB**2, repeat)
    full = o.supercell_hamiltonian(dx, shape, comps, HB**2, repeat)
    emb = o.embed_blocks(blocks, shape, repeat)
    np.testing.assert_allclose(emb, full, atol=1e-14)
    big = tuple(n * r for n, r in zip(shape, repeat))
    np.testing.assert_allclose(o.momentum_matrix_to_position(emb, big), o.momentum_matrix_to_position(full, big), atol=1e-14)
    np.testing.assert_allclose(np.sort(o.eigh_blocks(blocks)[0].ravel()), np.linalg.eigvalsh(full), atol=1e-12)


@pytest.mark.parametrize(("shape", "repeat"), GRID)
def test_potential_only_blocks(shape, repeat):
    """tests/test_bloch.py:89-126: tiled potential list == repeat_potential."""
    table = o.pad_components(shape, o.cos_potential_components(shape, 1.0))
    conv = o.convolution_matrix(table)
    n_k = int(np.prod(repeat))
    emb = o.embed_blocks(np.tile(conv[None], (n_k, 1, 1)), shape, repeat)
    rep = o.convolution_matrix(o.repeat_components(shape, o.cos_potential_components(shape, 1.0), repeat))
    np.testing.assert_allclose(emb, rep, atol=1e-15)


@pytest.mark.parametrize(("shape", "repeat"), [*GRID, ((3, 4, 5), (2, 3, 4)), ((16,), (7,)), ((7, 7, 7), (2, 2, 2))])
def test_permutation_closed_form_matches_literal(shape, repeat):
    m = int(np.prod(shape) * np.prod(repeat))
    v = np.arange(m)
    lit = o.bloch_from_supercell(v, shape, repeat)
    np.testing.assert_array_equal(lit, o.bloch_permutation(shape, repeat))
    np.testing.assert_array_equal(o.bloch_into_supercell(lit, shape, repeat), v)


def test_closed_form_convolution_equals_fft_densification():
    rng = np.random.default_rng(3)
    table = rng.normal(size=(4, 5)) + 1j * rng.normal(size=(4, 5))
    np.testing.assert_allclose(o.convolution_matrix(table), o.convolution_matrix_via_fft(table), atol=1e-14)


@pytest.mark.parametrize(("shape", "repeat"), [((5,), (4,)), ((4, 3), (3, 4)), ((3, 4, 3), (2, 2, 2))])
def test_evolution_against_dense_expm(shape, repeat):
    """dynamics/schrodinger.py has no reference test: cross-check the oracle with scipy expm."""
    dx = 2 * np.pi * np.eye(len(shape))
    comps = o.sin_potential_components(shape, 3.0, 0.3)
    blocks = o.bloch_blocks(dx, shape, comps, HB**2, repeat)
    w, t = o.eigh_blocks(blocks)
    big = tuple(n * r for n, r in zip(shape, repeat))
    m = int(np.prod(big))
    rng = np.random.default_rng(0)
    psi = rng.normal(size=m) + 1j * rng.normal(size=m)
    psi /= np.linalg.norm(psi)
    times = np.linspace(0, 3 * HB, 6)
    res = o.evolve_pipeline(psi, shape, repeat, w, t, times, HB)
    hx = o.momentum_matrix_to_position(o.supercell_hamiltonian(dx, shape, comps, HB**2, repeat), big)
    ref = np.array([scipy.linalg.expm(-1j * hx * tt / HB) @ psi for tt in times])
    np.testing.assert_allclose(res["position"], ref, atol=1e-12)


def test_frozen_fixtures():
    g = np.load(GOLD / "bloch_golden.npz")
    for name in ("sq1d", "sq2d", "hex2d", "cub3d"):
        shape, repeat = tuple(g[f"{name}/shape"]), tuple(g[f"{name}/repeat"])
        blocks = o.bloch_blocks(g[f"{name}/dx"], shape, g[f"{name}/comps"], HB**2, repeat)
        np.testing.assert_allclose(blocks, g[f"{name}/blocks"], atol=1e-14)
        w, t = o.eigh_blocks(blocks)
        np.testing.assert_allclose(w, g[f"{name}/eigenvalues"], atol=1e-12)
        res = o.evolve_pipeline(g[f"{name}/psi0"], shape, repeat, w, t, g[f"{name}/times"], HB)
        np.testing.assert_allclose(res["position"], g[f"{name}/position"], atol=1e-12)
        np.testing.assert_array_equal(o.bloch_permutation(shape, repeat), g[f"{name}/permutation"])
