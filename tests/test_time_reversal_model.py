"""CPU check of the time-reversal pairing of Bloch blocks that `sqb_bloch_time_reversal_mirror` (csrc/bloch_build.cu)
relies on: on an orthogonal lattice with a Hermitian-symmetric potential table the oracle's blocks satisfy
H_{-k}[pi g, pi g'] = conj(H_k[g, g']) with s = q + R g -> -s (mod R N per axis) on the supercell momentum index
(reference: kinetic diagonal operator/_build/_hamiltonian.py:121-146, block order bloch/build.py:138-143, potential part
bloch/build.py:123-128).  The index map below is the numpy twin of the kernel's `time_reversal_partner`."""
from __future__ import annotations

import numpy as np
import pytest

from oracle import bloch_oracle as o

HBAR = o.HBAR_SI


def _fft_value(i, n):
    return np.where(i < (n + 1) // 2, i, i - n)


def partner(block_index, repeat, shape):
    """Block multi-index -> (partner block multi-index, perm with perm[g_flat] = g'_flat)."""
    maps, partner_index = [], []
    for a in range(len(shape)):
        r, na = repeat[a], shape[a]
        m = r * na
        s = (_fft_value(block_index[a], r) + r * _fft_value(np.arange(na), na)) % m
        sp = (m - s) % m
        qi = sp % r
        assert np.all(qi == qi[0])
        maps.append(((sp - _fft_value(qi[0], r)) // r) % na)
        partner_index.append(int(qi[0]))
    grids = np.meshgrid(*maps, indexing="ij")
    return partner_index, np.ravel_multi_index([g.ravel() for g in grids], shape)


def _worst_mismatch(dx, shape, repeat, comps):
    blocks = o.bloch_blocks(dx, shape, comps, HBAR**2, repeat)
    worst = 0.0
    for b in range(blocks.shape[0]):
        pi, perm = partner(np.unravel_index(b, repeat), repeat, shape)
        bp = int(np.ravel_multi_index(pi, repeat))
        # the pairing is an involution on (block, in-block index)
        pi2, perm2 = partner(pi, repeat, shape)
        assert int(np.ravel_multi_index(pi2, repeat)) == b and np.array_equal(perm2[perm], np.arange(len(perm)))
        h, hp = blocks[b], blocks[bp]
        worst = max(worst, float(np.abs(hp[np.ix_(perm, perm)] - np.conj(h)).max() / np.abs(h).max()))
    return worst


@pytest.mark.parametrize(
    ("shape", "repeat", "kind"),
    [((8,), (6,), "cos"), ((8,), (5,), "sin"), ((7,), (4,), "sin"), ((6, 4), (4, 3), "cos"), ((5, 4), (3, 4), "sin"),
     ((3, 4, 2), (2, 3, 4), "cos")],
)
def test_blocks_of_k_and_minus_k_are_conjugate_permutations(shape, repeat, kind):
    dx = 2 * np.pi * np.eye(len(shape))
    comps = o.cos_potential_components(shape, 3.0) if kind == "cos" else o.sin_potential_components(shape, 4.0, 0.2)
    assert _worst_mismatch(dx, shape, repeat, comps) < 1e-14


def test_pairing_fails_on_the_hexagonal_lattice_and_the_solver_must_not_use_it():
    """Quirk Q1 (Cartesian-row Brillouin-zone wrap with lattice-vector norms) is not even in the lattice index: the
    solver only pairs blocks on orthogonal lattices (pipeline.BlochSolver._time_reversal_applies)."""
    hexv = 2 * np.pi * np.array([[1.0, 0.0], [0.5, np.sqrt(3) / 2]])
    assert _worst_mismatch(hexv, (4, 4), (4, 2), o.fcc_potential_components((4, 4), 5.0)) > 1e-3


def test_representative_blocks_are_a_prefix_and_cover_every_pair():
    """The blocks with first k-index <= R_0 / 2 (a contiguous prefix) contain a partner of every other block."""
    for shape, repeat in (((4,), (7,)), ((4, 4), (8, 3)), ((2, 2, 2), (5, 2, 3))):
        n_k = int(np.prod(repeat))
        n_rep = (repeat[0] // 2 + 1) * (n_k // repeat[0])
        for b in range(n_rep, n_k):
            pi, _ = partner(np.unravel_index(b, repeat), repeat, shape)
            assert int(np.ravel_multi_index(pi, repeat)) < n_rep
