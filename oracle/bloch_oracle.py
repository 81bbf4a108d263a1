"""CPU ORACLE (test infrastructure, NOT product code) for the Bloch build -> eigh -> evolve hot path.

This file is a plain numpy (float64 / complex128) restatement of what the reference
``Matt-Ord/slate_quantum`` computes on the path named by ``BASELINE.json:north_star``.
Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may import it; the shipped package ``slate_quantum_b200`` never does.

Why a restatement and not the reference itself
----------------------------------------------
* the reference needs CPython >= 3.14 (``/root/reference/pyproject.toml:7``) and fails to *parse* on
  the 3.12 interpreter of this image (PEP 696 type-parameter defaults), and
* every floating point operation of the path runs inside the un-vendored dependency
  ``slate_core @ git+https://github.com/Matt-Ord/slate.git@4aa5106`` (version 0.0.2,
  ``pyproject.toml:11``, ``uv.lock:931-933``) whose source is absent from ``/root/reference``.

The arithmetic slate_core reaches on this path is numpy's own (``np.linalg.eigh`` = LAPACK zheevd,
``np.fft`` = pocketfft, ``np.matmul``), so the restatement below calls exactly those.

Parity pinning status (see tests/test_oracle_golden.py, tests/golden/README.md)
-------------------------------------------------------------------------------
PINNED by the reference's own known-answer tests / invariants, re-expressed against this oracle:
  * kinetic literal ``[0, .5, 2, 2, .5]``                        (tests/test_operator.py:22-28)
  * split raw layout ``[0,0,0, 0,.5,2,2,.5]`` and dense diagonal  (tests/test_operator.py:31-59)
  * eigenvalue multiset equality through eigh                    (tests/test_operator.py:62-75)
  * FFT-order literal ``[0,1,2,-2,-1]`` and FFT directions        (tests/test_operator.py:184-210)
  * cos / fcc potential == analytic functions                    (tests/test_operator.py:305-353)
  * Bloch blocks == supercell Hamiltonian on the (shape, repeat) grid of
    tests/test_bloch.py:16-47, 50-86, 89-126 (orthogonal 2*pi lattice, atol 1e-14 / 1e-15)
PARITY UNPINNED by the reference (it has no test): eigenvectors of Bloch blocks, the whole of
``dynamics/schrodinger.py``, non-orthogonal lattices through ``bloch.*``.  For those this oracle is
cross-checked by independent identities (dense supercell ``scipy.linalg.expm`` evolution, residuals,
``eigvalsh`` of the dense supercell matrix) in tests/test_oracle_golden.py.

Every function cites the reference file:line it follows.
"""

from __future__ import annotations

import itertools
from typing import Sequence

import numpy as np

HBAR_SI = 1.054571817e-34  # scipy.constants.hbar (CODATA 2018), the reference's default


# --------------------------------------------------------------------------------------
# index conventions
# --------------------------------------------------------------------------------------
def nk_points(n: int) -> np.ndarray:
    """FFT-ordered integer frequencies, e.g. n=5 -> [0,1,2,-2,-1].

    slate_core ``fundamental_nk_points``; literal pinned at tests/test_operator.py:195.
    """
    return np.fft.ifftshift(np.arange((-n + 1) // 2, (n + 1) // 2))


def stacked_nk(shape: Sequence[int]) -> np.ndarray:
    """(nd, N) integer mesh, ``indexing="ij"`` C-order ravel.

    operator/_build/_hamiltonian.py:61-69 (_fundamental_stacked_n_kinetic).
    """
    mesh = np.meshgrid(*(nk_points(n) for n in shape), indexing="ij")
    return np.array([m.ravel() for m in mesh])


def stacked_dk(delta_x: np.ndarray) -> np.ndarray:
    """Reciprocal vectors, row i = dk_i with dk_i . a_j = 2 pi delta_ij (a_j = full cell vectors).

    slate_core ``fundamental_stacked_dk``; pinned 1-D (L=2pi => dk=1) at tests/test_operator.py:22-28
    and on the hexagonal cell at tests/test_operator.py:323-353.
    """
    delta_x = np.asarray(delta_x, dtype=np.float64)
    return 2 * np.pi * np.linalg.inv(delta_x).T


def sample_fractions(repeat: Sequence[int]) -> np.ndarray:
    """(nd, n_k) Bloch fractions, block order = meshgrid(indexing="ij") C-order ravel.

    bloch/build.py:40-42 (values = nk/R) and :138-143 (get_sample_fractions).
    """
    mesh = np.meshgrid(*(nk_points(r) / r for r in repeat), indexing="ij")
    return np.array([m.ravel() for m in mesh])


# --------------------------------------------------------------------------------------
# a-1 kinetic energy
# --------------------------------------------------------------------------------------
def wrap_k_points(k_points: np.ndarray, delta_k: Sequence[float]) -> np.ndarray:
    """operator/_build/_hamiltonian.py:32-37 -- half-open [-d/2, d/2) via np.mod (quirk Q2)."""
    shift = np.array(delta_k).reshape(-1, 1) / 2
    return np.mod(k_points + shift, 2 * shift) - shift


def kinetic_energy(
    delta_x: np.ndarray,
    shape: Sequence[int],
    mass: float,
    fraction: np.ndarray | None = None,
    *,
    hbar: float = HBAR_SI,
) -> np.ndarray:
    """Kinetic diagonal E[N] (complex128, imag 0) for one Bloch fraction.

    operator/_build/_hamiltonian.py:40-48 (_get_bloch_phase), :72-94
    (_fundamental_stacked_kinetic_points, incl. the Cartesian-row wrap with lattice-vector norms,
    quirk Q1), :121-146 (energy = sum (hbar k)^2 / 2m as complex128).
    All axes periodic ("Fourier").
    """
    nd = len(shape)
    dk = stacked_dk(delta_x)  # fundamental_stacked_dk
    fraction = np.zeros(nd) if fraction is None else np.asarray(fraction, dtype=np.float64)
    bloch_phase = np.tensordot(dk, fraction, axes=(0, 0))
    points = np.einsum("ij,ik->jk", dk, stacked_nk(shape).astype(np.float64))
    points = points + bloch_phase[:, np.newaxis]
    delta_k = dk * np.asarray(shape, dtype=np.float64)[:, np.newaxis]  # fundamental_stacked_delta_k
    points = wrap_k_points(points, tuple(np.linalg.norm(delta_k, axis=1)))
    return np.sum(np.square(hbar * points) / (2 * mass), axis=0, dtype=np.complex128)


# --------------------------------------------------------------------------------------
# a-2 potentials (raw data = Fourier components in a cropped transformed basis)
# --------------------------------------------------------------------------------------
def outer_product(*arrays: np.ndarray) -> np.ndarray:
    """_util/_prod.py:6-10."""
    grids = np.meshgrid(*arrays, indexing="ij")
    return np.prod(grids, axis=0)


def cos_potential_components(shape: Sequence[int], height: float, offset: float = 0.0) -> np.ndarray:
    """(3,)*nd components at frequencies [0,+1,-1] per axis. operator/_build/_potential.py:86-108."""
    nd = len(shape)
    size = int(np.prod(shape))
    data = outer_product(*(np.array([2 * (1 + offset), 1, 1], dtype=np.complex128),) * nd)
    return 0.25**nd * height * data * np.sqrt(size)


def sin_potential_components(shape: Sequence[int], height: float, offset: float = 0.0) -> np.ndarray:
    """operator/_build/_potential.py:111-131."""
    nd = len(shape)
    size = int(np.prod(shape))
    data = outer_product(*(np.array([2 * (1 + offset), 1j, -1j]),) * nd)
    return 0.25**nd * height * data * np.sqrt(size)


def fcc_potential_components(shape: Sequence[int], height: float) -> np.ndarray:
    """operator/_build/_potential.py:167-190 (2-D only, as in the reference)."""
    assert len(shape) == 2  # noqa: PLR2004
    size = int(np.prod(shape))
    data = np.array([[3, 1, 1], [1, 1, 0], [1, 0, 1]]).astype(np.complex128)
    return (1 / 3) ** 2 * data * height * np.sqrt(size)


def pad_components(shape: Sequence[int], components: np.ndarray) -> np.ndarray:
    """Zero-pad cropped Fourier components to ``shape`` in FFT order (CroppedBasis semantics).

    CroppedBasis(n, transformed) keeps the n lowest |frequency| entries in FFT order, i.e.
    cropped index j <-> frequency nk_points(n)[j] (assumption U4, pinned by
    tests/test_operator.py:305-320 and tests/test_bloch.py:89-126).
    """
    components = np.asarray(components, dtype=np.complex128)
    full = np.zeros(tuple(shape), dtype=np.complex128)
    idx = [nk_points(c) % n for c, n in zip(components.shape, shape, strict=True)]
    full[np.ix_(*idx)] = components
    return full


def potential_position_values(shape: Sequence[int], components: np.ndarray) -> np.ndarray:
    """V(x) on the grid.

    The diagonal of a position operator is stored on the *dual* outer basis
    (operator/_diagonal.py:64-73: ``RecastBasis(inner, inner_recast.dual_basis(),
    outer_recast.dual_basis())``), and a dual index goes transformed -> fundamental with
    fft(norm="ortho") (tests/test_operator.py:197-204).  Pinned by ``sin_potential(meta, 2) ==
    1 + sin x`` (tests/test_operator.py:316-320), which distinguishes fft from ifft.
    """
    return np.fft.fftn(pad_components(shape, components), norm="ortho")


def repeat_components(shape: Sequence[int], components: np.ndarray, repeat: Sequence[int]) -> np.ndarray:
    """Supercell Fourier table of the repeated potential, zero padded, FFT order.

    operator/_build/_potential.py:56-83: unit-cell component j goes to supercell index j*R
    (TruncatedBasis(Truncation(n=N_a, step=R_a, offset=0))), data * sqrt(prod R).
    """
    full = pad_components(shape, components)
    big = np.zeros(tuple(n * r for n, r in zip(shape, repeat, strict=True)), dtype=np.complex128)
    big[tuple(slice(None, None, r) for r in repeat)] = full * np.sqrt(np.prod(repeat))
    return big


def convolution_matrix(full_table: np.ndarray) -> np.ndarray:
    """Momentum-basis matrix of a position-diagonal operator: V[k,k'] = Vt[(k'-k) mod shape]/sqrt(N).

    With v(x) = fft_ortho(Vt) (see :func:`potential_position_values`) and the ket index going
    position -> momentum by fft(ortho), bra by ifft(ortho):
    V[k,k'] = (1/N) sum_x v(x) e^{-2 pi i (k-k') x/N} = Vt[(k'-k) mod N] / sqrt(N).
    (Symmetric tables such as cos/fcc cannot tell k-k' from k'-k; sin_potential can.)
    """
    shape = full_table.shape
    size = int(np.prod(shape))
    idx = np.array(np.unravel_index(np.arange(size), shape))  # (nd, N) array indices
    diff = (idx[:, None, :] - idx[:, :, None]) % np.array(shape)[:, None, None]
    flat = np.ravel_multi_index(tuple(diff), shape)
    return full_table.ravel()[flat] / np.sqrt(size)


def convolution_matrix_via_fft(full_table: np.ndarray) -> np.ndarray:
    """The reference's own densification route (slate_core): values -> diag in x -> 2-sided FFT.

    Ket index: position -> momentum is fft(ortho); bra (dual) index the opposite direction
    (tests/test_operator.py:197-204).  Independent check of :func:`convolution_matrix`.
    """
    shape = full_table.shape
    nd = len(shape)
    size = int(np.prod(shape))
    v_x = np.fft.fftn(full_table, norm="ortho").ravel()
    dense = np.diag(v_x).reshape(*shape, *shape)
    dense = np.fft.fftn(dense, axes=tuple(range(nd)), norm="ortho")
    dense = np.fft.ifftn(dense, axes=tuple(range(nd, 2 * nd)), norm="ortho")
    return dense.reshape(size, size)


# --------------------------------------------------------------------------------------
# a-3 / a-4 Hamiltonians
# --------------------------------------------------------------------------------------
def split_raw_data(components: np.ndarray, kinetic: np.ndarray) -> np.ndarray:
    """operator/_build/_hamiltonian.py:177-182: raw = concatenate([potential.raw, kinetic.raw])."""
    return np.concatenate([np.asarray(components, dtype=np.complex128), kinetic], axis=None)


def dense_hamiltonian(
    delta_x: np.ndarray,
    shape: Sequence[int],
    full_table: np.ndarray,
    mass: float,
    fraction: np.ndarray | None = None,
    *,
    hbar: float = HBAR_SI,
) -> np.ndarray:
    """One cell's H in its transformed (momentum) basis = diag(E) + convolution matrix."""
    energy = kinetic_energy(delta_x, shape, mass, fraction, hbar=hbar)
    return np.diag(energy) + convolution_matrix(full_table)


def bloch_blocks(
    delta_x: np.ndarray,
    shape: Sequence[int],
    components: np.ndarray,
    mass: float,
    repeat: Sequence[int],
    *,
    hbar: float = HBAR_SI,
    block_begin: int = 0,
    block_count: int | None = None,
) -> np.ndarray:
    """K1: raw data [n_k, N, N] of ``bloch.build.kinetic_hamiltonian`` (bloch/build.py:146-167).

    Row index = ket k, column = bra k', block order per :func:`sample_fractions`.
    """
    fractions = sample_fractions(repeat)
    n_k = fractions.shape[1]
    block_count = n_k - block_begin if block_count is None else block_count
    conv = convolution_matrix(pad_components(shape, components))
    out = np.empty((block_count, conv.shape[0], conv.shape[0]), dtype=np.complex128)
    for i in range(block_count):
        energy = kinetic_energy(delta_x, shape, mass, fractions[:, block_begin + i], hbar=hbar)
        out[i] = conv
        out[i][np.diag_indices(conv.shape[0])] += energy
    return out


def supercell_hamiltonian(
    delta_x: np.ndarray,
    shape: Sequence[int],
    components: np.ndarray,
    mass: float,
    repeat: Sequence[int],
    *,
    hbar: float = HBAR_SI,
) -> np.ndarray:
    """Dense H of the repeated cell in the supercell momentum basis (tests/test_bloch.py:32-45).

    repeat_potential + operator.build.kinetic_hamiltonian on the RepeatedMetadata volume
    (metadata/_repeat.py:9-25: size N*R, delta = R*delta).
    """
    big_shape = tuple(n * r for n, r in zip(shape, repeat, strict=True))
    big_delta_x = np.asarray(delta_x, dtype=np.float64) * np.asarray(repeat, dtype=np.float64)[:, None]
    table = repeat_components(shape, components, repeat)
    return dense_hamiltonian(big_delta_x, big_shape, table, mass, None, hbar=hbar)


# --------------------------------------------------------------------------------------
# a-5 K5 permutations, literal numpy restatements
# --------------------------------------------------------------------------------------
def shifted_from_inner(vectors: np.ndarray, n_states: int, n_repeats: int, axis: int = -1) -> np.ndarray:
    """bloch/_shifted_basis.py:87-111."""
    axis %= vectors.ndim
    shift = n_repeats // 2 + (n_repeats * (n_states // 2))
    shifted = np.roll(vectors, shift, axis=axis)
    stacked = shifted.reshape(*vectors.shape[:axis], n_states, n_repeats, *vectors.shape[axis + 1 :])
    return np.fft.ifftshift(stacked, axes=(axis, axis + 1)).reshape(vectors.shape)


def shifted_into_inner(vectors: np.ndarray, n_states: int, n_repeats: int, axis: int = -1) -> np.ndarray:
    """bloch/_shifted_basis.py:62-84."""
    axis %= vectors.ndim
    stacked = vectors.reshape(*vectors.shape[:axis], n_states, n_repeats, *vectors.shape[axis + 1 :])
    shifted = np.fft.fftshift(stacked, axes=(axis, axis + 1))
    shift = n_repeats // 2 + (n_repeats * (n_states // 2))
    return np.roll(shifted.reshape(vectors.shape), -shift, axis=axis)


def transposed_from_inner(
    vectors: np.ndarray, n_states: Sequence[int], n_repeats: Sequence[int], axis: int = -1
) -> np.ndarray:
    """bloch/_transposed_basis.py:136-167: [(in,blk),...] -> [(blk...),(in...)]."""
    axis %= vectors.ndim
    n_dim = len(n_repeats)
    stacked = vectors.reshape(
        *vectors.shape[:axis],
        *itertools.chain(*zip(n_states, n_repeats, strict=True)),
        *vectors.shape[axis + 1 :],
    )
    flipped = stacked.transpose(
        *range(axis),
        *range(axis + 1, axis + 2 * n_dim, 2),
        *range(axis, axis + 2 * n_dim, 2),
        *range(axis + 2 * n_dim, stacked.ndim),
    )
    return flipped.reshape(vectors.shape)


def transposed_into_inner(
    vectors: np.ndarray, n_states: Sequence[int], n_repeats: Sequence[int], axis: int = -1
) -> np.ndarray:
    """bloch/_transposed_basis.py:102-133: [(blk...),(in...)] -> [(in,blk),...]."""
    axis %= vectors.ndim
    n_dim = len(n_repeats)
    stacked = vectors.reshape(*vectors.shape[:axis], *n_repeats, *n_states, *vectors.shape[axis + 1 :])
    flipped = stacked.transpose(
        *range(axis),
        *itertools.chain(*((axis + i + n_dim, axis + i) for i in range(n_dim))),
        *range(axis + 2 * n_dim, stacked.ndim),
    )
    return flipped.reshape(vectors.shape)


def bloch_from_supercell(
    vectors: np.ndarray, shape: Sequence[int], repeat: Sequence[int], axis: int = -1
) -> np.ndarray:
    """Supercell-momentum (TupleBasis of TransformedBasis, C order) -> Bloch [(blk...),(in...)] order.

    = per-axis BlochShiftedBasis.__from_inner__ then BlochTransposedBasis.__from_inner__
    (basis nesting at bloch/build.py:79-88).
    """
    axis %= vectors.ndim
    nd = len(shape)
    big = tuple(n * r for n, r in zip(shape, repeat, strict=True))
    stacked = vectors.reshape(*vectors.shape[:axis], *big, *vectors.shape[axis + 1 :])
    for a in range(nd):
        stacked = shifted_from_inner(stacked, shape[a], repeat[a], axis=axis + a)
    flat = stacked.reshape(vectors.shape)
    return transposed_from_inner(flat, shape, repeat, axis=axis)


def bloch_into_supercell(
    vectors: np.ndarray, shape: Sequence[int], repeat: Sequence[int], axis: int = -1
) -> np.ndarray:
    """Inverse of :func:`bloch_from_supercell` (the two ``__into_inner__``)."""
    axis %= vectors.ndim
    nd = len(shape)
    big = tuple(n * r for n, r in zip(shape, repeat, strict=True))
    flat = transposed_into_inner(vectors, shape, repeat, axis=axis)
    stacked = flat.reshape(*vectors.shape[:axis], *big, *vectors.shape[axis + 1 :])
    for a in range(nd):
        stacked = shifted_into_inner(stacked, shape[a], repeat[a], axis=axis + a)
    return stacked.reshape(vectors.shape)


def bloch_permutation(shape: Sequence[int], repeat: Sequence[int]) -> np.ndarray:
    """perm with bloch[i] = supercell[perm[i]] -- closed form used by the CUDA kernel.

    Element (block b, inner p) with signed FFT-order integers (q_a, k_a) per axis sits at supercell
    index (k_a * R_a + q_a) mod (N_a R_a) per axis (C order over axes).
    """
    nd = len(shape)
    n_k = int(np.prod(repeat))
    size = int(np.prod(shape))
    b_idx = np.array(np.unravel_index(np.arange(n_k), repeat))  # (nd, n_k)
    p_idx = np.array(np.unravel_index(np.arange(size), shape))  # (nd, N)
    perm = np.zeros((n_k, size), dtype=np.int64)
    for a in range(nd):
        n, r = shape[a], repeat[a]
        q = nk_points(r)[b_idx[a]][:, None]
        k = nk_points(n)[p_idx[a]][None, :]
        perm = perm * (n * r) + (k * r + q) % (n * r)
    return perm.ravel()


def embed_blocks(blocks: np.ndarray, shape: Sequence[int], repeat: Sequence[int]) -> np.ndarray:
    """Block-diagonal [n_k,N,N] -> dense supercell momentum matrix [M,M] (BlockDiagonalBasis + K5)."""
    n_k, size, _ = blocks.shape
    m = n_k * size
    dense = np.zeros((m, m), dtype=blocks.dtype)
    for b in range(n_k):
        dense[b * size : (b + 1) * size, b * size : (b + 1) * size] = blocks[b]
    dense = bloch_into_supercell(dense, shape, repeat, axis=0)
    return bloch_into_supercell(dense, shape, repeat, axis=1)


# --------------------------------------------------------------------------------------
# a-8 K4 FFT basis change
# --------------------------------------------------------------------------------------
def position_to_momentum(vectors: np.ndarray, big_shape: Sequence[int], axis: int = -1) -> np.ndarray:
    """Ket index: fundamental -> transformed = fft(norm="ortho") per axis (tests/test_operator.py:197-204)."""
    axis %= vectors.ndim
    nd = len(big_shape)
    stacked = vectors.reshape(*vectors.shape[:axis], *big_shape, *vectors.shape[axis + 1 :])
    out = np.fft.fftn(stacked, axes=tuple(range(axis, axis + nd)), norm="ortho")
    return out.reshape(vectors.shape)


def momentum_to_position(vectors: np.ndarray, big_shape: Sequence[int], axis: int = -1) -> np.ndarray:
    """Ket index: transformed -> fundamental = ifft(norm="ortho") per axis."""
    axis %= vectors.ndim
    nd = len(big_shape)
    stacked = vectors.reshape(*vectors.shape[:axis], *big_shape, *vectors.shape[axis + 1 :])
    out = np.fft.ifftn(stacked, axes=tuple(range(axis, axis + nd)), norm="ortho")
    return out.reshape(vectors.shape)


# --------------------------------------------------------------------------------------
# a-6 K2 eigh, a-7 K3 evolve
# --------------------------------------------------------------------------------------
def eigh_blocks(blocks: np.ndarray) -> tuple[np.ndarray, np.ndarray]:
    """K2: per-block ``np.linalg.eigh`` (slate_core.linalg.into_diagonal_hermitian on a
    BlockDiagonalBasis operator; call site operator/_linalg.py:56).

    Returns (E [n_k, N] float64 ascending within a block (quirk Q3),
             transform [n_k, N(eigen index), N(component)] = eigh(...)[1]^T, rows = eigenvectors).
    """
    w, v = np.linalg.eigh(blocks)
    return w, np.ascontiguousarray(np.swapaxes(v, -1, -2))


def project(transform: np.ndarray, psi_bloch: np.ndarray) -> np.ndarray:
    """K3a: c[b,n] = sum_p conj(V_b[p,n]) psi[b,p]  (ExplicitUnitaryBasis inner -> explicit, U6)."""
    return np.einsum("bnp,bp->bn", transform.conj(), psi_bloch)


def evolve_eigenbasis(coefficients: np.ndarray, eigenvalues: np.ndarray, times: np.ndarray, hbar: float) -> np.ndarray:
    """K3b, literally dynamics/schrodinger.py:48-53 (operation order -1j * E * t / hbar, E complex128)."""
    coefficients = np.asarray(coefficients).ravel()
    eig = np.asarray(eigenvalues).ravel().astype(np.complex128)
    time_values = np.asarray(times, dtype=np.float64)
    return coefficients[np.newaxis, :] * np.exp(-1j * eig * time_values[:, np.newaxis] / hbar)


def back_transform(transform: np.ndarray, amplitudes: np.ndarray) -> np.ndarray:
    """K3c: psi[t,b,p] = sum_n a[t,b,n] V_b[p,n] (explicit -> inner)."""
    n_k, size, _ = transform.shape
    a = amplitudes.reshape(-1, n_k, size)
    return np.einsum("tbn,bnp->tbp", a, transform)


def evolve_pipeline(
    psi0_position: np.ndarray,
    shape: Sequence[int],
    repeat: Sequence[int],
    eigenvalues: np.ndarray,
    transform: np.ndarray,
    times: np.ndarray,
    hbar: float,
) -> dict[str, np.ndarray]:
    """Full K4 -> K5 -> K3a -> K3b -> K3c -> K5 -> K4 chain."""
    big = tuple(n * r for n, r in zip(shape, repeat, strict=True))
    n_k = int(np.prod(repeat))
    size = int(np.prod(shape))
    psi_k = position_to_momentum(psi0_position, big)
    psi_b = bloch_from_supercell(psi_k, shape, repeat).reshape(n_k, size)
    c = project(transform, psi_b)
    a = evolve_eigenbasis(c, eigenvalues, times, hbar)  # [T, M] eigenbasis
    psi_bt = back_transform(transform, a).reshape(len(times), n_k * size)  # bloch momentum
    psi_kt = bloch_into_supercell(psi_bt, shape, repeat, axis=-1)
    psi_xt = momentum_to_position(psi_kt, big, axis=-1)
    return {
        "coefficients": c,
        "eigen": a,
        "bloch": psi_bt,
        "momentum": psi_kt,
        "position": psi_xt,
    }


def momentum_matrix_to_position(dense: np.ndarray, big_shape: Sequence[int]) -> np.ndarray:
    """Operator.as_array(): ket index ifft(ortho), bra index fft(ortho) (tests/test_operator.py:197-204)."""
    nd = len(big_shape)
    m = dense.shape[0]
    stacked = dense.reshape(*big_shape, *big_shape)
    stacked = np.fft.ifftn(stacked, axes=tuple(range(nd)), norm="ortho")
    stacked = np.fft.fftn(stacked, axes=tuple(range(nd, 2 * nd)), norm="ortho")
    return stacked.reshape(m, m)


# --------------------------------------------------------------------------------------
# parity metrics shared by the tests (SURVEY §8d tolerances)
# --------------------------------------------------------------------------------------
def eigh_residuals(blocks: np.ndarray, eigenvalues: np.ndarray, transform: np.ndarray) -> tuple[float, float]:
    """max_b ||H V - V L||_F / ||H||_F and max_b ||V^H V - I||_F."""
    v = np.swapaxes(transform, -1, -2)  # columns = eigenvectors
    hv = blocks @ v
    res = np.linalg.norm(hv - v * eigenvalues[:, None, :], axis=(1, 2)) / np.maximum(
        np.linalg.norm(blocks, axis=(1, 2)), 1e-300
    )
    eye = np.eye(blocks.shape[-1])
    orth = np.linalg.norm(np.swapaxes(v.conj(), -1, -2) @ v - eye, axis=(1, 2))
    return float(res.max()), float(orth.max())


def subspace_projector_error(
    eigenvalues: np.ndarray, transform_a: np.ndarray, transform_b: np.ndarray, norm_h: float, tol: float = 1e-8
) -> float:
    """Compare eigenvectors through projectors onto (near-)degenerate clusters, one block.

    Clusters = runs with gaps < tol * ||H||; returns max_cluster ||P_a - P_b||_F.
    """
    n = len(eigenvalues)
    worst = 0.0
    start = 0
    for i in range(1, n + 1):
        if i == n or abs(eigenvalues[i] - eigenvalues[i - 1]) > tol * norm_h:
            va = transform_a[start:i]
            vb = transform_b[start:i]
            pa = va.T @ va.conj()
            pb = vb.T @ vb.conj()
            worst = max(worst, float(np.linalg.norm(pa - pb)))
            start = i
    return worst


# ---------------------------------------------------------------------------------------------
# Observables on position-basis state lists (reference operator/_measure.py).
# PINNED by the reference's tests/test_measure.py (position eigenstates, coherent states: x, periodic x with
# its sign, wrapped x, variance / width) re-expressed in tests/test_measure_oracle.py.  The wrap interval
# convention of slate_core's fundamental_stacked_x_points (which side of +-L/2 is closed) is NOT visible from
# the reference (assumption U9): the half-open interval [-L/2, L/2) is used, as everywhere else in the path.
def stacked_x_points(
    delta_x: np.ndarray, shape: Sequence[int], offset: Sequence[float] | None = None, wrapped: bool = False
) -> np.ndarray:
    """Cartesian coordinates [nd, prod(shape)] of the grid points, C-order ravel.

    ``offset`` is a Cartesian displacement added to every point (operator/_build/_position.py:236-239 builds it
    from the scalar ``offset`` of measure.x); ``wrapped`` folds the fractional coordinates into [-1/2, 1/2).
    """
    delta_x = np.asarray(delta_x, dtype=np.float64)
    nd = len(shape)
    a = delta_x[:nd, :nd]
    idx = np.array([g.ravel() for g in np.meshgrid(*(np.arange(n) for n in shape), indexing="ij")], dtype=np.float64)
    dx = a / np.asarray(shape, dtype=np.float64)[:, None]  # fundamental_stacked_dx: lattice vector / points
    points = np.einsum("ij,ik->jk", dx, idx)  # x = sum_i n_i dx_i: exactly i*dx in 1-D (tests/test_measure.py:15)
    if offset is not None:
        points = points + np.asarray(offset, dtype=np.float64)[:nd, None]
    if wrapped:
        frac = np.linalg.solve(a.T, points)
        points = points - np.einsum("ij,ik->jk", a, np.floor(frac + 0.5))
    return points


def _normalised_density(states: np.ndarray) -> np.ndarray:
    """|psi|^2 / sum |psi|^2 per state: normalize_each (state/_state.py:294-308) + a diagonal expectation."""
    density = np.square(np.abs(np.atleast_2d(states)))
    return density / density.sum(axis=1, keepdims=True)


def measure_all_x(
    states: np.ndarray, delta_x: np.ndarray, shape: Sequence[int], axis: int, offset: float = 0.0, wrapped: bool = False
) -> np.ndarray:
    """operator.measure.all_x (operator/_measure.py:191-210): <x_axis> of every normalised state."""
    inner_offset = [offset if i == axis else 0.0 for i in range(len(shape))]
    points = stacked_x_points(delta_x, shape, inner_offset, wrapped)[axis]
    return _normalised_density(states) @ points


def measure_all_scatter(states: np.ndarray, shape: Sequence[int], axis: int) -> np.ndarray:
    """<exp(+i q x)> for the fundamental q of ``axis`` (_get_all_fundamental_scatter, operator/_measure.py:213-239).

    Only its ANGLE is used downstream and pinned (tests/test_measure.py:33-38 fixes the sign); the modulus
    convention of slate_core's single-coefficient transformed operator (a 1/sqrt(N) factor) is not visible.
    """
    idx = np.unravel_index(np.arange(int(np.prod(shape))), shape)[axis]
    return _normalised_density(states) @ np.exp(2j * np.pi * idx / shape[axis])


def measure_all_periodic_x(states: np.ndarray, delta_x: np.ndarray, shape: Sequence[int], axis: int) -> np.ndarray:
    """operator.measure.all_periodic_x (operator/_measure.py:242-268)."""
    angle = np.angle(measure_all_scatter(states, shape, axis))
    wrapped = np.mod(angle, 2 * np.pi)
    return wrapped * (np.linalg.norm(np.asarray(delta_x)[axis]) / (2 * np.pi))


def measure_all_variance_x(states: np.ndarray, delta_x: np.ndarray, shape: Sequence[int], axis: int) -> np.ndarray:
    """operator.measure.all_variance_x (operator/_measure.py:271-292, variance_x :140-163):
    <x_w^2> with x_w the position wrapped around each state's own periodic <x>."""
    states = np.atleast_2d(states)
    centre = measure_all_periodic_x(states, delta_x, shape, axis)
    density = _normalised_density(states)
    out = np.empty(len(states))
    for i, x0 in enumerate(centre):
        inner_offset = [-x0 if a == axis else 0.0 for a in range(len(shape))]
        points = stacked_x_points(delta_x, shape, inner_offset, True)[axis]
        out[i] = density[i] @ np.square(points)
    return out


def measure_all_coherent_width(states: np.ndarray, delta_x: np.ndarray, shape: Sequence[int], axis: int) -> np.ndarray:
    """operator.measure.all_coherent_width (operator/_measure.py:295-305): sqrt(2 variance)."""
    return np.sqrt(2.0 * measure_all_variance_x(states, delta_x, shape, axis))


def coherent_state(
    delta_x: np.ndarray, shape: Sequence[int], x_0: Sequence[float], k_0: Sequence[float], sigma_0: Sequence[float]
) -> np.ndarray:
    """state.build.coherent (state/_build.py:79-99): periodic Gaussian wavepacket, normalised."""
    nd = len(shape)
    a = np.asarray(delta_x, dtype=np.float64)[:nd, :nd]
    disp = stacked_x_points(a, shape, [-x for x in x_0], True)
    distance = np.linalg.norm([d / s for d, s in zip(disp, sigma_0)], axis=0)
    phi = np.einsum("ij,i->j", disp, np.asarray(k_0, dtype=np.float64)[:nd])
    data = np.exp(1j * phi - np.square(distance) / 2)
    return data / np.sqrt(np.sum(np.square(np.abs(data))))


# momentum-space observables (operator/_measure.py:307-420).  The k points are the FFT-ordered integer frequencies
# times dk (pinned: the k-operator literal of tests/test_operator.py:184-196); only orthogonal axes are restated.
def stacked_k_points(
    delta_x: np.ndarray, shape: Sequence[int], offset: Sequence[float] | None = None, wrapped: bool = False
) -> np.ndarray:
    """Cartesian k of every momentum-basis index [nd, prod(shape)] (fundamental_stacked_k_points), optionally
    displaced by ``offset`` and folded into the first Brillouin zone [-K/2, K/2) of the grid, K_a = N_a dk_a."""
    nd = len(shape)
    dk = stacked_dk(np.asarray(delta_x, dtype=np.float64)[:nd, :nd])
    points = np.einsum("ij,ik->jk", dk, stacked_nk(shape).astype(np.float64))
    if offset is not None:
        points = points + np.asarray(offset, dtype=np.float64)[:nd, None]
    if wrapped:
        big = dk * np.asarray(shape, dtype=np.float64)[:, None]  # rows = delta_k vectors
        frac = np.linalg.solve(big.T, points)
        points = points - np.einsum("ij,ik->jk", big, np.floor(frac + 0.5))
    return points


def measure_all_k(states: np.ndarray, delta_x: np.ndarray, shape: Sequence[int], axis: int) -> np.ndarray:
    """operator.measure.all_k (operator/_measure.py:323-336): <k_axis> of every normalised position-basis state."""
    momentum = position_to_momentum(np.atleast_2d(states), shape)
    return _normalised_density(momentum) @ stacked_k_points(delta_x, shape)[axis]


def measure_all_variance_k(states: np.ndarray, delta_x: np.ndarray, shape: Sequence[int], axis: int) -> np.ndarray:
    """operator.measure.all_variance_k (operator/_measure.py:361-378, variance_k :338-358): <(k - <k>)^2> with the
    difference folded into the first Brillouin zone."""
    momentum = position_to_momentum(np.atleast_2d(states), shape)
    density = _normalised_density(momentum)
    centre = density @ stacked_k_points(delta_x, shape)[axis]
    out = np.empty(len(density))
    for i, k0 in enumerate(centre):
        offset = [-k0 if a == axis else 0.0 for a in range(len(shape))]
        out[i] = density[i] @ np.square(stacked_k_points(delta_x, shape, offset, True)[axis])
    return out


def measure_all_uncertainty(states: np.ndarray, delta_x: np.ndarray, shape: Sequence[int], axis: int) -> np.ndarray:
    """operator.measure.all_uncertainty (operator/_measure.py:381-420): sqrt(var_x var_k)."""
    return np.sqrt(measure_all_variance_x(states, delta_x, shape, axis) * measure_all_variance_k(states, delta_x, shape, axis))


# --------------------------------------------------------------------------------------
# K7: operator / state algebra - dense numpy restatements
# --------------------------------------------------------------------------------------
def x_operator_array(
    delta_x: np.ndarray, shape: Sequence[int], axis: int, offset: float = 0.0, wrapped: bool = False
) -> np.ndarray:
    """operator.build.x(...).as_array() (operator/_build/_position.py:228-242): diagonal in the position basis."""
    off = [offset if i == axis else 0.0 for i in range(len(shape))]
    return np.diag(stacked_x_points(delta_x, shape, off, wrapped)[axis]).astype(np.complex128)


def k_operator_array(delta_x: np.ndarray, shape: Sequence[int], axis: int, power: int = 1) -> np.ndarray:
    """operator.build.k / k_squared(...).as_array() (operator/_build/_momentum.py:37-62): diagonal in the momentum
    basis, written in the position basis with the convention pinned by tests/test_operator.py:197-204."""
    diag = np.diag(stacked_k_points(delta_x, shape)[axis].astype(np.complex128) ** power)
    return momentum_matrix_to_position(diag, shape)


def operator_matmul(lhs: np.ndarray, rhs: np.ndarray) -> np.ndarray:
    """operator.matmul: A_ik B_kj (operator/_linalg.py:108-119); also matmul_list_operator / matmul_operator_list
    (:172-196) when one operand carries a leading list axis."""
    return np.matmul(lhs, rhs)


def operator_commutator(lhs: np.ndarray, rhs: np.ndarray) -> np.ndarray:
    """operator.commute / get_commutator_operator_list: lhs rhs - rhs lhs (operator/_linalg.py:122-137, :199-215)."""
    return np.matmul(lhs, rhs) - np.matmul(rhs, lhs)


def operator_dagger(op: np.ndarray) -> np.ndarray:
    """operator.dagger / dagger_each (operator/_linalg.py:140-169)."""
    return np.conj(np.swapaxes(op, -1, -2))


def operator_apply(op: np.ndarray, state: np.ndarray) -> np.ndarray:
    """operator.apply: (O psi)_i = O_ij psi_j (operator/_operator.py:210-218)."""
    return op @ state


def expectation_of_each(op: np.ndarray, states: np.ndarray) -> np.ndarray:
    """operator.expectation_of_each: conj(psi_ai) O_ij psi_aj for every a (operator/_operator.py:196-207);
    NOT divided by the norm (the reference does not normalise here)."""
    states = np.atleast_2d(states)
    return np.einsum("ai,ij,aj->a", states.conj(), op, states)


def inner_product_each(states_0: np.ndarray, states_1: np.ndarray) -> np.ndarray:
    """state.inner_product_each (state/_state.py:208-218)."""
    return np.einsum("ji,ji->j", np.conj(states_0), states_1)


def all_inner_product(states_0: np.ndarray, states_1: np.ndarray) -> np.ndarray:
    """state.all_inner_product (state/_state.py:221-232)."""
    return np.conj(states_0) @ states_1.T


def occupations(states: np.ndarray) -> np.ndarray:
    """state.get_occupations / get_all_occupations: abs(conj(a) a) of the stored coefficients
    (state/_state.py:71-87, :265-308)."""
    return np.abs(np.conj(states) * states)


def normalize_each(states: np.ndarray) -> np.ndarray:
    """state.normalize_each: every row divided by sqrt(<psi|psi>) (state/_state.py:249-262)."""
    states = np.atleast_2d(states)
    return states / np.sqrt(inner_product_each(states, states))[:, np.newaxis]


BOLTZMANN_SI = 1.380649e-23


def temperature_corrected_operators(
    hamiltonian: np.ndarray, operators: np.ndarray, temperature: float, eta: float = 1.0, hbar: float = HBAR_SI
) -> np.ndarray:
    """noise.build.temperature_corrected_operators (noise/build/_build.py:141-155) on dense arrays [L, n, n]."""
    kb_t = BOLTZMANN_SI * temperature
    correction = operator_commutator(hamiltonian, operators) * (-1 * np.sqrt(eta / (8 * kb_t * hbar**2)))
    return correction + operators * np.sqrt(2 * eta * kb_t / hbar**2)


def hamiltonian_shift(hamiltonian: np.ndarray, operators: np.ndarray, eta: float, hbar: float = HBAR_SI) -> np.ndarray:
    """noise.build.hamiltonian_shift (noise/build/_build.py:158-172): i eta / (4 hbar) [H, sum_m L_m^dagger L_m]."""
    shift_product = np.einsum("ikj,ikl->jl", np.conj(operators), operators)
    return operator_commutator(hamiltonian, shift_product) * (1j * eta / (4 * hbar))


def measure_all_potential_from_function(
    states: np.ndarray, delta_x: np.ndarray, shape: Sequence[int], fn, offset=None, wrapped: bool = False
) -> np.ndarray:
    """operator.measure.all_potential_from_function (operator/_measure.py:41-61): <f(x)> of the normalised states."""
    weights = np.asarray(fn(tuple(stacked_x_points(delta_x, shape, offset, wrapped)))).ravel()
    return np.real(_normalised_density(np.atleast_2d(states)) @ weights)


def measure_all_momentum_from_function(
    states: np.ndarray, delta_x: np.ndarray, shape: Sequence[int], fn, offset=None, wrapped: bool = False
) -> np.ndarray:
    """operator.measure.momentum_from_function (operator/_measure.py:64-83) for every state: <f(k)>."""
    weights = np.asarray(fn(tuple(stacked_k_points(delta_x, shape, offset, wrapped)))).ravel()
    momentum = position_to_momentum(np.atleast_2d(states), shape)
    return np.real(_normalised_density(momentum) @ weights)


# --------------------------------------------------------------------------------------
# sampled-block parity at full size (checker for bench.py's parity gate and the full-size tests)
# --------------------------------------------------------------------------------------
def evolve_sampled_blocks(
    delta_x: np.ndarray,
    shape: Sequence[int],
    components: np.ndarray,
    mass: float,
    repeat: Sequence[int],
    psi0_position: np.ndarray,
    times: np.ndarray,
    blocks: Sequence[int],
    hbar: float = HBAR_SI,
) -> np.ndarray:
    """Bloch-ordered momentum components of the sampled k-blocks at `times`: [len(times), len(blocks), N].

    The chain of :func:`evolve_pipeline` (K4 -> K5 -> K3a -> K3b -> K3c) restricted to a few blocks: every block
    evolves on its own (the Hamiltonian is block diagonal, bloch/build.py:131-135), so the oracle can check
    sampled blocks of a full-size configuration at ANY time row in milliseconds per block.
    """
    big = tuple(n * r for n, r in zip(shape, repeat, strict=True))
    n_k, size = int(np.prod(repeat)), int(np.prod(shape))
    psi_k = position_to_momentum(np.asarray(psi0_position, dtype=np.complex128).reshape(1, -1), big)
    psi_b = bloch_from_supercell(psi_k, shape, repeat).reshape(n_k, size)
    times = np.asarray(times, dtype=np.float64)
    out = np.empty((times.size, len(blocks), size), dtype=np.complex128)
    for i, b in enumerate(blocks):
        h = bloch_blocks(delta_x, shape, components, mass, repeat, hbar=hbar, block_begin=int(b), block_count=1)
        w, t = eigh_blocks(h)
        c = project(t, psi_b[int(b)][None, :])
        a = evolve_eigenbasis(c, w, times, hbar)
        out[:, i, :] = back_transform(t, a).reshape(times.size, size)
    return out


def position_rows_to_blocks(
    rows: np.ndarray, shape: Sequence[int], repeat: Sequence[int], blocks: Sequence[int]
) -> np.ndarray:
    """Position-basis supercell states [R, M] -> their Bloch-ordered momentum components in the sampled blocks
    [R, len(blocks), N] (K4 forward + K5 from_inner, the inverse of the tail of :func:`evolve_pipeline`)."""
    big = tuple(n * r for n, r in zip(shape, repeat, strict=True))
    n_k, size = int(np.prod(repeat)), int(np.prod(shape))
    rows = np.asarray(rows, dtype=np.complex128).reshape(-1, n_k * size)
    psi_b = bloch_from_supercell(position_to_momentum(rows, big), shape, repeat).reshape(rows.shape[0], n_k, size)
    return psi_b[:, [int(b) for b in blocks], :]


def sampled_block_parity(
    delta_x: np.ndarray,
    shape: Sequence[int],
    components: np.ndarray,
    mass: float,
    repeat: Sequence[int],
    psi0_position: np.ndarray,
    times: np.ndarray,
    rows: np.ndarray,
    blocks: Sequence[int],
    hbar: float = HBAR_SI,
    basis: str = "position",
) -> float:
    """max |device - oracle| over the sampled blocks of evolved states `rows` (one per entry of `times`).

    basis "position": rows are supercell position-basis states [R, M]; "bloch": rows are Bloch-ordered
    momentum states [R, M] (the k-sharded output)."""
    want = evolve_sampled_blocks(delta_x, shape, components, mass, repeat, psi0_position, times, blocks, hbar)
    n_k, size = int(np.prod(repeat)), int(np.prod(shape))
    if basis == "position":
        got = position_rows_to_blocks(rows, shape, repeat, blocks)
    else:
        got = np.asarray(rows).reshape(-1, n_k, size)[:, [int(b) for b in blocks], :]
    return float(np.abs(got - want).max())


# --------------------------------------------------------------------------------------
# noise kernels and their eigenvalue diagonalisation
# --------------------------------------------------------------------------------------
def noise_kernel_from_operators(values: np.ndarray, operators: np.ndarray) -> np.ndarray:
    """noise/_kernel.py:38-66: beta[i, j, k, l] = sum_a lambda_a conj(L_a[j, i]) L_a[k, l]."""
    ops = np.asarray(operators, dtype=np.complex128)
    return np.einsum("a,aji,akl->ijkl", np.asarray(values), np.conj(ops), ops)


def diagonal_noise_kernel_from_operators(values: np.ndarray, diagonals: np.ndarray) -> np.ndarray:
    """noise/_kernel.py:145-177: beta[i, j] = sum_a lambda_a conj(d_a[i]) d_a[j]."""
    d = np.asarray(diagonals, dtype=np.complex128)
    return np.einsum("a,ai,aj->ij", np.asarray(values), np.conj(d), d)


def noise_operators_diagonal_eigenvalue(kernel: np.ndarray) -> tuple[np.ndarray, np.ndarray]:
    """noise/diagonalize/_eigenvalue.py:80-131: (eigenvalues, conj(U^T)) of eigh(beta); row b = noise operator b."""
    np.testing.assert_allclose(kernel, np.conj(kernel.T), atol=1e-10 * max(np.abs(kernel).max(), 1e-300))
    res = np.linalg.eigh(kernel)
    return res.eigenvalues, np.conj(res.eigenvectors.T)


def noise_operators_eigenvalue(kernel: np.ndarray) -> tuple[np.ndarray, np.ndarray]:
    """noise/diagonalize/_eigenvalue.py:22-77 for beta[i, j, k, l]: swap the first two indices, eigh of the
    [n^2, n^2] matrix, operators = conj(U^T) reshaped [n^2, n, n]."""
    n = kernel.shape[0]
    data = kernel.swapaxes(0, 1).reshape(n * n, n * n)
    w, ops = noise_operators_diagonal_eigenvalue(data)
    return w, ops.reshape(n * n, n, n)
