"""Parity of every CUDA kernel (through the C ABI) against the CPU oracle.  All tests need a B200."""

import itertools

import numpy as np
import pytest

from oracle import bloch_oracle as o

pytestmark = pytest.mark.gpu

HBAR = o.HBAR_SI
SHAPES_1D = [(x,) for x in [3, 5]]
SHAPES_2D = [*itertools.product([3, 4], [3, 4]), (5, 5)]
REF_GRID = [*itertools.product(SHAPES_1D, SHAPES_1D), *itertools.product(SHAPES_2D, SHAPES_2D)]


def _gpu():
    import torch

    from slate_quantum_b200 import kernels

    return torch, kernels


def _build_gpu(delta_x, shape, comps, mass, repeat, hbar=HBAR, **kw):
    torch, kernels = _gpu()
    table = kernels.to_device(o.pad_components(shape, comps).ravel(), torch.complex128)
    h = kernels.bloch_build(shape, repeat, o.stacked_dk(delta_x), table, mass, hbar, **kw)
    return h


# ---------------------------------------------------------------------------------------- K1
@pytest.mark.parametrize(("shape", "repeat"), REF_GRID)
@pytest.mark.parametrize("kind", ["cos1", "cos0", "sin"])
def test_bloch_build_reference_grid(cuda, shape, repeat, kind):
    """tests/test_bloch.py:16-47 / 50-86 shapes; tolerance = the reference's own (1e-14 / 1e-15)."""
    nd = len(shape)
    dx = 2 * np.pi * np.eye(nd)
    comps = {
        "cos1": o.cos_potential_components(shape, 1.0),
        "cos0": o.cos_potential_components(shape, 0.0),
        "sin": o.sin_potential_components(shape, 0.7, 0.1),
    }[kind]
    got = _build_gpu(dx, shape, comps, HBAR**2, repeat).cpu().numpy()
    want = o.bloch_blocks(dx, shape, comps, HBAR**2, repeat)
    np.testing.assert_allclose(got, want, rtol=0, atol=1e-14)
    # and against the dense supercell Hamiltonian, as the reference test does
    full = o.supercell_hamiltonian(dx, shape, comps, HBAR**2, repeat)
    np.testing.assert_allclose(o.embed_blocks(got, shape, repeat), full, rtol=0, atol=1e-14)


def test_bloch_build_hexagonal_quirk_q1(cuda):
    """Non-orthogonal lattice: the literal Cartesian-row wrap of _hamiltonian.py:87-93 (quirk Q1)."""
    dx = 2 * np.pi * np.array([[np.sin(np.pi / 3), np.cos(np.pi / 3)], [0.0, 1.0]])
    shape, repeat = (6, 4), (5, 4)
    comps = o.fcc_potential_components(shape, 10.0)
    got = _build_gpu(dx, shape, comps, HBAR**2, repeat).cpu().numpy()
    want = o.bloch_blocks(dx, shape, comps, HBAR**2, repeat)
    scale = np.abs(want).max()
    np.testing.assert_allclose(got, want, rtol=0, atol=1e-14 * scale)


def test_bloch_build_block_range_and_3d(cuda):
    dx = 2 * np.pi * np.eye(3)
    shape, repeat = (3, 4, 5), (2, 3, 2)
    comps = o.cos_potential_components(shape, 4.0)
    want = o.bloch_blocks(dx, shape, comps, HBAR**2, repeat)
    got = _build_gpu(dx, shape, comps, HBAR**2, repeat, block_begin=5, block_count=4).cpu().numpy()
    np.testing.assert_allclose(got, want[5:9], rtol=0, atol=1e-13)
    torch, kernels = _gpu()
    e = kernels.bloch_kinetic(shape, repeat, o.stacked_dk(dx), HBAR**2, HBAR).cpu().numpy()
    fr = o.sample_fractions(repeat)
    for b in range(fr.shape[1]):
        np.testing.assert_allclose(e[b], o.kinetic_energy(dx, shape, HBAR**2, fr[:, b]), rtol=1e-14, atol=1e-14)


def test_kinetic_literal(cuda):
    """tests/test_operator.py:22-28 literal through the CUDA kernel."""
    torch, kernels = _gpu()
    e = kernels.bloch_kinetic((5,), (1,), o.stacked_dk(np.array([[2 * np.pi]])), HBAR**2, HBAR).cpu().numpy()
    np.testing.assert_allclose(e[0], [0.0, 0.5, 2.0, 2.0, 0.5])
    assert np.all(e.imag == 0)


# ---------------------------------------------------------------------------------------- K5
@pytest.mark.parametrize(("shape", "repeat"), [*REF_GRID, ((3, 4, 5), (2, 3, 4)), ((16,), (7,))])
def test_bloch_permute(cuda, shape, repeat):
    torch, kernels = _gpu()
    m = int(np.prod(shape) * np.prod(repeat))
    rng = np.random.default_rng(1)
    v = rng.normal(size=(3, m)) + 1j * rng.normal(size=(3, m))
    vd = kernels.to_device(v, torch.complex128)
    fwd = kernels.bloch_permute(shape, repeat, vd, 0)
    np.testing.assert_array_equal(fwd.cpu().numpy(), o.bloch_from_supercell(v, shape, repeat))
    back = kernels.bloch_permute(shape, repeat, fwd, 1)
    np.testing.assert_array_equal(back.cpu().numpy(), v)
    np.testing.assert_array_equal(
        kernels.bloch_permute(shape, repeat, vd, 1).cpu().numpy(), o.bloch_into_supercell(v, shape, repeat)
    )


# ---------------------------------------------------------------------------------------- K4
@pytest.mark.parametrize(
    "dims",
    [(5,), (9,), (12,), (16,), (25,), (64,), (224,), (1024,), (6400,), (3 * 5 * 7 * 11,), (8192,), (131072,),
     (12, 15), (16, 20), (64, 48), (9, 16, 10), (28, 28, 28), (13 * 4,), (31 * 2,), (8192, 6), (6400, 3, 2)],
)
def test_fft_matches_numpy(cuda, dims):
    torch, kernels = _gpu()
    rng = np.random.default_rng(2)
    batch = 3 if np.prod(dims) < 100000 else 1
    x = rng.normal(size=(batch, *dims)) + 1j * rng.normal(size=(batch, *dims))
    xd = kernels.to_device(x, torch.complex128)
    axes = tuple(range(1, 1 + len(dims)))
    fwd = kernels.fft_c2c(xd, dims, -1).cpu().numpy()
    np.testing.assert_allclose(fwd, np.fft.fftn(x, axes=axes, norm="ortho"), rtol=0, atol=2e-13)
    bwd = kernels.fft_c2c(xd, dims, +1).cpu().numpy()
    np.testing.assert_allclose(bwd, np.fft.ifftn(x, axes=axes, norm="ortho"), rtol=0, atol=2e-13)
    # in place round trip
    y = xd.clone()
    kernels.fft_c2c(y, dims, -1, out=y)
    kernels.fft_c2c(y, dims, +1, out=y)
    np.testing.assert_allclose(y.cpu().numpy(), x, rtol=0, atol=2e-13)


def test_fft_axis(cuda):
    torch, kernels = _gpu()
    rng = np.random.default_rng(3)
    x = rng.normal(size=(4, 12, 10)) + 1j * rng.normal(size=(4, 12, 10))
    xd = kernels.to_device(x, torch.complex128)
    got = kernels.fft_axis(xd, 4, 12, 10, -1).cpu().numpy()
    np.testing.assert_allclose(got, np.fft.fft(x, axis=1, norm="ortho"), rtol=0, atol=1e-13)


# ---------------------------------------------------------------------------------------- K2
def _check_eigh(h, w, t, info):
    """north_star tolerances: eigenvalues rel 1e-10, residual < 1e-10, projectors."""
    assert np.all(info == 0), info
    norm = np.linalg.norm(h, axis=(1, 2))
    w_ref, t_ref = o.eigh_blocks(h)
    scale = np.abs(w_ref).max(axis=1, keepdims=True)
    assert np.all(np.diff(w, axis=1) >= 0)
    np.testing.assert_allclose(w, w_ref, rtol=1e-10, atol=0) if np.all(np.abs(w_ref) > 1e-3 * scale) else None
    assert np.max(np.abs(w - w_ref) / np.maximum(scale, 1e-300)) < 1e-10
    res, orth = o.eigh_residuals(h, w, t)
    assert res < 1e-10, res
    assert orth < 1e-10, orth
    for b in range(min(len(h), 8)):
        err = o.subspace_projector_error(w_ref[b], t[b], t_ref[b], norm[b])
        assert err < 1e-8, err


@pytest.mark.parametrize("n", [1, 2, 3, 5, 16, 33, 64, 100, 128, 200, 256])
def test_eigh_random_hermitian(cuda, n):
    torch, kernels = _gpu()
    rng = np.random.default_rng(0)
    batch = 5
    g = rng.normal(size=(batch, n, n)) + 1j * rng.normal(size=(batch, n, n))
    h = g + np.swapaxes(g.conj(), 1, 2)
    a = kernels.to_device(h, torch.complex128)
    w, t, info = kernels.eigh_batched(a)
    _check_eigh(h, w.cpu().numpy(), t.cpu().numpy(), info.cpu().numpy())


@pytest.mark.parametrize("n", [343, 512])
def test_eigh_large_blocks(cuda, n):
    torch, kernels = _gpu()
    rng = np.random.default_rng(4)
    g = rng.normal(size=(2, n, n)) + 1j * rng.normal(size=(2, n, n))
    h = g + np.swapaxes(g.conj(), 1, 2)
    a = kernels.to_device(h, torch.complex128)
    w, t, info = kernels.eigh_batched(a)
    _check_eigh(h, w.cpu().numpy(), t.cpu().numpy(), info.cpu().numpy())


@pytest.mark.parametrize("height", [0.0, 10.0])
def test_eigh_bloch_blocks_degenerate(cuda, height):
    """Free-particle blocks (cos_potential(meta, 0), in the reference tests) have exact +-k degeneracies."""
    torch, kernels = _gpu()
    dx = 2 * np.pi * np.eye(2)
    shape, repeat = (8, 8), (3, 4)
    comps = o.cos_potential_components(shape, height)
    hd = _build_gpu(dx, shape, comps, HBAR**2, repeat)
    h = hd.cpu().numpy()
    w, t, info = kernels.eigh_batched(hd)
    _check_eigh(h, w.cpu().numpy(), t.cpu().numpy(), info.cpu().numpy())


def test_eigh_chunked_workspace(cuda):
    """A workspace that only holds 3 matrices forces the chunk loop."""
    torch, kernels = _gpu()
    from slate_quantum_b200 import _lib

    n, batch = 48, 10
    rng = np.random.default_rng(5)
    g = rng.normal(size=(batch, n, n)) + 1j * rng.normal(size=(batch, n, n))
    h = g + np.swapaxes(g.conj(), 1, 2)
    a = kernels.to_device(h, torch.complex128)
    per3 = _lib.load().sqb_eigh_workspace_bytes(n, 3)
    ws = torch.empty(per3, dtype=torch.uint8, device="cuda")
    w, t, info = kernels.eigh_batched(a, workspace=ws)
    _check_eigh(h, w.cpu().numpy(), t.cpu().numpy(), info.cpu().numpy())


# ---------------------------------------------------------------------------------------- K3
@pytest.mark.parametrize(("n", "batch", "n_t"), [(5, 3, 7), (64, 4, 100), (100, 3, 65), (256, 2, 130), (343, 1, 64)])
def test_project_and_evolve(cuda, n, batch, n_t):
    torch, kernels = _gpu()
    rng = np.random.default_rng(6)
    g = rng.normal(size=(batch, n, n)) + 1j * rng.normal(size=(batch, n, n))
    h = g + np.swapaxes(g.conj(), 1, 2)
    w_ref, t_ref = o.eigh_blocks(h)
    psi = rng.normal(size=(batch, n)) + 1j * rng.normal(size=(batch, n))
    times = np.linspace(0.0, 5.0 * HBAR, n_t)
    td = kernels.to_device(t_ref, torch.complex128)
    wd = kernels.to_device(w_ref, torch.float64)
    c = kernels.project(td, kernels.to_device(psi, torch.complex128))
    c_ref = o.project(t_ref, psi)
    np.testing.assert_allclose(c.cpu().numpy(), c_ref, rtol=0, atol=1e-12)
    timed = kernels.to_device(times, torch.float64)
    a = kernels.evolve_eigen(wd.reshape(-1), c.reshape(-1), timed, HBAR).cpu().numpy()
    a_ref = o.evolve_eigenbasis(c_ref, w_ref, times, HBAR)
    np.testing.assert_allclose(a, a_ref, rtol=0, atol=1e-9)
    out = kernels.evolve_bloch(td, wd, c, timed, HBAR).cpu().numpy()
    out_ref = o.back_transform(t_ref, a_ref).reshape(n_t, batch * n)
    np.testing.assert_allclose(out, out_ref, rtol=0, atol=1e-9)
    # uniform-grid fast path (phase recurrence in the producer warps) must agree to the same tolerance
    dt = kernels.uniform_step(times)
    assert dt is not None
    out_u = kernels.evolve_bloch(td, wd, c, timed, HBAR, uniform_dt=dt).cpu().numpy()
    np.testing.assert_allclose(out_u, out_ref, rtol=0, atol=1e-9)
    assert kernels.uniform_step(times**2 / max(times[-1], 1e-300)) is None


def test_full_pipeline_vs_expm(cuda):
    """K1 -> K2 -> K4 -> K5 -> K3 -> K5 -> K4 against dense scipy expm of the supercell Hamiltonian."""
    import scipy.linalg

    torch, kernels = _gpu()
    dx = 2 * np.pi * np.eye(2)
    shape, repeat = (4, 3), (3, 4)
    big = tuple(n * r for n, r in zip(shape, repeat))
    n, n_k = int(np.prod(shape)), int(np.prod(repeat))
    comps = o.sin_potential_components(shape, 3.0, 0.3)
    hd = _build_gpu(dx, shape, comps, HBAR**2, repeat)
    w, t, info = kernels.eigh_batched(hd)
    assert int(info.abs().sum()) == 0
    rng = np.random.default_rng(7)
    psi = rng.normal(size=n * n_k) + 1j * rng.normal(size=n * n_k)
    psi /= np.linalg.norm(psi)
    times = np.linspace(0, 3 * HBAR, 9)
    psid = kernels.to_device(psi, torch.complex128)
    psik = kernels.fft_c2c(psid.reshape(1, -1), big, -1)
    psib = kernels.bloch_permute(shape, repeat, psik, 0).reshape(n_k, n)
    c = kernels.project(t, psib)
    out = kernels.evolve_bloch(t, w, c, kernels.to_device(times, torch.float64), HBAR)
    outk = kernels.bloch_permute(shape, repeat, out, 1)
    outx = kernels.fft_c2c(outk, big, +1).cpu().numpy()
    hk = o.supercell_hamiltonian(dx, shape, comps, HBAR**2, repeat)
    hx = o.momentum_matrix_to_position(hk, big)
    ref = np.array([scipy.linalg.expm(-1j * hx * tt / HBAR) @ psi for tt in times])
    np.testing.assert_allclose(outx, ref, rtol=0, atol=1e-9)


@pytest.mark.parametrize(
    ("shape", "repeat", "kind"),
    [((8,), (6,), "sin"), ((7,), (5,), "sin"), ((6, 4), (4, 3), "cos"), ((5, 4), (3, 4), "sin"), ((3, 4, 2), (4, 3, 2), "cos"),
     ((8, 8), (8, 4), "cos"), ((6, 4), (4, 3), "complex"), ((4, 4), (4, 2), "hex")],
)
def test_time_reversal_eigensolve(cuda, shape, repeat, kind):
    """BlochSolver.diagonalise with time_reversal = True: the blocks of half the k-grid are diagonalised, the rest take
    eigenvalues and permuted, conjugated eigenvectors from their partners (H_{-k}[pi g, pi g'] = conj(H_k[g, g'])).
    Against the oracle on EVERY block (eigenvalues, residual, orthogonality) and through the evolution; a table that is
    not Hermitian-symmetric and a non-orthogonal lattice must fall back to the full eigensolve."""
    torch, kernels = _gpu()
    from slate_quantum_b200.pipeline import BlochSolver

    nd = len(shape)
    dx = 2 * np.pi * np.eye(nd)
    if kind == "hex":
        dx = 2 * np.pi * np.array([[1.0, 0.0], [0.5, np.sqrt(3) / 2]])
        comps = o.fcc_potential_components(shape, 5.0)
    elif kind == "sin":
        comps = o.sin_potential_components(shape, 4.0, 0.2)
    else:
        comps = o.cos_potential_components(shape, 6.0)
    table = o.pad_components(shape, comps).astype(np.complex128)
    rng = np.random.default_rng(3)
    if kind == "complex":  # V(x) not real: the k / -k pairing does not hold (the blocks are not even Hermitian pairs)
        table = table + 0.3j * np.roll(table, 1, axis=0)
    solver = BlochSolver(shape, repeat, o.stacked_dk(dx), HBAR**2, HBAR)
    solver.time_reversal = True
    h = solver.build(kernels.to_device(table.ravel(), torch.complex128)).clone()
    blocks = h.cpu().numpy()
    applies = solver._time_reversal_applies()  # noqa: SLF001
    assert applies == (kind in ("sin", "cos"))
    if kind == "complex":
        return  # such blocks are not Hermitian: nothing to diagonalise, only the fallback decision is checked
    launches = kernels.LAUNCHES
    w, v = solver.diagonalise()
    assert int(solver.info.abs().sum()) == 0
    w_ref, t_ref = o.eigh_blocks(blocks)
    assert np.abs(w.cpu().numpy() - w_ref).max() < 1e-10 * np.abs(w_ref).max()
    res, orth = o.eigh_residuals(blocks, w.cpu().numpy(), v.cpu().numpy())
    assert res < 1e-10 and orth < 1e-10, (res, orth)
    del launches
    # ... and through the evolution to the position basis
    psi = rng.normal(size=solver.m) + 1j * rng.normal(size=solver.m)
    psi /= np.linalg.norm(psi)
    times = np.linspace(0, 3 * HBAR, 7)
    c = solver.project(kernels.to_device(psi, torch.complex128))
    pos = solver.evolve_position(c, kernels.to_device(times, torch.float64)).cpu().numpy()
    want = o.evolve_pipeline(psi, shape, repeat, w_ref, t_ref, times, HBAR)["position"]
    np.testing.assert_allclose(pos, want, rtol=0, atol=1e-9)


# ---------------------------------------------------------------------------------------- K4b cell FFT
@pytest.mark.parametrize(("shape", "repeat"), [((5,), (4,)), ((8,), (6,)), ((4, 3), (3, 4)), ((4, 4), (8, 2)),
                                                ((3, 4, 5), (2, 3, 4)), ((7, 7, 7), (2, 4, 2))])
def test_cell_fft_and_fold_equal_permute_plus_supercell_fft(cuda, shape, repeat):
    """fold(V) + cell FFT must equal K5 + supercell ifftn (the literal reference chain) on the same data."""
    torch, kernels = _gpu()
    n, n_k = int(np.prod(shape)), int(np.prod(repeat))
    big = tuple(a * b for a, b in zip(shape, repeat))
    rng = np.random.default_rng(8)
    v = rng.normal(size=(n_k, n, n)) + 1j * rng.normal(size=(n_k, n, n))  # any "transform"
    a = rng.normal(size=(3, n_k, n)) + 1j * rng.normal(size=(3, n_k, n))  # amplitudes per (t, block, e)
    bloch = np.einsum("tbe,bek->tbk", a, v).reshape(3, -1)
    want = o.momentum_to_position(o.bloch_into_supercell(bloch, shape, repeat), big)
    vd = kernels.to_device(v, torch.complex128)
    folded = kernels.fold_transform(shape, repeat, vd)
    cell = kernels.apply_transform(folded, kernels.to_device(a.reshape(3, -1), torch.complex128), 0)
    got = kernels.cell_fft(shape, repeat, cell, 1).cpu().numpy()
    np.testing.assert_allclose(got, want, rtol=0, atol=1e-12)
    # direction 0 is the inverse map
    back = kernels.cell_fft(shape, repeat, kernels.to_device(want, torch.complex128), 0)
    cell_ref = kernels.apply_transform(folded, kernels.to_device(a.reshape(3, -1), torch.complex128), 0)
    np.testing.assert_allclose(back.cpu().numpy(), cell_ref.cpu().numpy(), rtol=0, atol=1e-12)


@pytest.mark.parametrize(
    ("shape", "repeat"),
    [((16, 16), (3, 2)), ((8, 8), (2, 5)), ((7, 7, 7), (2, 1, 3)), ((5, 4), (4, 3)), ((16,), (6,)), ((3, 2, 8), (2, 2, 2)),
     ((32, 16), (2, 2)), ((12,), (5,))],
)
def test_fold_transform_against_numpy(cuda, shape, repeat):
    """sqb_fold_transform = in-cell ifftn(ortho) of every eigenvector x the block's Bloch twiddle
    exp(2 pi i sum_a q_a j_a / (N_a R_a)); cells whose axes are single radices take the fused one-pass kernel
    (fold_fused_kernel), the rest ((32, 16), (12,)) the per-axis passes - both against the same numpy expression."""
    torch, kernels = _gpu()
    n, n_k = int(np.prod(shape)), int(np.prod(repeat))
    rng = np.random.default_rng(n + n_k)
    v = _rand_c(rng, n_k, n, n)
    got = kernels.fold_transform(shape, repeat, kernels.to_device(v, torch.complex128)).cpu().numpy()
    nd = len(shape)
    cell = np.fft.ifftn(v.reshape(n_k, n, *shape), axes=tuple(range(2, 2 + nd)), norm="ortho").reshape(n_k, n, n)
    blocks = np.stack(np.unravel_index(np.arange(n_k), repeat), axis=-1)          # [n_k, nd]
    q = np.where(blocks < (np.asarray(repeat) + 1) // 2, blocks, blocks - np.asarray(repeat))
    j = np.stack(np.unravel_index(np.arange(n), shape), axis=-1)                  # [n, nd]
    phase = np.einsum("ba,ja->bj", q / (np.asarray(shape) * np.asarray(repeat)), j)
    want = cell * np.exp(2j * np.pi * phase)[:, None, :]
    np.testing.assert_allclose(got, want, rtol=0, atol=1e-13)


def test_fold_transform_block_range(cuda):
    torch, kernels = _gpu()
    shape, repeat = (4, 3), (3, 4)
    n, n_k = 12, 12
    rng = np.random.default_rng(9)
    v = rng.normal(size=(n_k, n, n)) + 1j * rng.normal(size=(n_k, n, n))
    vd = kernels.to_device(v, torch.complex128)
    full = kernels.fold_transform(shape, repeat, vd).cpu().numpy()
    part = kernels.fold_transform(shape, repeat, vd[5:9].contiguous(), block_begin=5).cpu().numpy()
    np.testing.assert_array_equal(part, full[5:9])


def test_evolve_position_pipeline(cuda):
    """BlochSolver.evolve_position (folded eigenvectors + cell FFT) against the oracle chain."""
    torch, kernels = _gpu()
    from slate_quantum_b200.pipeline import BlochSolver

    dx = 2 * np.pi * np.eye(2)
    shape, repeat = (6, 4), (4, 6)
    comps = o.sin_potential_components(shape, 4.0, 0.2)
    solver = BlochSolver(shape, repeat, o.stacked_dk(dx), HBAR**2, HBAR)
    solver.build(kernels.to_device(o.pad_components(shape, comps).ravel(), torch.complex128))
    blocks = solver.hamiltonian.cpu().numpy()
    solver.diagonalise()
    rng = np.random.default_rng(10)
    psi = rng.normal(size=solver.m) + 1j * rng.normal(size=solver.m)
    psi /= np.linalg.norm(psi)
    times = np.linspace(0, 5 * HBAR, 11)
    c = solver.project(kernels.to_device(psi, torch.complex128))
    got = solver.evolve_position(c, kernels.to_device(times, torch.float64)).cpu().numpy()
    w_ref, t_ref = o.eigh_blocks(blocks)
    want = o.evolve_pipeline(psi, shape, repeat, w_ref, t_ref, times, HBAR)["position"]
    np.testing.assert_allclose(got, want, rtol=0, atol=1e-9)


@pytest.mark.parametrize(("shape", "repeat", "ranks"), [((5,), (6,), 2), ((4, 3), (4, 3), 2), ((4, 4), (8, 2), 4),
                                                         ((3, 2, 3), (4, 3, 2), 2)])
def test_cell_fft_sharded_reads_rank_major_buffer(cuda, shape, repeat, ranks):
    """The all-to-all receive buffer [ranks, lead, n_k/ranks, N] must transform like the transposed copy."""
    torch, kernels = _gpu()
    n, n_k = int(np.prod(shape)), int(np.prod(repeat))
    lead = 3
    rng = np.random.default_rng(11)
    cell = rng.normal(size=(lead, n_k, n)) + 1j * rng.normal(size=(lead, n_k, n))
    want = kernels.cell_fft(shape, repeat, kernels.to_device(cell.reshape(lead, -1), torch.complex128), 1).cpu().numpy()
    per = n_k // ranks
    rank_major = np.ascontiguousarray(cell.reshape(lead, ranks, per, n).transpose(1, 0, 2, 3))
    got = kernels.cell_fft(shape, repeat, kernels.to_device(rank_major.reshape(-1), torch.complex128), 1, ranks=ranks)
    np.testing.assert_allclose(got.cpu().numpy(), want, rtol=0, atol=1e-13)


# ---------------------------------------------------------------------------------------- edge cases
def test_eigh_special_matrices(cuda):
    """Zero, identity, diagonal (incl. repeated entries), rank-one and 2x2 complex matrices."""
    torch, kernels = _gpu()
    n = 24
    rng = np.random.default_rng(12)
    u = rng.normal(size=n) + 1j * rng.normal(size=n)
    mats = [
        np.zeros((n, n), dtype=complex),
        np.eye(n, dtype=complex),
        np.diag(np.repeat(np.arange(n // 4), 4)).astype(complex),
        np.diag(rng.normal(size=n)).astype(complex),
        np.outer(u, u.conj()),
        np.diag(np.geomspace(1e-8, 1e8, n)).astype(complex),
    ]
    h = np.stack(mats)
    w, t, info = kernels.eigh_batched(kernels.to_device(h, torch.complex128))
    w, t, info = w.cpu().numpy(), t.cpu().numpy(), info.cpu().numpy()
    assert np.all(info == 0)
    w_ref = np.linalg.eigvalsh(h)
    scale = np.maximum(np.abs(w_ref).max(axis=1, keepdims=True), 1e-300)
    assert np.max(np.abs(w - w_ref) / scale) < 1e-12
    v = np.swapaxes(t, 1, 2)
    assert np.abs(np.swapaxes(v.conj(), 1, 2) @ v - np.eye(n)).max() < 1e-12
    res = np.abs(h @ v - v * w[:, None, :]).max(axis=(1, 2)) / scale[:, 0]
    assert res.max() < 1e-12
    h2 = np.array([[[1.0, 2 - 3j], [2 + 3j, -1.0]]])
    w2, t2, info2 = kernels.eigh_batched(kernels.to_device(h2, torch.complex128))
    np.testing.assert_allclose(w2.cpu().numpy()[0], np.linalg.eigvalsh(h2[0]), atol=1e-14)


# n_t = 513 / 1100: the uniform-grid kernel walks 8 time tiles of 64 rows per CTA (one full group + a ragged one)
@pytest.mark.parametrize(("n", "n_t"), [(7, 1), (9, 63), (130, 65), (64, 129), (40, 513), (70, 1100)])
def test_evolve_ragged_sizes(cuda, n, n_t):
    """Block sizes / time counts that are not multiples of the GEMM tiles, general and uniform-grid kernels."""
    torch, kernels = _gpu()
    rng = np.random.default_rng(13)
    batch = 3
    g = rng.normal(size=(batch, n, n)) + 1j * rng.normal(size=(batch, n, n))
    w_ref, t_ref = o.eigh_blocks(g + np.swapaxes(g.conj(), 1, 2))
    c = rng.normal(size=(batch, n)) + 1j * rng.normal(size=(batch, n))
    times = 0.3 * HBAR + np.arange(n_t) * 0.17 * HBAR
    want = o.back_transform(t_ref, o.evolve_eigenbasis(c, w_ref, times, HBAR)).reshape(n_t, batch * n)
    args = (kernels.to_device(t_ref, torch.complex128), kernels.to_device(w_ref, torch.float64),
            kernels.to_device(c, torch.complex128), kernels.to_device(times, torch.float64), HBAR)
    np.testing.assert_allclose(kernels.evolve_bloch(*args).cpu().numpy(), want, rtol=0, atol=1e-9)
    if n_t > 1:
        got = kernels.evolve_bloch(*args, uniform_dt=kernels.uniform_step(times)).cpu().numpy()
        np.testing.assert_allclose(got, want, rtol=0, atol=1e-9)


@pytest.mark.parametrize(
    ("n", "n_t", "spread"),
    [(32, 192, 1.0), (100, 512, 30.0), (256, 513, 3.0), (255, 1500, 500.0), (64, 2048, 0.0), (343, 512, 8.0),
     (512, 600, 2.0), (300, 1100, 0.0), (256, 2048, 1.0), (128, 1324, 0.5), (100, 4096, 0.3), (72, 1024, 1.2)],
)
def test_evolve_uniform_nufft_route(cuda, n, n_t, spread):
    """Uniform grids with <= 512 states per block and >= 192 rows take the NUFFT route (csrc/evolve_nufft.cuh): against
    the oracle's direct sum for ragged tiles, clustered / fully degenerate spectra (spread = 0: every source in ONE
    window), phases of hundreds of radians per step, blocks of more than 256 states (gathered in two passes of sorted sources),
    narrow bands (small spread: the kernel interleaves the rows of 2 / 4 / 8 tiles) and a column-slice output."""
    torch, kernels = _gpu()
    rng = np.random.default_rng(n + n_t)
    batch = 3
    assert kernels.evolve_uniform_route(n, batch, n_t) == "nufft"
    assert kernels.evolve_uniform_route(n, batch, 64) == "gemm"
    assert kernels.evolve_uniform_route(513, batch, n_t) == "gemm"
    g = rng.normal(size=(batch, n, n)) + 1j * rng.normal(size=(batch, n, n))
    _, t_ref = o.eigh_blocks(g + np.swapaxes(g.conj(), 1, 2))
    w_ref = np.sort(rng.uniform(-0.2, 1.0, size=(batch, n)) * spread, axis=1) + 0.7
    w_ref[0, : n // 2] = w_ref[0, 0]  # a degenerate band
    c = (rng.normal(size=(batch, n)) + 1j * rng.normal(size=(batch, n))) / np.sqrt(n)
    times = 0.3 * HBAR + np.arange(n_t) * 0.37 * HBAR
    want = o.back_transform(t_ref, o.evolve_eigenbasis(c, w_ref, times, HBAR)).reshape(n_t, batch * n)
    args = (kernels.to_device(t_ref, torch.complex128), kernels.to_device(w_ref, torch.float64),
            kernels.to_device(c, torch.complex128), kernels.to_device(times, torch.float64), HBAR)
    dt = kernels.uniform_step(times)
    got = kernels.evolve_bloch(*args, uniform_dt=dt).cpu().numpy()
    # the oracle's own phases carry |w t / hbar| * 1e-16 of rounding (up to ~3e-11 at spread = 500)
    tol = 1e-11 * max(1.0, spread / 10.0)
    np.testing.assert_allclose(got, want, rtol=0, atol=tol)
    # output as a column slice of a wider buffer (the multi-GPU segment layout)
    wide = torch.zeros((n_t, batch * n + 40), dtype=torch.complex128, device="cuda")
    kernels.evolve_bloch(*args, uniform_dt=dt, out=wide[:, 24 : 24 + batch * n])
    np.testing.assert_allclose(wide[:, 24 : 24 + batch * n].cpu().numpy(), want, rtol=0, atol=tol)
    assert float(wide[:, :24].abs().max()) == 0.0 and float(wide[:, 24 + batch * n :].abs().max()) == 0.0


def test_evolve_uniform_routes_share_one_workspace(cuda):
    """Time tiles of different length take different routes (host_tiles ramps 64, 128, ... rows): the GEMM route's
    Re+Im plane must be rebuilt when the NUFFT route served the first call on a new transform."""
    torch, kernels = _gpu()
    rng = np.random.default_rng(99)
    n, batch = 96, 4
    times = 0.1 * HBAR + np.arange(600) * 0.21 * HBAR
    dt = kernels.uniform_step(times)
    ws = torch.empty(kernels.evolve_uniform_workspace_bytes(n, batch, 600), dtype=torch.uint8, device="cuda")
    ws.fill_(0xFF)
    for trial in range(2):  # second trial: a NEW transform in the same workspace, first served by the NUFFT route
        g = rng.normal(size=(batch, n, n)) + 1j * rng.normal(size=(batch, n, n))
        w_ref, t_ref = o.eigh_blocks(g + np.swapaxes(g.conj(), 1, 2))
        c = (rng.normal(size=(batch, n)) + 1j * rng.normal(size=(batch, n))) / np.sqrt(n)
        want = o.back_transform(t_ref, o.evolve_eigenbasis(c, w_ref, times, HBAR)).reshape(600, batch * n)
        td, wd, cd = (kernels.to_device(t_ref, torch.complex128), kernels.to_device(w_ref, torch.float64),
                      kernels.to_device(c, torch.complex128))
        reuse = False
        for t0, t1 in ((0, 500), (500, 564), (564, 600)) if trial else ((0, 64), (64, 564), (564, 600)):
            tt = kernels.to_device(times[t0:t1], torch.float64)
            got = kernels.evolve_bloch(td, wd, cd, tt, HBAR, uniform_dt=dt, workspace=ws, reuse_plane=reuse)
            np.testing.assert_allclose(got.cpu().numpy(), want[t0:t1], rtol=0, atol=1e-10)
            reuse = True


@pytest.mark.parametrize("dims", [(11,), (13 * 17,), (19 * 2,), (23, 4), (29,), (31, 3), (1,), (2, 1, 3)])
def test_fft_generic_primes_and_unit_axes(cuda, dims):
    torch, kernels = _gpu()
    rng = np.random.default_rng(14)
    x = rng.normal(size=(2, *dims)) + 1j * rng.normal(size=(2, *dims))
    axes = tuple(range(1, 1 + len(dims)))
    got = kernels.fft_c2c(kernels.to_device(x, torch.complex128), dims, -1).cpu().numpy()
    np.testing.assert_allclose(got, np.fft.fftn(x, axes=axes, norm="ortho"), rtol=0, atol=1e-13)


@pytest.mark.parametrize("dims", [(37,), (101,), (74, 3), (2, 41, 4), (6, 43), (4099,), (2 * 2053,)])
@pytest.mark.parametrize("sign", [-1, 1])
def test_fft_lengths_with_large_prime_factors(cuda, dims, sign):
    """Axis lengths with a prime factor above 31 run as a chirp-z (Bluestein) convolution on the power-of-two passes
    (the reference takes any length through numpy): strided and contiguous axes, in place, and M = 8192 > the
    shared-memory line limit (4099, 4106 points: four-step passes inside the convolution)."""
    torch, kernels = _gpu()
    rng = np.random.default_rng(sum(dims))
    x = rng.normal(size=(2, *dims)) + 1j * rng.normal(size=(2, *dims))
    axes = tuple(range(1, 1 + len(dims)))
    want = (np.fft.fftn if sign < 0 else np.fft.ifftn)(x, axes=axes, norm="ortho")
    xd = kernels.to_device(x, torch.complex128)
    got = kernels.fft_c2c(xd, dims, sign).cpu().numpy()
    np.testing.assert_allclose(got, want, rtol=0, atol=2e-12)
    kernels.fft_c2c(xd, dims, sign, out=xd)
    np.testing.assert_allclose(xd.cpu().numpy(), want, rtol=0, atol=2e-12)


def test_operator_basis_change_on_a_37_point_axis(cuda):
    """The momentum <-> position conversion of the mirror on a prime axis (slate_core goes through numpy there)."""
    torch, kernels = _gpu()
    rng = np.random.default_rng(37)
    x = _rand_c(rng, 3, 37)
    got = kernels.fft_axis(kernels.to_device(x, torch.complex128), 3, 37, 1, +1).cpu().numpy()
    np.testing.assert_allclose(got.reshape(3, 37), np.fft.ifft(x, axis=1, norm="ortho"), rtol=0, atol=1e-12)


@pytest.mark.parametrize(("shape", "repeat"), [((4,), (1,)), ((3, 4), (1, 5)), ((3, 4), (5, 1)), ((2, 3, 4), (3, 1, 2))])
def test_unit_repeats(cuda, shape, repeat):
    """repeat = 1 along an axis: permutation, cell FFT and build must degrade gracefully."""
    torch, kernels = _gpu()
    n, n_k = int(np.prod(shape)), int(np.prod(repeat))
    big = tuple(a * b for a, b in zip(shape, repeat))
    rng = np.random.default_rng(15)
    v = rng.normal(size=(2, n * n_k)) + 1j * rng.normal(size=(2, n * n_k))
    vd = kernels.to_device(v, torch.complex128)
    np.testing.assert_array_equal(kernels.bloch_permute(shape, repeat, vd, 0).cpu().numpy(),
                                  o.bloch_from_supercell(v, shape, repeat))
    tr = rng.normal(size=(n_k, n, n)) + 1j * rng.normal(size=(n_k, n, n))
    a = rng.normal(size=(2, n_k, n)) + 1j * rng.normal(size=(2, n_k, n))
    bloch = np.einsum("tbe,bek->tbk", a, tr).reshape(2, -1)
    want = o.momentum_to_position(o.bloch_into_supercell(bloch, shape, repeat), big)
    folded = kernels.fold_transform(shape, repeat, kernels.to_device(tr, torch.complex128))
    cell = kernels.apply_transform(folded, kernels.to_device(a.reshape(2, -1), torch.complex128), 0)
    np.testing.assert_allclose(kernels.cell_fft(shape, repeat, cell, 1).cpu().numpy(), want, rtol=0, atol=1e-12)


# ---------------------------------------------------------------------------------------------
# divide-and-conquer tridiagonal stage of K2 on its own (csrc/eigh_dc.cuh; model: tests/dc_model.py)
def _run_tridiag(d, e):
    torch, kernels = _gpu()
    dd = kernels.to_device(np.ascontiguousarray(d), torch.float64)
    ee = kernels.to_device(np.ascontiguousarray(e), torch.float64)
    w, q, info = kernels.tridiag_eigh_batched(dd, ee)
    return w.cpu().numpy(), q.cpu().numpy(), info.cpu().numpy()


def test_tridiag_dc_hard_cases(cuda):
    from test_dc_model import CASES, check_tridiagonal

    for name in sorted(CASES):
        d, e = CASES[name]
        w, q, info = _run_tridiag(d[None, :], e[None, :])
        assert info[0] == 0, name
        check_tridiagonal(d, e, w[0], q[0], tol_scale=2.0)


@pytest.mark.parametrize("n", [1, 2, 7, 32, 33, 64, 100, 256, 343, 512])
def test_tridiag_dc_random_batches(cuda, n):
    from test_dc_model import check_tridiagonal

    rng = np.random.default_rng(n)
    batch = 5
    d = rng.normal(size=(batch, n))
    e = rng.normal(size=(batch, n))
    w, q, info = _run_tridiag(d, e)
    assert not info.any()
    for b in range(batch):
        check_tridiagonal(d[b], e[b], w[b], q[b], tol_scale=2.0)


def test_tridiag_dc_matches_model_eigenvalues(cuda):
    import dc_model

    rng = np.random.default_rng(11)
    d, e = rng.normal(size=(1, 200)), rng.normal(size=(1, 200))
    w, q, info = _run_tridiag(d, e)
    w_model, _ = dc_model.dc_eigh_tridiagonal(d[0], e[0])
    assert info[0] == 0
    assert np.abs(w[0] - w_model).max() < 1e-13 * np.abs(w_model).max()


def test_tridiag_dc_nan_input_is_flagged(cuda):
    d = np.ones((2, 70))
    e = np.full((2, 70), 0.5)
    d[1, 40] = np.nan
    w, q, info = _run_tridiag(d, e)
    assert info[0] == 0 and info[1] != 0


# ---------------------------------------------------------------------------------------------
# K6 observables (csrc/measure.cu) against the oracle restatement of operator/_measure.py
@pytest.mark.parametrize(("grid", "axis"), [((300,), 0), ((24, 36), 0), ((24, 36), 1), ((5, 7, 6), 1), ((4096,), 0), ((5000,), 0)])
def test_measure_moments_against_oracle(cuda, grid, axis):
    torch, kernels = _gpu()
    rng = np.random.default_rng(21)
    m = int(np.prod(grid))
    lead = 5
    states = rng.normal(size=(lead, m)) + 1j * rng.normal(size=(lead, m))
    states *= np.exp(-np.linspace(-3, 3, m) ** 2)[None, :]  # localised, un-normalised
    lengths = np.array([2 * np.pi, 3.0, 1.7])[: len(grid)]
    delta_x = np.diag(lengths)
    dev = kernels.to_device(states, torch.complex128)
    spacing = lengths[axis] / grid[axis]
    mom = kernels.measure_moments(dev, grid, axis, spacing, lengths[axis]).cpu().numpy()
    x = mom[:, 1] / mom[:, 0]
    np.testing.assert_allclose(x, o.measure_all_x(states, delta_x, grid, axis), rtol=1e-12, atol=1e-13)
    periodic = np.mod(np.arctan2(mom[:, 4], mom[:, 3]), 2 * np.pi) * lengths[axis] / (2 * np.pi)
    want_p = o.measure_all_periodic_x(states, delta_x, grid, axis)
    np.testing.assert_allclose(periodic, want_p, rtol=1e-11, atol=1e-12)
    off = kernels.to_device(-want_p, torch.float64)
    mom2 = kernels.measure_moments(dev, grid, axis, spacing, lengths[axis], off, True).cpu().numpy()
    np.testing.assert_allclose(mom2[:, 2] / mom2[:, 0], o.measure_all_variance_x(states, delta_x, grid, axis),
                               rtol=1e-10, atol=1e-13)
    # shifted, unwrapped x (measure.all_x(offset=...))
    mom3 = kernels.measure_moments(dev, grid, axis, spacing, lengths[axis],
                                   kernels.to_device(np.full(lead, 0.37), torch.float64), False).cpu().numpy()
    np.testing.assert_allclose(mom3[:, 1] / mom3[:, 0], o.measure_all_x(states, delta_x, grid, axis, offset=0.37),
                               rtol=1e-12, atol=1e-13)


def test_measure_position_eigenstates_exact(cuda):
    """reference tests/test_measure.py:8-15 through the CUDA kernel: <x> of |i> is exactly i dx."""
    torch, kernels = _gpu()
    n = 50
    dx = 2 * np.pi / n
    mom = kernels.measure_moments(kernels.to_device(np.eye(n, dtype=np.complex128), torch.complex128), (n,), 0, dx,
                                  2 * np.pi).cpu().numpy()
    np.testing.assert_array_equal(mom[:, 1] / mom[:, 0], np.arange(n) * dx)


def test_measure_on_evolved_bloch_states(cuda):
    """K1 -> K2 -> K3 -> cell FFT -> K6: observables of the evolved wavepacket against the oracle chain."""
    torch, kernels = _gpu()
    from slate_quantum_b200.pipeline import BlochSolver

    shape, repeat = (8, 8), (4, 3)
    big = tuple(s * r for s, r in zip(shape, repeat))
    dx = 2 * np.pi * np.eye(2)
    big_dx = dx * np.array(repeat)[:, None]
    comps = o.cos_potential_components(shape, 10.0)
    solver = BlochSolver(shape, repeat, o.stacked_dk(dx), HBAR**2, HBAR)
    solver.build(kernels.to_device(o.pad_components(shape, comps).ravel(), torch.complex128))
    solver.diagonalise()
    psi = o.coherent_state(big_dx, big, (np.pi * repeat[0], np.pi * repeat[1]), (0.0, 0.3), (2.0, 2.5))
    times = np.linspace(0, 6 * HBAR, 24)
    c = solver.project(kernels.to_device(psi, torch.complex128))
    pos = solver.bloch_to_position(solver.evolve_bloch(c, kernels.to_device(times, torch.float64)))
    blocks = o.bloch_blocks(dx, shape, comps, HBAR**2, repeat)
    w_ref, t_ref = o.eigh_blocks(blocks)
    ref = o.evolve_pipeline(psi, shape, repeat, w_ref, t_ref, times, HBAR)["position"]
    for axis in (0, 1):
        length = big_dx[axis, axis]
        mom = kernels.measure_moments(pos.reshape(len(times), -1).contiguous(), big, axis, length / big[axis],
                                      length).cpu().numpy()
        np.testing.assert_allclose(mom[:, 0], 1.0, atol=1e-9)  # unitary evolution keeps the norm
        got = np.mod(np.arctan2(mom[:, 4], mom[:, 3]), 2 * np.pi) * length / (2 * np.pi)
        np.testing.assert_allclose(got, o.measure_all_periodic_x(ref, big_dx, big, axis), atol=1e-8)
        np.testing.assert_allclose(mom[:, 1] / mom[:, 0], o.measure_all_x(ref, big_dx, big, axis), atol=1e-8)


def test_eigensystem_shard_dump_and_resume(cuda, tmp_path):
    """(E, V) shard dump: a fresh solver resumed from the shard evolves identically, and a
    shard of a different problem / a corrupted shard is refused."""
    torch, kernels = _gpu()
    from slate_quantum_b200.pipeline import BlochSolver

    shape, repeat = (6, 5), (3, 4)
    dx = 2 * np.pi * np.eye(2)
    comps = o.cos_potential_components(shape, 7.0)
    table = kernels.to_device(o.pad_components(shape, comps).ravel(), torch.complex128)

    def make(mass=HBAR**2):
        return BlochSolver(shape, repeat, o.stacked_dk(dx), mass, HBAR, block_begin=2, block_count=7)

    a = make()
    a.build(table)
    a.diagonalise()
    a.save_eigensystem(str(tmp_path), tag="rank0")
    b = make()
    b.load_eigensystem(str(tmp_path), tag="rank0")
    rng = np.random.default_rng(5)
    c = kernels.to_device(rng.normal(size=(7, a.n)) + 1j * rng.normal(size=(7, a.n)), torch.complex128)
    times = kernels.to_device(np.linspace(0, 3 * HBAR, 20), torch.float64)
    np.testing.assert_array_equal(a.evolve_bloch(c, times).cpu().numpy(), b.evolve_bloch(c, times).cpu().numpy())
    with pytest.raises(ValueError, match="different problem"):
        make(mass=2 * HBAR**2).load_eigensystem(str(tmp_path), tag="rank0")
    w = np.load(tmp_path / "rank0.eigenvalues.npy")
    w[0, 0] += 1.0
    np.save(tmp_path / "rank0.eigenvalues.npy", w)
    with pytest.raises(ValueError, match="checksum"):
        make().load_eigensystem(str(tmp_path), tag="rank0")


# ---------------------------------------------------------------------------------------------
# BASELINE.json's full sizes through size-independent properties (the oracle cannot run these in seconds)
@pytest.mark.parametrize("name", ["cfg2", "cfg3"])
def test_full_size_properties(cuda, name):
    """cfg2 (1024 x N=128) / cfg3 (4096 x N=256): trace and ordering of every block's spectrum, residual and
    orthogonality of sampled blocks, t = 0 round trip position -> Bloch -> eigenbasis -> position, unitarity
    (norm of every evolved state), and the observables' normalisation."""
    torch, kernels = _gpu()
    from slate_quantum_b200.configs import CONFIGS
    from slate_quantum_b200.pipeline import BlochSolver

    cfg = CONFIGS[name]
    n, n_k = cfg.n, cfg.n_k
    solver = BlochSolver(cfg.shape, cfg.repeat, cfg.dk(), cfg.mass(), HBAR)
    table = kernels.to_device(cfg.potential_table().ravel(), torch.complex128)
    h = solver.build(table)
    trace = torch.diagonal(h, dim1=1, dim2=2).real.sum(dim=1).clone()
    h_norm = torch.linalg.matrix_norm(h).clone()
    rng = np.random.default_rng(31)
    sample = torch.as_tensor(np.sort(rng.choice(n_k, size=48, replace=False)), device="cuda")
    h_sample = h[sample].clone()
    w, v = solver.diagonalise()
    assert int(solver.info.abs().sum()) == 0
    # every block: ascending eigenvalues whose sum is the trace of the block
    assert bool((w[:, 1:] >= w[:, :-1]).all())
    assert float(((w.sum(dim=1) - trace).abs() / h_norm).max()) < 1e-12
    # sampled blocks: residual and orthogonality (rows of v = eigenvectors)
    vs = v[sample]
    vcols = vs.transpose(1, 2)
    resid = torch.linalg.matrix_norm(h_sample @ vcols - vcols * w[sample][:, None, :]) / h_norm[sample]
    assert float(resid.max()) < 1e-10
    eye = torch.eye(n, dtype=torch.complex128, device="cuda")
    assert float(torch.linalg.matrix_norm(vs.conj() @ vcols - eye).max()) < 1e-10
    del h_sample, vs, vcols
    # evolution: t = 0 reproduces psi0, every later state has the same norm
    psi0 = kernels.to_device(cfg.initial_state(), torch.complex128)
    c = solver.project(psi0)
    n_t = 64
    times = kernels.to_device(cfg.times()[:n_t].copy(), torch.float64)
    assert float(times[0]) == 0.0
    pos = solver.evolve_position(c, times, uniform_dt=kernels.uniform_step(cfg.times()[:n_t]))
    assert float((pos[0] - psi0.reshape(-1)).abs().max()) < 1e-9
    norms = torch.linalg.vector_norm(pos, dim=1)
    assert float((norms - norms[0]).abs().max()) < 1e-9
    # general (non-uniform) kernel on a few of the same times agrees with the uniform-grid 3M kernel
    few = solver.evolve_position(c, times[:8].contiguous())
    assert float((few - pos[:8]).abs().max()) < 1e-9
    big = cfg.supercell
    length = float(np.linalg.norm(cfg.vectors()[0])) * cfg.repeat[0]
    mom = kernels.measure_moments(pos, big, 0, length / big[0], length)
    assert float((mom[:, 0] - norms**2).abs().max()) < 1e-9


@pytest.mark.parametrize(("name", "count"), [("cfg4", 96), ("cfg5", 160)])
def test_sharded_config_block_range_properties(cuda, name, count):
    """cfg4 (hexagonal, N=512) / cfg5 (cubic, N=343) need k-sharding over GPUs: a contiguous block range (what one
    rank owns) is built and diagonalised at full block size and checked through trace, ordering, residual and
    orthogonality, and against numpy on three of its blocks."""
    torch, kernels = _gpu()
    from slate_quantum_b200.configs import CONFIGS
    from slate_quantum_b200.pipeline import BlochSolver

    cfg = CONFIGS[name]
    n = cfg.n
    begin = cfg.n_k // 3
    solver = BlochSolver(cfg.shape, cfg.repeat, cfg.dk(), cfg.mass(), HBAR, block_begin=begin, block_count=count)
    h = solver.build(kernels.to_device(cfg.potential_table().ravel(), torch.complex128)).clone()
    w, v = solver.diagonalise()
    assert int(solver.info.abs().sum()) == 0
    h_norm = torch.linalg.matrix_norm(h)
    trace = torch.diagonal(h, dim1=1, dim2=2).real.sum(dim=1)
    assert bool((w[:, 1:] >= w[:, :-1]).all())
    assert float(((w.sum(dim=1) - trace).abs() / h_norm).max()) < 1e-12
    vcols = v.transpose(1, 2)
    resid = torch.linalg.matrix_norm(h @ vcols - vcols * w[:, None, :]) / h_norm
    assert float(resid.max()) < 1e-10
    eye = torch.eye(n, dtype=torch.complex128, device="cuda")
    assert float(torch.linalg.matrix_norm(v.conj() @ vcols - eye).max()) < 1e-10
    for b in (0, count // 2, count - 1):
        w_ref = np.linalg.eigvalsh(h[b].cpu().numpy())
        np.testing.assert_allclose(w[b].cpu().numpy(), w_ref, rtol=0, atol=1e-10 * np.abs(w_ref).max())


# ---------------------------------------------------------------------------------------- K7
def _rand_c(rng, *shape):
    return rng.normal(size=shape) + 1j * rng.normal(size=shape)


@pytest.mark.parametrize(("m", "n", "k"), [(1, 1, 1), (5, 5, 5), (64, 64, 64), (70, 33, 129), (1, 257, 64),
                                            (130, 1, 9), (300, 300, 300), (96, 200, 7)])
@pytest.mark.parametrize("batch", [None, 3])
def test_zgemm_against_numpy(cuda, m, n, k, batch):
    """sqb_zgemm_batched against the oracle's operator_matmul on ragged sizes (tile edges in m, n and k)."""
    torch, kernels = _gpu()
    rng = np.random.default_rng(m * 1000 + n * 10 + k)
    lead = () if batch is None else (batch,)
    a, b = _rand_c(rng, *lead, m, k), _rand_c(rng, *lead, k, n)
    want = o.operator_matmul(a, b)
    got = kernels.zgemm(kernels.to_device(a, torch.complex128), kernels.to_device(b, torch.complex128))
    assert got.shape == want.shape
    np.testing.assert_allclose(got.cpu().numpy(), want, rtol=0, atol=1e-13 * max(1, k))


def test_zgemm_strided_conjugated_and_broadcast_operands(cuda):
    """Transposes / daggers / conjugates are stride swaps + a flag, a 2-d operand is shared by the batch, and
    alpha / beta accumulate in place: out = alpha * op(A) op(B) + beta * out."""
    torch, kernels = _gpu()
    rng = np.random.default_rng(77)
    m, n, k, batch = 75, 50, 91, 4
    a, b = _rand_c(rng, batch, k, m), _rand_c(rng, n, k)  # stored transposed
    c0 = _rand_c(rng, batch, m, n)
    ad, bd = kernels.to_device(a, torch.complex128), kernels.to_device(b, torch.complex128)
    alpha, beta = 0.3 - 1.1j, -0.5 + 0.25j
    cases = {
        "T,T": (ad.mT, bd.mT, np.swapaxes(a, 1, 2), b.T),
        "H,T": (ad.mH, bd.mT, np.conj(np.swapaxes(a, 1, 2)), b.T),
        "T,H": (ad.mT, bd.mH, np.swapaxes(a, 1, 2), b.conj().T),
        "H,H": (ad.mH, bd.mH, np.conj(np.swapaxes(a, 1, 2)), b.conj().T),
    }
    for name, (xa, xb, na, nb) in cases.items():
        out = kernels.to_device(c0, torch.complex128)
        kernels.zgemm(xa, xb, alpha=alpha, beta=beta, out=out)
        want = alpha * (na @ nb) + beta * c0
        np.testing.assert_allclose(out.cpu().numpy(), want, rtol=0, atol=2e-12, err_msg=name)
    # sliced (non-compact) operands: a column block of a wider matrix
    wide = _rand_c(rng, 40, 200)
    wd = kernels.to_device(wide, torch.complex128)
    got = kernels.zgemm(wd[:, 30:101], wd[:, 90:161].mH)
    np.testing.assert_allclose(got.cpu().numpy(), wide[:, 30:101] @ wide[:, 90:161].conj().T, rtol=0, atol=1e-11)
    # k = 0: out = beta * out
    out = kernels.to_device(c0[0], torch.complex128)
    kernels.zgemm(ad[0, :0].mT, bd[:, :0].mT, alpha=1.0, beta=2.0, out=out)
    np.testing.assert_allclose(out.cpu().numpy(), 2 * c0[0], rtol=0, atol=0)
    with pytest.raises(ValueError):
        kernels.zgemm(ad[0], bd)  # inner dimensions differ


def test_zgemm_commutator_is_exactly_zero_for_equal_operands(cuda):
    """reference tests/test_operator.py:248-259 asserts array_equal(commute(A, A), 0): AB and BA must be
    accumulated in the same order so that the in-place subtraction cancels bit for bit."""
    torch, kernels = _gpu()
    a = kernels.to_device(_rand_c(np.random.default_rng(3), 150, 150), torch.complex128)
    out = kernels.zgemm(a, a)
    kernels.zgemm(a, a, alpha=-1.0, beta=1.0, out=out)
    assert float(out.abs().max().item()) == 0.0


def test_rowwise_reductions_against_oracle(cuda):
    """sqb_rowwise_vdot / sqb_abs2 / sqb_normalize_rows vs the oracle's inner_product_each, occupations,
    normalize_each (state/_state.py:208-308)."""
    torch, kernels = _gpu()
    rng = np.random.default_rng(5)
    for rows, length in [(1, 1), (3, 5), (7, 1000), (2, 70001)]:
        x, y = _rand_c(rng, rows, length), _rand_c(rng, rows, length)
        xd, yd = kernels.to_device(x, torch.complex128), kernels.to_device(y, torch.complex128)
        scale = np.sqrt(length)
        np.testing.assert_allclose(kernels.rowwise_vdot(xd, yd).cpu().numpy(), o.inner_product_each(x, y),
                                   rtol=0, atol=1e-13 * length)
        np.testing.assert_allclose(kernels.abs2(xd).cpu().numpy(), o.occupations(x), rtol=1e-15, atol=0)
        np.testing.assert_allclose(kernels.normalize_rows(xd).cpu().numpy(), o.normalize_each(x), rtol=0,
                                   atol=1e-15 * scale)
    # strided rows (a column block)
    wide = _rand_c(rng, 6, 300)
    wd = kernels.to_device(wide, torch.complex128)
    np.testing.assert_allclose(kernels.rowwise_vdot(wd[:, 10:150], wd[:, 150:290]).cpu().numpy(),
                               o.inner_product_each(wide[:, 10:150], wide[:, 150:290]), rtol=0, atol=1e-12)


def test_rowwise_weighted_abs2(cuda):
    """sqb_rowwise_weighted_abs2 = the oracle's expectation_of_each of a diagonal operator."""
    torch, kernels = _gpu()
    rng = np.random.default_rng(6)
    for rows, length in [(1, 1), (4, 37), (3, 5000)]:
        x, w = _rand_c(rng, rows, length), _rand_c(rng, length)
        got = kernels.rowwise_weighted_abs2(kernels.to_device(x, torch.complex128), kernels.to_device(w, torch.complex128))
        np.testing.assert_allclose(got.cpu().numpy(), o.expectation_of_each(np.diag(w), x), rtol=0, atol=1e-13 * length)
    with pytest.raises(ValueError):
        kernels.rowwise_weighted_abs2(kernels.to_device(x, torch.complex128), kernels.to_device(w[:-1], torch.complex128))


@pytest.mark.parametrize("length", [256, 512])
@pytest.mark.parametrize(("outer", "inner"), [(1, 3), (2, 24), (3, 16), (1, 37), (2, 1)])
@pytest.mark.parametrize("sign", [-1, 1])
def test_fft_axis_long_strided_lines(cuda, length, outer, inner, sign):
    """Three-stage power-of-two lines with adjacent lines in memory (the across-block cell FFT of a 256- or
    512-wide k-grid) take the one-buffer variant with an in-place middle stage; ragged line counts included."""
    torch, kernels = _gpu()
    rng = np.random.default_rng(length + 7 * outer + inner)
    x = _rand_c(rng, outer, length, inner)
    got = kernels.fft_axis(kernels.to_device(x, torch.complex128), outer, length, inner, sign).cpu().numpy()
    want = (np.fft.fft if sign < 0 else np.fft.ifft)(x, axis=1, norm="ortho")
    np.testing.assert_allclose(got.reshape(x.shape), want, rtol=0, atol=1e-13)


@pytest.mark.parametrize("length", [128, 256, 1024, 2048, 4096])
@pytest.mark.parametrize(("outer", "inner"), [(2, 1), (1, 40), (3, 33)])
@pytest.mark.parametrize("sign", [-1, 1])
def test_fft_axis_radix16_lengths(cuda, length, outer, inner, sign):
    """Power-of-two lines of 128 points and more lead with radix-16 stages (dft_small<16>: 128 = 16*8 and
    256 = 16*16 in two stages, 1024 .. 4096 in three), contiguous lines and adjacent strided lines (up to 32 per CTA,
    ragged counts included)."""
    torch, kernels = _gpu()
    rng = np.random.default_rng(length + 5 * outer + inner)
    x = _rand_c(rng, outer, length, inner)
    got = kernels.fft_axis(kernels.to_device(x, torch.complex128), outer, length, inner, sign).cpu().numpy()
    want = (np.fft.fft if sign < 0 else np.fft.ifft)(x, axis=1, norm="ortho")
    np.testing.assert_allclose(got.reshape(x.shape), want, rtol=0, atol=2e-13)


@pytest.mark.parametrize("shape", [(16, 16), (2, 8), (3, 24)])
def test_cell_fft_64x64_block_grid_cluster_kernel(cuda, shape):
    """64 x 64 block grids (cfg3) take the one-pass cluster kernel (cell_fft2_cluster_kernel: axis 1 per CTA, axis 0
    across the 8 CTAs of a cluster through distributed shared memory).  Against numpy: ifft2(ortho) over the two block
    axes of the cell layout [t, b0, b1, j0, j1], scattered to x_a = c_a N_a + j_a; 3 states (ragged against nothing:
    every (state, column group) is one cluster)."""
    torch, kernels = _gpu()
    repeat = (64, 64)
    n = shape[0] * shape[1]
    rng = np.random.default_rng(shape[0])
    x = _rand_c(rng, 3, 64, 64, shape[0], shape[1])
    want = np.fft.ifft2(x, axes=(1, 2), norm="ortho").transpose(0, 1, 3, 2, 4).reshape(3, -1)
    xd = kernels.to_device(x.reshape(3, -1), torch.complex128)
    got = kernels.cell_fft(shape, repeat, xd, 1).cpu().numpy()
    np.testing.assert_allclose(got, want, rtol=0, atol=1e-13)
    assert got.shape == (3, 4096 * n)


@pytest.mark.parametrize(("shape", "repeat"), [((2, 2), (512, 3)), ((3,), (256,)), ((2, 2), (4, 512)), ((2, 2), (128, 4))])
def test_cell_fft_long_block_axis(cuda, shape, repeat):
    """cell FFT across 256 / 512 blocks (8-GPU weak-scaling k-grids) against the literal supercell chain."""
    torch, kernels = _gpu()
    n, n_k = int(np.prod(shape)), int(np.prod(repeat))
    big = tuple(a * b for a, b in zip(shape, repeat))
    rng = np.random.default_rng(12)
    v = _rand_c(rng, n_k, n, n)
    a = _rand_c(rng, 2, n_k, n)
    bloch = np.einsum("tbe,bek->tbk", a, v).reshape(2, -1)
    want = o.momentum_to_position(o.bloch_into_supercell(bloch, shape, repeat), big)
    folded = kernels.fold_transform(shape, repeat, kernels.to_device(v, torch.complex128))
    cell = kernels.apply_transform(folded, kernels.to_device(a.reshape(2, -1), torch.complex128), 0)
    got = kernels.cell_fft(shape, repeat, cell, 1).cpu().numpy()
    np.testing.assert_allclose(got, want, rtol=0, atol=1e-12)


# ---------------------------------------------------------------------------------------- round 2 additions
@pytest.mark.parametrize("name", ["cfg2", "cfg3"])
def test_full_size_sampled_blocks_against_the_oracle(cuda, name):
    """cfg2 / cfg3 at BASELINE.json's FULL sizes: the position-basis states of the first, a middle and the LAST time
    rows of the full time grid (T = 1024 / 4096), from the uniform-grid 3M kernel the bench times AND from the
    general kernel, against the CPU oracle on sampled k-blocks at 1e-9 (the evolution is per block, so the oracle
    checks any time row of the full configuration in milliseconds per block); eigenvalues of the sampled blocks at
    rel 1e-10 and their eigenvectors through degenerate-subspace projectors."""
    torch, kernels = _gpu()
    from slate_quantum_b200.configs import CONFIGS
    from slate_quantum_b200.pipeline import BlochSolver

    cfg = CONFIGS[name]
    T = cfg.n_times
    solver = BlochSolver(cfg.shape, cfg.repeat, cfg.dk(), cfg.mass(), HBAR)
    solver.build(kernels.to_device(cfg.potential_table().ravel(), torch.complex128))
    w, v = solver.diagonalise()
    assert int(solver.info.abs().sum()) == 0
    psi_np = cfg.initial_state()
    c = solver.project(kernels.to_device(psi_np, torch.complex128))
    times_np = cfg.times()
    dt = kernels.uniform_step(times_np)
    assert dt is not None
    comps = cfg.potential_components()
    blocks = sorted({0, cfg.n_k // 3, cfg.n_k // 2 + 1, cfg.n_k - 1})
    # the 3M kernel derives row r of a tile from tilebase * step8^r * pow8: check rows deep inside LONG tiles
    tile = 2048 if T >= 2048 else T
    want_rows = sorted({0, tile - 1, T // 2 - 1, T - 1})
    got = {}
    states = torch.empty((tile, solver.m), dtype=torch.complex128, device="cuda")
    pos = torch.empty_like(states)
    for t0 in range(0, T, tile):
        tt = min(tile, T - t0)
        solver.evolve_cell(c, kernels.to_device(times_np[t0 : t0 + tt], torch.float64), out=states[:tt], uniform_dt=dt)
        solver.cell_to_position(states[:tt], out=pos[:tt])
        for r in want_rows:
            if t0 <= r < t0 + tt:
                got[r] = pos[r - t0].cpu().numpy()
    err = o.sampled_block_parity(cfg.vectors(), cfg.shape, comps, cfg.mass(), cfg.repeat, psi_np, times_np[want_rows],
                                 np.stack([got[r] for r in want_rows]), blocks, HBAR)
    assert err < 1e-9, err
    # the general (non-uniform) kernel on the same rows
    t_sel = kernels.to_device(times_np[want_rows], torch.float64)
    general = solver.evolve_position(c, t_sel).cpu().numpy()
    err = o.sampled_block_parity(cfg.vectors(), cfg.shape, comps, cfg.mass(), cfg.repeat, psi_np, times_np[want_rows],
                                 general, blocks, HBAR)
    assert err < 1e-9, err
    # the k-sharded (Bloch momentum) output of the same rows, uniform kernel, long tile
    bl = solver.evolve_bloch(c, kernels.to_device(times_np[T - tile :], torch.float64), out=states[:tile], uniform_dt=dt)
    err = o.sampled_block_parity(cfg.vectors(), cfg.shape, comps, cfg.mass(), cfg.repeat, psi_np, times_np[[T - 1]],
                                 bl[tile - 1 : tile].cpu().numpy(), blocks, HBAR, basis="bloch")
    assert err < 1e-9, err
    for b in blocks:
        h = o.bloch_blocks(cfg.vectors(), cfg.shape, comps, cfg.mass(), cfg.repeat, block_begin=b, block_count=1)
        w_ref, t_ref = o.eigh_blocks(h)
        np.testing.assert_allclose(w[b].cpu().numpy(), w_ref[0], rtol=0, atol=1e-10 * np.abs(w_ref).max())
        assert o.subspace_projector_error(w_ref[0], t_ref[0], v[b].cpu().numpy(), float(np.linalg.norm(h[0]))) < 1e-6


@pytest.mark.parametrize(("name", "count"), [("cfg5", 64), ("cfg4", 24)])
def test_sharded_config_block_range_evolution_against_the_oracle(cuda, name, count):
    """cfg5 (cubic, N = 343, T = 512) / cfg4 (hexagonal, N = 512): one rank's contiguous block range at full block
    size is built, diagonalised and evolved over the config's FULL time grid (k-sharded Bloch-momentum output, what a
    rank of the 8-GPU run produces); first / last rows of sampled blocks against the CPU oracle at 1e-9."""
    torch, kernels = _gpu()
    from slate_quantum_b200.configs import CONFIGS
    from slate_quantum_b200.pipeline import BlochSolver

    cfg = CONFIGS[name]
    begin = 5 * cfg.n_k // 8  # a block range in the middle of another rank's share
    solver = BlochSolver(cfg.shape, cfg.repeat, cfg.dk(), cfg.mass(), HBAR, block_begin=begin, block_count=count)
    solver.build(kernels.to_device(cfg.potential_table().ravel(), torch.complex128))
    solver.diagonalise()
    assert int(solver.info.abs().sum()) == 0
    psi_np = cfg.initial_state()
    c = solver.project(kernels.to_device(psi_np, torch.complex128))
    times_np = cfg.times()
    out = solver.evolve_bloch(c, kernels.to_device(times_np, torch.float64), uniform_dt=kernels.uniform_step(times_np))
    rows = [0, cfg.n_times - 1]
    blocks = sorted({begin, begin + count // 2, begin + count - 1})
    want = o.evolve_sampled_blocks(cfg.vectors(), cfg.shape, cfg.potential_components(), cfg.mass(), cfg.repeat, psi_np,
                                   times_np[rows], blocks, HBAR)
    got = out[rows].cpu().numpy().reshape(len(rows), count, cfg.n)[:, [b - begin for b in blocks], :]
    assert np.abs(got - want).max() < 1e-9


@pytest.mark.parametrize("case", ["sq1d", "sq2d", "hex2d", "cub3d"])
def test_golden_fixtures_against_the_cuda_path(cuda, case):
    """tests/golden/bloch_golden.npz (incl. the hexagonal quirk-Q1 case) against the CUDA path: blocks, eigenvalues,
    eigenbasis amplitudes (up to each eigenvector's phase) and evolved position-basis states."""
    from pathlib import Path

    torch, kernels = _gpu()
    from slate_quantum_b200.pipeline import BlochSolver

    g = np.load(Path(__file__).parent / "golden" / "bloch_golden.npz")
    shape, repeat = tuple(int(x) for x in g[f"{case}/shape"]), tuple(int(x) for x in g[f"{case}/repeat"])
    dx = g[f"{case}/dx"]
    solver = BlochSolver(shape, repeat, o.stacked_dk(dx), HBAR**2, HBAR)
    table = kernels.to_device(o.pad_components(shape, g[f"{case}/comps"]).ravel(), torch.complex128)
    h = solver.build(table).clone()
    np.testing.assert_allclose(h.cpu().numpy(), g[f"{case}/blocks"], atol=1e-13)
    w, _ = solver.diagonalise()
    w_ref = g[f"{case}/eigenvalues"]
    np.testing.assert_allclose(w.cpu().numpy(), w_ref, rtol=0, atol=1e-10 * np.abs(w_ref).max())
    psi = kernels.to_device(g[f"{case}/psi0"], torch.complex128)
    times = kernels.to_device(g[f"{case}/times"], torch.float64)
    c = solver.project(psi)
    amplitudes = solver.evolve_eigen(c, times).cpu().numpy()
    # eigenvector phases are a gauge choice: |a_n(t)| is not; within degenerate clusters only the cluster norm is
    np.testing.assert_allclose(np.linalg.norm(amplitudes, axis=1), np.linalg.norm(g[f"{case}/eigen_amplitudes"], axis=1),
                               atol=1e-10)
    pos = solver.bloch_to_position(solver.evolve_bloch(c, times)).cpu().numpy()
    np.testing.assert_allclose(pos, g[f"{case}/position"], atol=1e-9)
    pos2 = solver.evolve_position(c, times, uniform_dt=kernels.uniform_step(g[f"{case}/times"])).cpu().numpy()
    np.testing.assert_allclose(pos2, g[f"{case}/position"], atol=1e-9)
    perm = kernels.bloch_permute(shape, repeat, kernels.to_device(np.arange(solver.m).astype(np.complex128), torch.complex128), 0)
    np.testing.assert_array_equal(perm.cpu().numpy().real.astype(np.int64), g[f"{case}/permutation"])


@pytest.mark.parametrize("n", [600, 1024])
def test_eigh_dense_above_the_single_sm_limit(cuda, n):
    """One dense Hermitian matrix with n > 512: two-sided block Jacobi over batched K2
    sub-problems and K7 GEMMs (core/jacobi.py), same contract as the kernel path - ascending eigenvalues at rel
    1e-10, residual and orthogonality below 1e-10, rows = eigenvectors."""
    torch, kernels = _gpu()
    rng = np.random.default_rng(n)
    g = rng.normal(size=(n, n)) + 1j * rng.normal(size=(n, n))
    h = g + g.conj().T
    w, t, info = kernels.eigh_batched(kernels.to_device(h[None], torch.complex128))
    assert int(info.abs().sum()) == 0
    _check_eigh(h[None], w.cpu().numpy(), t.cpu().numpy(), info.cpu().numpy())


def test_round_robin_pairs_cover_every_pair_once():
    from slate_quantum_b200.core.jacobi import _round_robin

    for nb in (2, 4, 6, 10):
        rounds = _round_robin(nb)
        assert len(rounds) == nb - 1
        seen = [p for r in rounds for p in r]
        assert len(seen) == len(set(seen)) == nb * (nb - 1) // 2
        for r in rounds:  # disjoint within a round
            flat = [x for p in r for x in p]
            assert sorted(flat) == list(range(nb))


@pytest.mark.parametrize(("shape", "repeat"), [((6,), (10,)), ((4, 3), (4, 6)), ((4, 4), (8, 8)), ((3, 2, 2), (2, 3, 4)),
                                                ((7, 7, 7), (2, 2, 2))])
def test_cell_fft_moments_fused_equals_cell_fft_then_measure(cuda, shape, repeat):
    """sqb_cell_fft_moments (K6 fused into the last cell-FFT pass, the position-basis states are never written)
    against sqb_cell_fft + sqb_measure_moments and against the oracle, for every axis, with and without per-state
    offsets and wrapping (the variance pass re-reduces the SAME tile: stage 1)."""
    torch, kernels = _gpu()
    nd = len(shape)
    rng = np.random.default_rng(77)
    big = tuple(n * r for n, r in zip(shape, repeat))
    m = int(np.prod(big))
    lead = 5
    cell_np = rng.normal(size=(lead, m)) + 1j * rng.normal(size=(lead, m))
    lengths = [2 * np.pi * r * (1.0 + 0.25 * a) for a, r in enumerate(repeat)]
    spacing = [ln / g for ln, g in zip(lengths, big)]
    pos = kernels.cell_fft(shape, repeat, kernels.to_device(cell_np, torch.complex128), 1)
    offsets_np = rng.uniform(-3, 3, size=(nd, lead))
    offsets = kernels.to_device(offsets_np, torch.float64)
    tile = kernels.to_device(cell_np, torch.complex128)
    fused0 = kernels.cell_fft_moments(shape, repeat, tile, spacing, lengths, stage=0).cpu().numpy()
    fused1 = kernels.cell_fft_moments(shape, repeat, tile, spacing, lengths, stage=1, offsets=offsets, wrapped=True).cpu().numpy()
    for axis in range(nd):
        want0 = kernels.measure_moments(pos, big, axis, spacing[axis], lengths[axis]).cpu().numpy()
        want1 = kernels.measure_moments(pos, big, axis, spacing[axis], lengths[axis], offsets[axis].contiguous(), True).cpu().numpy()
        scale = np.abs(want0).max()
        np.testing.assert_allclose(fused0[:, axis, :], want0, atol=1e-11 * scale)
        np.testing.assert_allclose(fused1[:, axis, :3], want1[:, :3], atol=1e-11 * np.abs(want1).max())
    # and the oracle: <x> along every axis of the position-basis states
    pos_np = pos.cpu().numpy()
    dx = np.diag(lengths)
    for axis in range(nd):
        np.testing.assert_allclose(fused0[:, axis, 1] / fused0[:, axis, 0], o.measure_all_x(pos_np, dx, big, axis), atol=1e-10)
