"""CPU: host-side logic - metadata, basis bookkeeping, workload configs, sharding (gloo, world_size 2),
and the numpy model of the GPU eigensolver algorithm."""

import os
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

from oracle import bloch_oracle as o

ROOT = Path(__file__).resolve().parent.parent


def test_metadata_and_bases_bookkeeping():
    from slate_quantum_b200.bloch.build import BlochFractionMetadata, basis_from_metadata, get_sample_fractions
    from slate_quantum_b200.core import basis, metadata
    from slate_quantum_b200.metadata import RepeatedMetadata, repeat_volume_metadata, unit_cell_metadata

    meta = metadata.volume.spaced_volume_metadata_from_stacked_delta_x(
        (np.array([2 * np.pi, 0]), np.array([0, 2 * np.pi])), (4, 3)
    )
    assert meta.shape == (4, 3)
    np.testing.assert_allclose(metadata.fundamental_stacked_dk(meta), np.eye(2), atol=1e-15)
    np.testing.assert_allclose(metadata.fundamental_stacked_delta_k(meta), np.diag([4.0, 3.0]), atol=1e-15)
    np.testing.assert_array_equal(metadata.fundamental_nk_points(5), [0, 1, 2, -2, -1])
    rep = repeat_volume_metadata(meta, (3, 2))
    assert isinstance(rep.children[0], RepeatedMetadata)
    assert rep.shape == (12, 6)
    assert rep.children[0].domain.delta == pytest.approx(3 * 2 * np.pi)
    assert unit_cell_metadata(rep) == meta
    assert rep == repeat_volume_metadata(meta, (3, 2))
    assert rep != repeat_volume_metadata(meta, (2, 3))
    fr = get_sample_fractions(BlochFractionMetadata.from_repeats((4, 2)))
    np.testing.assert_array_equal(np.array(fr), o.sample_fractions((4, 2)))
    bloch_basis = basis_from_metadata(rep)
    assert bloch_basis.size == 72
    assert bloch_basis.metadata() == rep
    assert bloch_basis == basis_from_metadata(rep)
    assert bloch_basis.dual_basis() != bloch_basis
    assert bloch_basis.inner.supercell_momentum_basis() == basis.transformed_from_metadata(rep)
    tb = basis.transformed_from_metadata(meta)
    assert tb.is_dual == (False, False) and tb.dual_basis().is_dual == (True, True)


def test_potential_descriptions_match_oracle():
    from slate_quantum_b200.core import metadata
    from slate_quantum_b200.operator import build

    meta = metadata.volume.spaced_volume_metadata_from_stacked_delta_x(
        (np.array([2 * np.pi, 0]), np.array([0, 2 * np.pi])), (8, 6)
    )
    np.testing.assert_allclose(build.cos_potential(meta, 3.0, offset=0.2).raw_data,
                               o.cos_potential_components((8, 6), 3.0, 0.2).ravel())
    np.testing.assert_allclose(build.sin_potential(meta, 3.0).raw_data, o.sin_potential_components((8, 6), 3.0).ravel())
    np.testing.assert_allclose(build.fcc_potential(meta, 2.0).raw_data, o.fcc_potential_components((8, 6), 2.0).ravel())
    pot = build.cos_potential(meta, 1.0)
    assert pot.basis.size == 9
    assert pot.basis.metadata().children[0] == meta


def test_workload_configs():
    from slate_quantum_b200.configs import CONFIGS, shard_blocks

    c3 = CONFIGS["cfg3"]
    assert (c3.n, c3.n_k, c3.m, c3.n_times) == (256, 4096, 1 << 20, 4096)
    assert CONFIGS["cfg5"].n == 343 and CONFIGS["cfg4"].n == 512
    np.testing.assert_allclose(c3.potential_table(), o.pad_components(c3.shape, o.cos_potential_components(c3.shape, 10.0)))
    c4 = CONFIGS["cfg4"]
    np.testing.assert_allclose(c4.dk(), o.stacked_dk(c4.vectors()))
    psi = CONFIGS["cfg1"].initial_state()
    assert psi.shape == (6400,) and abs(np.linalg.norm(psi) - 1) < 1e-12
    for n_k, world in [(4096, 8), (10, 4), (7, 8)]:
        ranges = [shard_blocks(n_k, r, world) for r in range(world)]
        assert sum(c for _, c in ranges) == n_k
        assert all(ranges[i][0] + ranges[i][1] == ranges[i + 1][0] for i in range(world - 1))


@pytest.mark.parametrize("n", [1, 2, 7, 40])
def test_eigensolver_algorithm_model(n):
    from eigh_model import eigh_model

    rng = np.random.default_rng(n)
    g = rng.normal(size=(n, n)) + 1j * rng.normal(size=(n, n))
    h = g + g.conj().T
    w, t, fails, nrot, _ = eigh_model(h)
    assert fails == 0
    np.testing.assert_allclose(w, np.linalg.eigvalsh(h), atol=1e-12 * max(1.0, np.abs(h).max() * n))
    v = t.T
    assert np.linalg.norm(h @ v - v * w) / max(np.linalg.norm(h), 1e-300) < 1e-12
    assert nrot <= 5 * n * n // 2 + 64 * n + 64  # rotation-record capacity of csrc/eigh.cu


_WORKER = r"""
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.environ["SQB_ROOT"])
from oracle import bloch_oracle as o
from slate_quantum_b200.configs import shard_blocks
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:" + os.environ["SQB_PORT"],
                        rank=int(os.environ["RANK"]), world_size=2)
rank = dist.get_rank()
shape, repeat = (4, 3), (3, 3)
dx = 2 * np.pi * np.eye(2)
comps = o.cos_potential_components(shape, 2.0)
n_k = 9
begin, count = shard_blocks(n_k, rank, 2)
blocks = o.bloch_blocks(dx, shape, comps, o.HBAR_SI**2, repeat, block_begin=begin, block_count=count)
w, _ = o.eigh_blocks(blocks)
# ragged all-gather of eigenvalues, as bench.py / the solver do with NCCL
counts = [shard_blocks(n_k, r, 2)[1] for r in range(2)]
pad = max(counts)
mine = torch.zeros((pad, 12), dtype=torch.float64)
mine[:count] = torch.from_numpy(w)
gathered = [torch.zeros_like(mine) for _ in range(2)]
dist.all_gather(gathered, mine)
full = torch.cat([g[:c] for g, c in zip(gathered, counts)]).numpy()
w_all, _ = o.eigh_blocks(o.bloch_blocks(dx, shape, comps, o.HBAR_SI**2, repeat))
assert np.abs(full - w_all).max() < 1e-13, np.abs(full - w_all).max()
# k-shards -> time-shards exchange used before the across-block cell FFT (NCCL all_to_all on GPUs;
# grouped send/recv here).  Equal block counts per rank as in bench.py's weak-scaling layout.
from slate_quantum_b200.pipeline import exchange_k_to_time
T, width = 6, 5
glob = np.arange(T * 2 * width).reshape(T, 2 * width) * (1 + 0.5j)
mine_k = torch.from_numpy(glob[:, rank * width:(rank + 1) * width].copy())
got = exchange_k_to_time(mine_k).numpy()
want = glob[rank * (T // 2):(rank + 1) * (T // 2)]
assert got.shape == want.shape and np.array_equal(got, want), (got, want)
dist.barrier()
dist.destroy_process_group()
print("rank", rank, "ok")
"""


def test_block_sharding_and_eigenvalue_gather_world_size_2(tmp_path):
    """N>1 path on CPU: contiguous k-block ranges + all-gather of eigenvalues over gloo, world_size 2."""
    script = tmp_path / "worker.py"
    script.write_text(_WORKER)
    port = str(29500 + os.getpid() % 2000)
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), SQB_ROOT=str(ROOT), SQB_PORT=port, MASTER_ADDR="127.0.0.1")
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT, text=True))
    for p in procs:
        out, _ = p.communicate(timeout=180)
        assert p.returncode == 0, out


_RING_WORKER = r"""
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.environ["SQB_ROOT"])
from slate_quantum_b200.pipeline import ring_exchange, ring_segments
world = int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:" + os.environ["SQB_PORT"],
                        rank=int(os.environ["RANK"]), world_size=world)
rank = dist.get_rank()
blocks, n = 3, 4
def transform_of(r):  # what rank r's folded eigenvectors look like
    return torch.from_numpy((np.arange(blocks * n * n).reshape(blocks, n, n) + 1000 * r) * (1 - 2j))
local = transform_of(rank)
gathered = torch.full((world, blocks, n, n), np.nan + 0j, dtype=torch.complex128)
order = []
ring_exchange(local, gathered, on_arrival=order.append)
assert order == list(range(1, world)), order
segs = ring_segments(rank, world, blocks)
assert [s[0] for s in segs] == [(rank - k) % world for k in range(world)]
assert sorted(s[1] for s in segs) == [r * blocks for r in range(world)] and all(s[2] == blocks for s in segs)
for k, (src, b0, cnt) in enumerate(segs):
    if k == 0:
        assert src == rank and torch.isnan(gathered[src].real).all()  # own slot untouched
    else:  # the segment visited k-th is the one that arrived in ring step k
        assert torch.equal(gathered[src], transform_of(src)), (rank, src)
# real tensors travel as they are
gr = torch.zeros((world, 5), dtype=torch.float64)
ring_exchange(torch.full((5,), float(rank), dtype=torch.float64), gr)
assert all(torch.equal(gr[r], torch.full((5,), float(r), dtype=torch.float64)) for r in range(world) if r != rank)
try:
    ring_exchange(local, gathered[:, :1])
except ValueError:
    pass
else:
    raise AssertionError("shape mismatch not caught")
dist.barrier()
dist.destroy_process_group()
print("rank", rank, "ok")
"""


@pytest.mark.parametrize("world", [2, 3])
def test_ring_exchange_of_eigenvectors_gloo(tmp_path, world):
    """The eigenvector exchange of the time-sharded evolve (pipeline.TimeShardedEvolver) over gloo on CPU."""
    script = tmp_path / "ring_worker.py"
    script.write_text(_RING_WORKER)
    port = str(31600 + os.getpid() % 2000 + world)
    procs = []
    for rank in range(world):
        env = dict(os.environ, RANK=str(rank), WORLD_SIZE=str(world), SQB_ROOT=str(ROOT), SQB_PORT=port,
                   MASTER_ADDR="127.0.0.1")
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT, text=True))
    for p in procs:
        out, _ = p.communicate(timeout=180)
        assert p.returncode == 0, out


def test_ring_segments_validation():
    from slate_quantum_b200.pipeline import ring_segments

    assert ring_segments(0, 1, 7) == [(0, 0, 7)]
    assert ring_segments(1, 4, 2) == [(1, 2, 2), (0, 0, 2), (3, 6, 2), (2, 4, 2)]
    with pytest.raises(ValueError):
        ring_segments(4, 4, 2)


@pytest.mark.parametrize(("tiles_m", "tiles_n"), [(1, 1), (3, 5), (8, 8), (9, 2), (64, 64), (17, 33), (1, 40)])
def test_zgemm_grouped_tile_order_is_a_bijection(tiles_m, tiles_n):
    """csrc/zgemm.cu maps the linear CTA index to (tile row, tile column) in groups of kZgGroup = 8 tile rows walked
    column by column; every tile must be visited exactly once, and a group's CTAs must be contiguous."""
    group = 8
    seen = []
    for bid in range(tiles_m * tiles_n):
        width = group * tiles_n
        first = (bid // width) * group
        gsize = min(tiles_m - first, group)
        seen.append((first + (bid % width) % gsize, (bid % width) // gsize))
    assert sorted(seen) == [(i, j) for i in range(tiles_m) for j in range(tiles_n)]
    # the first min(8, tiles_m) CTAs share one B panel (same tile column)
    head = seen[: min(group, tiles_m)]
    assert len({j for _, j in head}) == 1


def test_x_k_p_builders_bookkeeping_on_cpu():
    """operator.build.{x, k, k_squared, p, momentum_from_function} and noise.build.caldeira_leggett_operators are
    host-side constructors: their stored diagonals and bases can be checked without a GPU (reference
    tests/test_operator.py:170-210, :451-462)."""
    from slate_quantum_b200 import noise, operator
    from slate_quantum_b200.configs import HBAR
    from slate_quantum_b200.core.basis import DiagonalBasis, RecastBasis, _strip
    from slate_quantum_b200.core.metadata import spaced_volume_metadata_from_stacked_delta_x, volume

    meta = spaced_volume_metadata_from_stacked_delta_x((np.array([2 * np.pi]),), (5,))
    x = operator.build.x(meta, axis=0)
    np.testing.assert_allclose(x.raw_data, np.linspace(0, 2 * np.pi, 5, endpoint=False))
    k = operator.build.k(meta, axis=0)
    np.testing.assert_array_equal(k.raw_data, [0, 1, 2, -2, -1])
    np.testing.assert_array_equal(operator.build.k_squared(meta, axis=0).raw_data, [0, 1, 4, 4, 1])
    np.testing.assert_allclose(operator.build.p(meta, axis=0).raw_data, HBAR * np.array([0, 1, 2, -2, -1]))
    for op in (x, k):
        stored = _strip(op.basis)
        assert isinstance(stored, RecastBasis) and isinstance(_strip(stored.inner), DiagonalBasis)
        assert op.basis.metadata().children[0] == meta
    shifted = operator.build.x(meta, axis=0, offset=-np.pi, wrapped=True)
    np.testing.assert_allclose(shifted.raw_data, o.stacked_x_points(np.array([[2 * np.pi]]), (5,), (-np.pi,), True)[0])
    f_k = operator.build.momentum_from_function(meta, lambda kk: kk[0] ** 3, wrapped=True, offset=(0.25,))
    np.testing.assert_allclose(f_k.raw_data, o.stacked_k_points(np.array([[2 * np.pi]]), (5,), (0.25,), True)[0] ** 3)
    np.testing.assert_allclose(volume.fundamental_stacked_dx(meta), [[2 * np.pi / 5]])
    cl = noise.build.caldeira_leggett_operators(meta)
    assert len(cl) == 1
    np.testing.assert_allclose(cl.raw_data.ravel(), x.raw_data)
    np.testing.assert_array_equal(cl.basis.metadata().children[0].values, [1])


def test_time_rows_of_a_rank_partition_the_grid_and_stay_uniform():
    """pipeline.shard_time_rows: contiguous or interleaved ownership of the time rows; every row has exactly one owner,
    and the interleaved rows of a uniform grid are a uniform grid with step G dt (what the NUFFT evolve is given)."""
    import numpy as np
    import pytest

    from slate_quantum_b200 import kernels
    from slate_quantum_b200.pipeline import shard_time_rows

    times = 0.25 + 0.125 * np.arange(48)
    for world in (1, 2, 3, 8):
        for interleaved in (False, True):
            owned = np.concatenate([np.arange(48)[shard_time_rows(r, world, 48, interleaved)] for r in range(world)])
            assert sorted(owned.tolist()) == list(range(48))
            for r in range(world):
                local = times[shard_time_rows(r, world, 48, interleaved)]
                assert len(local) == 48 // world
                step = kernels.uniform_step(local) if len(local) > 1 else None
                if len(local) > 1:
                    assert step == pytest.approx(0.125 * (world if interleaved else 1))
    with pytest.raises(ValueError, match="divisible"):
        shard_time_rows(0, 5, 48)


def test_time_reversal_eligibility_decisions_on_the_host():
    """BlochSolver._time_reversal_applies: the pairing of the blocks of k and -k is used only when it provably holds -
    orthogonal lattice, the whole k-grid on this solver, a Hermitian-symmetric potential table (a real V(x))."""
    import numpy as np
    import torch

    from oracle import bloch_oracle as o
    from slate_quantum_b200.pipeline import BlochSolver

    hbar = o.HBAR_SI
    shape, repeat = (6, 4), (4, 3)
    dk = o.stacked_dk(2 * np.pi * np.eye(2))

    def solver(**kw):
        s = BlochSolver(shape, repeat, kw.pop("dk", dk), hbar**2, hbar, **kw)
        s.time_reversal = True
        return s

    cos = torch.from_numpy(o.pad_components(shape, o.cos_potential_components(shape, 6.0)).astype(np.complex128).ravel())
    sin = torch.from_numpy(o.pad_components(shape, o.sin_potential_components(shape, 4.0, 0.2)).astype(np.complex128).ravel())
    s = solver()
    assert not s._time_reversal_applies()  # noqa: SLF001  (no table seen yet)
    for table in (cos, sin):  # sin: an asymmetric but still Hermitian-symmetric table, Vt[-q] = conj(Vt[q])
        s._check_table(table)  # noqa: SLF001
        assert s._time_reversal_applies()  # noqa: SLF001
    bad = cos.clone()
    bad[1] += 0.25j  # V(x) no longer real
    s._check_table(bad)  # noqa: SLF001
    assert not s._time_reversal_applies()  # noqa: SLF001
    s._check_table(cos)  # noqa: SLF001
    s.time_reversal = False
    assert not s._time_reversal_applies()  # noqa: SLF001
    # a rank's share of the k-grid does not contain its partners
    part = solver(block_begin=0, block_count=6)
    part._check_table(cos)  # noqa: SLF001
    assert not part._time_reversal_applies()  # noqa: SLF001
    # the hexagonal lattice (quirk Q1) breaks the pairing
    hexa = solver(dk=o.stacked_dk(2 * np.pi * np.array([[1.0, 0.0], [0.5, np.sqrt(3) / 2]])))
    hexa._check_table(cos)  # noqa: SLF001
    assert not hexa._time_reversal_applies()  # noqa: SLF001
