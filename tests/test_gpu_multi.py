"""GPU, world_size >= 2: the time-sharded position-basis evolve (pipeline.TimeShardedEvolver) over NCCL against
the oracle chain.  Skipped on a single-GPU box; run it with `gpurun --gpus 2 -- python -m pytest tests -m gpu`."""

from __future__ import annotations

import os
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
pytestmark = pytest.mark.gpu

_WORKER = r"""
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.environ["SQB_ROOT"])
from oracle import bloch_oracle as o
from slate_quantum_b200 import kernels
from slate_quantum_b200.pipeline import BlochSolver, TimeShardedEvolver
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", init_method="tcp://127.0.0.1:" + os.environ["SQB_PORT"], rank=rank,
                        world_size=world, device_id=torch.device("cuda", rank))
HBAR = o.HBAR_SI
dx = 2 * np.pi * np.eye(2)
shape, repeat = (8, 8), (4, 2 * world)
n_k = int(np.prod(repeat)); cnt = n_k // world
comps = o.sin_potential_components(shape, 4.0, 0.2)
solver = BlochSolver(shape, repeat, o.stacked_dk(dx), HBAR**2, HBAR, rank * cnt, cnt)
solver.build(kernels.to_device(o.pad_components(shape, comps).ravel(), torch.complex128))
solver.diagonalise()
assert int(solver.info.abs().sum().item()) == 0
rng = np.random.default_rng(10)
psi = rng.normal(size=solver.m) + 1j * rng.normal(size=solver.m)
psi /= np.linalg.norm(psi)
T = 24 * world
blocks = o.bloch_blocks(dx, shape, comps, HBAR**2, repeat)
w_ref, t_ref = o.eigh_blocks(blocks)
c = solver.project(kernels.to_device(psi, torch.complex128))
for mode in ("p2p", "nccl"):
    sharded = TimeShardedEvolver(solver, T, exchange=mode)
    assert sharded.exchange_mode == mode
    assert sharded.rows == 24 and sharded.time_begin == 24 * rank
    eig = sharded.gather_eigenvalues().cpu().numpy()
    assert np.abs(eig - w_ref).max() <= 1e-10 * np.abs(w_ref).max()
    try:
        sharded.start_exchange(c)
    except RuntimeError:
        pass
    else:
        raise AssertionError("start_exchange before fold() must fail")
    for uniform in (True, False):
        times = np.linspace(0, 5 * HBAR, T) if uniform else np.sort(rng.uniform(0, 5 * HBAR, T))
        dt = kernels.uniform_step(times) if uniform else None
        assert (dt is not None) == uniform
        want = o.evolve_pipeline(psi, shape, repeat, w_ref, t_ref, times, HBAR)["position"]
        lo = sharded.time_begin
        t_dev = kernels.to_device(times[lo:lo + sharded.rows], torch.float64)
        for t_tile in (sharded.rows, 10):  # one tile / ragged tiles reusing the Re+Im plane
            sharded.fold()  # a new generation each time: alternates the two symmetric buffers
            sharded.start_exchange(c)
            states = torch.empty((t_tile, solver.m), dtype=torch.complex128, device="cuda")
            out = torch.empty_like(states)
            got = np.empty((sharded.rows, solver.m), dtype=np.complex128)
            def keep(t0, tt, tile):
                got[t0:t0 + tt] = tile.cpu().numpy()
            sharded.evolve_position(t_dev, states, out, dt, on_tile=keep)
            torch.cuda.synchronize()
            err = np.abs(got - want[lo:lo + sharded.rows]).max()
            assert err < 1e-9, (rank, mode, uniform, t_tile, err)
    # interleaved time ownership (rank r evolves rows r, r + G, ...: what bench.py uses on uniform grids) with enough rows
    # per rank for the NUFFT route of the uniform-grid evolve (n = 64 states, 600 rows: one full tile + a ragged one)
    T2 = 600 * world
    times = np.linspace(0, 40 * HBAR, T2)
    dt = kernels.uniform_step(times)
    assert kernels.evolve_uniform_route(solver.n, cnt, 600) == "nufft"
    sharded2 = TimeShardedEvolver(solver, T2, exchange=mode)
    sharded2.gather_eigenvalues()
    rows_sl = sharded2.time_rows(interleaved=True)
    assert np.arange(T2)[rows_sl].tolist() == list(range(rank, T2, world))
    sel = [0, 1, 255, 511, 512, 599]  # local rows checked against the oracle (the oracle evolves only those times)
    want = o.evolve_pipeline(psi, shape, repeat, w_ref, t_ref, times[rows_sl][sel], HBAR)["position"]
    for t_tile, o_tile in ((600, 600), (600, 150), (256, 256)):
        sharded2.fold()
        sharded2.start_exchange(c)
        states = torch.empty((t_tile, solver.m), dtype=torch.complex128, device="cuda")
        out = torch.empty((o_tile, solver.m), dtype=torch.complex128, device="cuda")
        got = np.empty((600, solver.m), dtype=np.complex128)
        def keep2(t0, tt, tile):
            got[t0:t0 + tt] = tile.cpu().numpy()
        sharded2.evolve_position(kernels.to_device(times[rows_sl], torch.float64), states, out, dt * world, on_tile=keep2)
        torch.cuda.synchronize()
        err = np.abs(got[sel] - want).max()
        assert err < 1e-9, (rank, mode, "interleaved", t_tile, o_tile, err)
    del sharded2
    dist.barrier()
dist.barrier()
dist.destroy_process_group()
print("rank", rank, "ok")
"""


def test_time_sharded_evolver_nccl(tmp_path):
    import torch

    world = min(torch.cuda.device_count(), 4)
    if world < 2:
        pytest.skip("needs at least 2 GPUs")
    script = tmp_path / "worker.py"
    script.write_text(_WORKER)
    port = str(33100 + os.getpid() % 2000)
    procs = []
    for rank in range(world):
        env = dict(os.environ, RANK=str(rank), WORLD_SIZE=str(world), SQB_ROOT=str(ROOT), SQB_PORT=port,
                   MASTER_ADDR="127.0.0.1")
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT, text=True))
    for p in procs:
        out, _ = p.communicate(timeout=600)
        assert p.returncode == 0, out
