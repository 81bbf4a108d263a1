"""Where the time of the API observables leg goes (GPU dev probe)."""
import sys, time
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import bench  # noqa: E402
from slate_quantum_b200 import bloch, dynamics, kernels  # noqa: E402
from slate_quantum_b200.operator import measure  # noqa: E402
from slate_quantum_b200.state import State  # noqa: E402

cfg = bench.CONFIGS["cfg3"]
cell_meta, _, position_basis, times, build_potential = bench.api_objects(cfg)
psi_h = torch.from_numpy(cfg.initial_state()).pin_memory()


def tick(label, t0):
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    print(f"{label:40s} {1e3 * (t1 - t0):8.1f} ms")
    return t1


for rep in range(2):
    print("--- rep", rep)
    torch.cuda.synchronize()
    t = time.perf_counter()
    potential = build_potential(cell_meta, cfg.height)
    hamiltonian = bloch.build.kinetic_hamiltonian(potential, cfg.mass(), cfg.repeat)
    t = tick("build", t)
    evolution = dynamics.solve_schrodinger_equation_decomposition(State(position_basis, psi_h), times, hamiltonian)
    t = tick("solve (eigh + project)", t)
    lazy = evolution.with_state_basis(position_basis)
    lazy._transform()
    t = tick("fold", t)
    shape, repeat = lazy.bloch_dims()
    space = evolution.basis.metadata().children[1]
    geom = [measure._axis_geometry(space, a) for a in range(cfg.nd)]
    spacing, length = [g[1] for g in geom], [g[2] for g in geom]
    for t0_, tile in lazy.cell_tiles():
        t = tick(f"evolve tile {tuple(tile.shape)}", t)
        mom = kernels.cell_fft_moments(shape, repeat, tile, spacing, length, stage=0)
        t = tick("fused stage 0 (axis-0 pass + fused last)", t)
        off = torch.zeros((cfg.nd, tile.shape[0]), dtype=torch.float64, device="cuda")
        mom2 = kernels.cell_fft_moments(shape, repeat, tile, spacing, length, stage=1, offsets=off, wrapped=True)
        t = tick("fused stage 1", t)
    del lazy, evolution, hamiltonian
