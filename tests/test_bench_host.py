"""CPU tests of bench.py's host-side pieces and of the oracle helpers the parity gate uses."""

import json
import subprocess
import sys
from pathlib import Path

import numpy as np

from oracle import bloch_oracle as o

ROOT = Path(__file__).resolve().parents[1]


def _bench():
    sys.path.insert(0, str(ROOT))
    import bench

    return bench


def test_config_dict_is_identical_for_both_arms_and_scalings():
    bench = _bench()
    cfg = bench.CONFIGS["cfg3"]
    weak = bench.config_dict(cfg, 8, False, True)
    assert weak["repeat_global"] == [512, 64] and weak["k_blocks_per_gpu"] == 4096
    strong = bench.config_dict(cfg, 8, True, True)
    assert strong["repeat_global"] == [64, 64] and strong["k_blocks_per_gpu"] == 512
    assert weak["workload"] == cfg.name and weak["parallelism"] == "k-block sharding x8"
    assert "L2" in weak["l2_policy"]


def test_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the oracle port on the host cores): one JSON line, the same `config` dict as the
    GPU arm, its own BLAS thread count even when the launcher exports OMP_NUM_THREADS=1, an auditable sample."""
    bench = _bench()
    import os

    env = dict(os.environ, OMP_NUM_THREADS="1", WORLD_SIZE="2", RANK="0")
    out = subprocess.run(
        [sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--config", "cfg1", "--steps", "1", "--warmup", "1",
         "--gpus", "2"],
        capture_output=True, text=True, timeout=600, env=env, check=True,
    ).stdout.strip().splitlines()
    assert len(out) == 1
    line = json.loads(out[0])
    assert line["impl"] == "reference" and line["metric"] == bench.METRIC and line["unit"] == bench.UNIT
    # default scaling mode = strong: the config's own k-grid split over the GPUs, same dict as the GPU arm
    assert line["scaling"] == "strong"
    assert line["config"] == bench.config_dict(bench.CONFIGS["cfg1"], 2, True, True)
    assert line["e2e"] == {"value": line["value"], "unit": bench.UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["sample_wall_s"] > 0
    assert line["cpu_baseline"]["blas_threads"] >= 1 and line["n_gpus"] == 2
    # another rank of the same launch prints nothing
    env["RANK"] = "1"
    quiet = subprocess.run(
        [sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--config", "cfg1", "--steps", "1", "--warmup", "1",
         "--gpus", "2"],
        capture_output=True, text=True, timeout=600, env=env, check=True,
    ).stdout.strip()
    assert quiet == ""


def test_ncu_traffic_is_read_from_the_committed_capture():
    bench = _bench()
    index = json.loads((ROOT / "profiles" / "traffic_index.json").read_text())
    entry = index["evolve_bloch_3m_kernel"]
    traffic, src = bench.ncu_traffic("evolve_bloch_3m_kernel", "cfg3-2d-square-16x16-64x64k-T4096", entry["t_tile"])
    assert src == f"profiles/{entry['csv']}" and 3e10 < traffic < 6e10
    assert bench.ncu_traffic("evolve_bloch_3m_kernel", "cfg3-2d-square-16x16-64x64k-T4096", 1000) == (None, None)
    assert bench.ncu_traffic("evolve_bloch_3m_kernel", "cfg2-1d", entry["t_tile"]) == (None, None)
    assert bench.ncu_traffic("no_such_kernel", "cfg3", 2048) == (None, None)


def test_sample_blocks_span_the_grid():
    bench = _bench()
    assert bench.sample_blocks(4096) == [0, 1365, 2049, 4095]
    assert bench.sample_blocks(1) == [0]
    for n_k in (2, 3, 100, 32768):
        got = bench.sample_blocks(n_k)
        assert got == sorted(set(got)) and got[0] == 0 and got[-1] == n_k - 1


def test_sampled_block_parity_helpers_against_the_full_pipeline():
    """evolve_sampled_blocks / position_rows_to_blocks / sampled_block_parity (the checker of bench.py's parity gate
    and of the full-size GPU tests) against the oracle's own full pipeline, incl. a hexagonal cell and a wrong row."""
    hb = o.HBAR_SI
    rng = np.random.default_rng(3)
    for dx, shape, repeat, comps in (
        (2 * np.pi * np.eye(2), (4, 3), (3, 4), o.cos_potential_components((4, 3), 10.0)),
        (2 * np.pi * np.array([[np.sin(np.pi / 3), np.cos(np.pi / 3)], [0.0, 1.0]]), (4, 4), (2, 3),
         o.fcc_potential_components((4, 4), 5.0)),
        (2 * np.pi * np.eye(1), (6,), (5,), o.sin_potential_components((6,), 2.0, 0.1)),
    ):
        w, t = o.eigh_blocks(o.bloch_blocks(dx, shape, comps, hb**2, repeat))
        m = int(np.prod(shape) * np.prod(repeat))
        psi = rng.normal(size=m) + 1j * rng.normal(size=m)
        psi /= np.linalg.norm(psi)
        times = np.linspace(0, 4 * hb, 9)
        ref = o.evolve_pipeline(psi, shape, repeat, w, t, times, hb)
        n_k = int(np.prod(repeat))
        blocks = sorted({0, n_k // 2, n_k - 1})
        rows = [0, 4, 8]
        assert o.sampled_block_parity(dx, shape, comps, hb**2, repeat, psi, times[rows], ref["position"][rows], blocks, hb) < 1e-12
        assert o.sampled_block_parity(dx, shape, comps, hb**2, repeat, psi, times[rows], ref["bloch"][rows], blocks, hb,
                                      basis="bloch") < 1e-12
        wrong = ref["position"][rows].copy()
        wrong[2] = ref["position"][7]  # a state of the wrong time must be caught
        assert o.sampled_block_parity(dx, shape, comps, hb**2, repeat, psi, times[rows], wrong, blocks, hb) > 1e-3


def test_is_smooth():
    from slate_quantum_b200 import kernels

    assert all(kernels.is_smooth(n) for n in (1, 2, 64, 100, 343, 1024, 31 * 29, 4096))
    assert not any(kernels.is_smooth(n) for n in (0, 37, 2 * 37, 1009))


def test_tile_sizing_keeps_whole_evolve_tiles_where_two_position_tiles_do_not_fit():
    """bench.choose_tiles: one GPU takes the whole time grid when two tiles fit, else equal halves; the time-sharded
    multi-GPU layout first shrinks the POSITION tile (down to a quarter of the evolve tile) so that a rank's share stays
    one whole 512-row NUFFT tile; without the position stage a single buffer may use 80 % of the memory."""
    bench = _bench()
    gb = 1 << 30
    row = 4096 * 256 * 16  # cfg3 on one GPU: 16.8 MB per time row
    assert bench.choose_tiles(4096, row, 165 * gb, True, False) == (4096, 4096)
    assert bench.choose_tiles(4096, row, 150 * gb, True, False) == (2048, 2048)
    assert bench.choose_tiles(4096, row, 60 * gb, True, False) == (1024, 1024)
    # 8 GPUs, weak scaling: 512 rows of a 512 x 64 block grid = 68.7 GB per tile
    assert bench.choose_tiles(512, 8 * row, 135 * gb, True, True) == (512, 256)
    assert bench.choose_tiles(512, 8 * row, 100 * gb, True, True) == (512, 128)
    assert bench.choose_tiles(512, 8 * row, 70 * gb, True, True) == (256, 128)
    # strong scaling: everything fits
    assert bench.choose_tiles(512, row, 150 * gb, True, True) == (512, 512)
    assert bench.choose_tiles(4096, row, 60 * gb, False, False) == (2048, 2048)
    t, o_ = bench.choose_tiles(4096, row, 1 * gb, True, False)
    assert t == o_ and t <= 64
