#!/usr/bin/env python
"""bench.py -- one "step" = one pass of the hot path over one synthetic Bloch workload:

    K1 build every k-block  ->  K2 eigh of every block  ->  K4+K5+K3a project psi0  ->
    K3c evolve over T times (NUFFT on uniform grids, else the fused phase x eigenvector DMMA GEMM)  ->  K4b cell FFT to
    the position basis

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config cfg3] [--impl ours|reference]

Prints ONE JSON line (rank 0).  Multi-GPU: one process per GPU (torchrun), k-blocks sharded in contiguous ranges.
`value` is the STRONG-scaling figure by default - the config's own k-grid split over the ranks, the north_star's "cfg3
... scales >= 7x at 8 GPUs" - and the same line carries `weak_scaling` (each rank owns the config's k-grid, enlarged
along axis 0; `--scaling weak` swaps the two), `parity_check` (sampled blocks of the LAST time rows
against the CPU oracle, every N), `e2e` (through the reference-shaped API at N=1; position-basis, time-sharded states
to host at N>1), and short runs of the other BASELINE.json configs (`other_configs`).  See DESIGN.md "Measurement".
"""

from __future__ import annotations

import argparse
import csv
import ctypes
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

from slate_quantum_b200.configs import CONFIGS, HBAR, BlochWorkload, shard_blocks  # noqa: E402

METRIC = "bloch_k_blocks_per_s"
UNIT = "k-blocks/s (each block built, diagonalised and evolved over the config's time batch)"
PARITY_TOL = 1e-9  # north_star: evolved states max-abs 1e-9


# ----------------------------------------------------------------------------------------------------
# clocks sampling (B200_PROFILING.md recipe)
class ClockSampler:
    QUERY = (
        "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
        "clocks_event_reasons.sw_power_cap"
    )

    def __init__(self, index: int) -> None:
        self.index = index
        self.rows: list[list[str]] = []
        self._stop = threading.Event()
        self._thread: threading.Thread | None = None

    def _run(self) -> None:
        while not self._stop.is_set():
            try:
                out = subprocess.run(
                    ["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                    capture_output=True, text=True, timeout=5, check=False,
                ).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.splitlines()[0].split(",")])
            except Exception:  # noqa: BLE001
                pass
            self._stop.wait(0.2)

    def __enter__(self) -> "ClockSampler":
        self._thread = threading.Thread(target=self._run, daemon=True)
        self._thread.start()
        return self

    def __exit__(self, *exc) -> None:
        self._stop.set()
        if self._thread:
            self._thread.join(timeout=6)

    def summary(self) -> dict:
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm = sorted(float(r[1]) for r in self.rows if r[1].replace(".", "").isdigit())
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            for name, val in zip(names, r[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {
            "sm_mhz": sm[len(sm) // 2] if sm else None,
            "sm_max_mhz": float(self.rows[0][2]) if self.rows[0][2].replace(".", "").isdigit() else None,
            "power_w_max": max((float(r[3]) for r in self.rows if r[3].replace(".", "").isdigit()), default=None),
            "samples": len(self.rows),
            "reasons": sorted(reasons),
        }


def choose_tiles(rows: int, row_bytes: int, free_bytes: int, with_position: bool, gathered: bool) -> tuple[int, int]:
    """(evolve tile, position tile) in time rows for `rows` rows of `row_bytes` each and `free_bytes` of HBM.

    The two tiles need not be equal: the NUFFT evolve works in tiles of 512 rows whatever the call asks for, so the
    evolve tile (`states`) keeps as many rows as fit and - in the time-sharded multi-GPU layout, where two full tiles of
    the enlarged block grid do not fit beside the exchanged eigenvectors - the cell FFT runs in chunks of a smaller
    position tile (`pos_out`, down to a quarter of the evolve tile) before the evolve tile itself is halved."""
    t_tile = o_tile = int(rows)
    if not with_position:
        while t_tile * row_bytes > int(free_bytes * 0.8) and t_tile > 64:
            t_tile //= 2
        return t_tile, t_tile
    while (t_tile + o_tile) * row_bytes > int(free_bytes * 0.85) and o_tile > 64:
        if gathered and o_tile * 4 > t_tile:
            o_tile //= 2
        else:
            t_tile //= 2
            o_tile = min(o_tile, t_tile)
    return t_tile, o_tile


def config_dict(cfg: BlochWorkload, world: int, strong: bool, with_position: bool, **extra) -> dict:
    """The `config` object of the JSON line; the reference arm prints the SAME dict (same keys and values)."""
    repeat_global = tuple(cfg.repeat) if strong else (cfg.repeat[0] * world, *cfg.repeat[1:])
    n_k_local = cfg.n_k // world if strong else cfg.n_k
    out = {
        "workload": cfg.name,
        "shape": list(cfg.shape),
        "repeat_per_gpu": list(cfg.repeat),
        "repeat_global": list(repeat_global),
        "block_size_N": cfg.n,
        "k_blocks_per_gpu": n_k_local,
        "n_times": cfg.n_times,
        "position_basis": bool(with_position),
        "l2_policy": "inputs larger than L2: every step rebuilds and re-reads the "
        f"{n_k_local * cfg.n * cfg.n * 16 / 1e9:.2f} GB block array (L2 = 126 MB)",
        "parallelism": f"k-block sharding x{world}",
    }
    assert not extra
    return out


# ----------------------------------------------------------------------------------------------------
# CPU arm: the oracle (numpy restatement of the reference; the reference itself cannot be imported here)
def _blas_threads(want: int | None = None) -> int:
    """Threads the BLAS / LAPACK pools will use; `want` raises the limit (torchrun exports OMP_NUM_THREADS=1)."""
    try:
        from threadpoolctl import threadpool_info, threadpool_limits

        if want is not None:
            threadpool_limits(limits=want)
        return max((p.get("num_threads", 1) for p in threadpool_info()), default=1)
    except Exception:  # noqa: BLE001
        return 1


def cpu_sample(cfg: BlochWorkload, n_blocks: int, n_times: int, with_position: bool) -> dict:
    """Time the oracle on a bounded sample of `cfg`; extrapolate linearly to one full step."""
    from oracle import bloch_oracle as o

    cores = len(os.sched_getaffinity(0))
    blas_threads = _blas_threads(cores)  # all host threads, whatever OMP_NUM_THREADS the launcher exported
    comps = cfg.potential_components()
    n_blocks = min(n_blocks, cfg.n_k)
    n_times = min(n_times, cfg.n_times)
    times = cfg.times()[:n_times]

    def best(fn, reps=2):
        ts = []
        out = None
        for _ in range(reps):
            t0 = time.perf_counter()
            out = fn()
            ts.append(time.perf_counter() - t0)
        return min(ts), out

    t_build, blocks = best(lambda: o.bloch_blocks(cfg.vectors(), cfg.shape, comps, cfg.mass(), cfg.repeat, hbar=HBAR,
                                                    block_begin=0, block_count=n_blocks))
    t_eigh, (w, v) = best(lambda: o.eigh_blocks(blocks))
    rng = np.random.default_rng(0)
    psi_b = rng.normal(size=(n_blocks, cfg.n)) + 1j * rng.normal(size=(n_blocks, cfg.n))

    def evolve():
        c = o.project(v, psi_b)
        a = o.evolve_eigenbasis(c, w, times, HBAR)  # [T, B*N] - what the reference returns
        per_block = np.matmul(a.reshape(n_times, n_blocks, cfg.n).transpose(1, 0, 2), v)  # zgemm per block
        return np.ascontiguousarray(per_block.transpose(1, 0, 2)).reshape(n_times, -1)

    t_evolve, _ = best(evolve)
    t_pos = 0.0
    n_pos = min(2, n_times)
    if with_position:
        full = rng.normal(size=(n_pos, cfg.m)) + 1j * rng.normal(size=(n_pos, cfg.m))
        t_pos, _ = best(lambda: o.momentum_to_position(o.bloch_into_supercell(full, cfg.shape, cfg.repeat), cfg.supercell))
    per_block = (t_build + t_eigh) / n_blocks
    per_block_time = t_evolve / (n_blocks * n_times)
    step_s = cfg.n_k * per_block + cfg.n_k * cfg.n_times * per_block_time + (cfg.n_times * t_pos / n_pos if with_position else 0.0)
    return {
        "value": cfg.n_k / step_s,
        "unit": UNIT,
        "cores": cores,
        "blas_threads": blas_threads,
        "kind": "port",
        "sample": (
            f"{n_blocks} of {cfg.n_k} k-blocks (build+np.linalg.eigh), evolution of those blocks over {n_times} of "
            f"{cfg.n_times} times (np.exp + np.matmul)"
            + (f", {n_pos} supercell permute+ifftn" if with_position else "")
            + "; extrapolated linearly to the full step (uniform, embarrassingly parallel work)"
        ),
        "step_s_extrapolated": step_s,
        "blocks_diagonalised_per_s": n_blocks / (t_build + t_eigh),
        "state_time_evolutions_per_s": 1.0 / (cfg.n_k * per_block_time + (t_pos / n_pos if with_position else 0.0)),
        "sample_wall_s": t_build + t_eigh + t_evolve + t_pos,
    }


def run_reference(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = int(os.environ.get("WORLD_SIZE", str(args.gpus)))
    cfg = CONFIGS[args.config]
    strong = args.scaling == "strong"
    # bounded sample per step so that K+W steps end within a few minutes
    nb = 64 if cfg.n <= 256 else 16
    nt = 256
    for _ in range(args.warmup):
        cpu_sample(cfg, 2, 8, args.position)
    t0 = time.perf_counter()
    vals = [cpu_sample(cfg, nb, nt, args.position) for _ in range(args.steps)]
    wall = time.perf_counter() - t0
    best = max(vals, key=lambda d: d["value"])
    # the CPU arm has one set of host cores whatever N is: its per-block rate does not depend on the k-grid, so the
    # weak-scaling workload of N GPUs (N x the blocks) takes N x the time at the same k-blocks/s
    line = {
        "impl": "reference",
        "metric": METRIC,
        "value": best["value"],
        "unit": UNIT,
        "n_gpus": args.gpus,
        "steps": args.steps,
        "warmup": args.warmup,
        "ms_per_step": best["step_s_extrapolated"] * 1e3 * (1 if strong else world),
        "ms_per_step_is": "linear extrapolation of the timed sample to the full step (not a timed step); "
        "`sample_wall_s` is what was timed",
        "higher_is_better": True,
        "scaling": args.scaling,
        "vs_baseline": None,
        "dtype": "c128 (f64)",
        "data": "synthetic",
        "config": config_dict(cfg, world, strong, bool(args.position)),
        "cpu_baseline": {k: best[k] for k in ("value", "unit", "cores", "blas_threads", "kind", "sample", "sample_wall_s")},
        "e2e": {"value": best["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "sample_wall_s": best["sample_wall_s"],
        "note": "numpy/LAPACK restatement of the reference (reference needs Python>=3.14 + un-vendored slate_core); "
        f"all {vals[0]['blas_threads']} BLAS threads of {vals[0]['cores']} cores; total wall {wall:.1f}s",
    }
    print(json.dumps(line))


# ----------------------------------------------------------------------------------------------------
def _hbm_peak() -> float:
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"])
        except Exception:  # noqa: BLE001
            pass
    return 6650.0  # B200_PROFILING.md fallback


_UNIT_SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def ncu_traffic(kernel: str, cfg_name: str, t_tile: int) -> tuple[float | None, str | None]:
    """dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of `kernel`, read from the committed ncu
    `--set full` raw page named in profiles/traffic_index.json for exactly this (config, time tile); else None."""
    index = ROOT / "profiles" / "traffic_index.json"
    try:
        entry = json.loads(index.read_text()).get(kernel)
        if not entry or not cfg_name.startswith(entry["config"]) or int(entry["t_tile"]) != int(t_tile):
            return None, None
        rows = list(csv.reader((ROOT / "profiles" / entry["csv"]).open()))
        head, units = rows[0], rows[1]
        k_col = head.index("Kernel Name")
        total = 0.0
        for row in rows[2:]:
            if kernel in row[k_col]:
                for name in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                    c = head.index(name)
                    total += float(row[c].replace(",", "")) * _UNIT_SCALE[units[c]]
                return total, f"profiles/{entry['csv']}"
    except Exception:  # noqa: BLE001
        return None, None
    return None, None


def ncu_metric(kernel: str, cfg_name: str, t_tile: int, metric: str) -> float | None:
    """One metric of `kernel`'s row in the committed ncu `--set full` raw page named in profiles/traffic_index.json for
    exactly this (config, time tile); None when there is no matching capture (never a guess)."""
    try:
        entry = json.loads((ROOT / "profiles" / "traffic_index.json").read_text()).get(kernel)
        if not entry or not cfg_name.startswith(entry["config"]) or int(entry["t_tile"]) != int(t_tile):
            return None
        rows = list(csv.reader((ROOT / "profiles" / entry["csv"]).open()))
        head = rows[0]
        k_col, m_col = head.index("Kernel Name"), head.index(metric)
        for row in rows[2:]:
            if kernel in row[k_col]:
                return float(row[m_col].replace(",", ""))
    except Exception:  # noqa: BLE001
        return None
    return None


def bind_to_gpu_numa_node(local_rank: int) -> int | None:
    """Pin this process (and so its first-touch pinned allocations) to the NUMA node of its GPU."""
    try:
        import torch

        prop = torch.cuda.get_device_properties(local_rank)
        bdf = f"{prop.pci_domain_id:04x}:{prop.pci_bus_id:02x}:{prop.pci_device_id:02x}.0"
        node = int(Path(f"/sys/bus/pci/devices/{bdf}/numa_node").read_text().strip())
        if node < 0:
            return None
        cpus: set[int] = set()
        for part in Path(f"/sys/devices/system/node/node{node}/cpulist").read_text().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = cpus & os.sched_getaffinity(0)
        if allowed:
            os.sched_setaffinity(0, allowed)
            return node
    except Exception:  # noqa: BLE001
        return None
    return None


class Ctx:
    def __init__(self) -> None:
        import torch
        import torch.distributed as dist

        from slate_quantum_b200 import _lib

        self.torch, self.dist = torch, dist
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a CUDA device; slate_quantum_b200 has no CPU fallback")
        torch.cuda.set_device(self.local_rank)
        self.numa_node = bind_to_gpu_numa_node(self.local_rank) if self.world > 1 else None
        if self.world > 1:
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local_rank))
        self.lib = _lib.load()

    def sync_all(self) -> None:
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
            self.torch.cuda.synchronize()

    def max_over_ranks(self, values: list[float]) -> list[float]:
        if self.world == 1:
            return list(values)
        t = self.torch.tensor(values, dtype=self.torch.float64, device="cuda")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return t.tolist()

    def ev(self):
        return self.torch.cuda.Event(enable_timing=True)


def sample_blocks(n_k: int) -> list[int]:
    """A few k-blocks spread over the whole grid (so at N > 1 they lie in several ranks' ranges)."""
    return sorted({0, n_k // 3, min(n_k // 2 + 1, n_k - 1), n_k - 1})


def oracle_gate(cfg: BlochWorkload, repeat_global, psi0: np.ndarray, times_sel: np.ndarray, rows: np.ndarray,
                basis: str) -> tuple[float, list[int]]:
    """max |device - oracle| over sampled blocks of the given evolved rows (the CPU oracle is the checker)."""
    from oracle import bloch_oracle as o

    n_k = int(np.prod(repeat_global))
    blocks = sample_blocks(n_k)
    err = o.sampled_block_parity(
        cfg.vectors(), cfg.shape, cfg.potential_components(), cfg.mass(), repeat_global, psi0, times_sel, rows, blocks,
        HBAR, basis=basis,
    )
    return err, blocks


# ----------------------------------------------------------------------------------------------------
class PositionRun:
    """The measured step with position-basis output (pipeline.BlochSolver at N=1, + TimeShardedEvolver at N>1)."""

    STAGES = ["build", "eigh", "fold", "project", "evolve", "exchange", "position"]

    def __init__(self, ctx: Ctx, cfg: BlochWorkload, strong: bool, with_position: bool = True) -> None:
        from slate_quantum_b200 import kernels
        from slate_quantum_b200.pipeline import BlochSolver, TimeShardedEvolver

        torch = ctx.torch
        self.ctx, self.cfg, self.strong, self.with_position = ctx, cfg, strong, with_position
        world, rank = ctx.world, ctx.rank
        if strong and cfg.n_k % world:
            raise SystemExit(f"strong scaling needs n_k={cfg.n_k} divisible by the number of GPUs")
        self.repeat_global = tuple(cfg.repeat) if strong else (cfg.repeat[0] * world, *cfg.repeat[1:])
        self.n, self.T = cfg.n, cfg.n_times
        self.n_k_local = cfg.n_k // world if strong else cfg.n_k
        self.n_k_global = self.n_k_local * world
        n, T = self.n, self.T
        self.solver = BlochSolver(cfg.shape, self.repeat_global, cfg.dk(), cfg.mass(), HBAR, rank * self.n_k_local,
                                  self.n_k_local)
        # ---- host inputs (pinned) -> device
        self.table_h = torch.from_numpy(cfg.potential_table().ravel().copy()).pin_memory()
        wl_global = BlochWorkload(cfg.name, cfg.delta_x, cfg.shape, self.repeat_global, cfg.potential, cfg.height, T,
                                  cfg.t_max_over_hbar)
        self.psi_np = wl_global.initial_state()
        self.psi_h = torch.from_numpy(self.psi_np).pin_memory()
        self.times_np = cfg.times()
        self.times_h = torch.from_numpy(self.times_np).pin_memory()
        self.table = self.table_h.cuda(non_blocking=True)
        self.psi0 = self.psi_h.cuda(non_blocking=True)
        self.times = self.times_h.cuda(non_blocking=True)
        self.dt_uniform = kernels.uniform_step(self.times_np)  # uniform grid -> table-driven phases (3M kernel)
        # ---- buffers
        self.h_buf = torch.empty((self.n_k_local, n, n), dtype=torch.complex128, device="cuda")
        ws_bytes = ctx.lib.sqb_eigh_workspace_bytes(n, self.n_k_local)
        self.solver._eigh_ws = torch.empty(ws_bytes, dtype=torch.uint8, device="cuda")  # noqa: SLF001
        self.gathered = world > 1 and with_position  # time-sharded evolve over exchanged eigenvectors
        self.sharded = TimeShardedEvolver(self.solver, T) if self.gathered else None
        # time sharding at N > 1: on a uniform grid rank r evolves rows r, r + N, r + 2N ... (again a uniform grid, step
        # N dt: the wider step spreads the NUFFT evolve's angles when a rank's share is a single tile), else a
        # contiguous range
        self.interleaved = self.gathered and self.dt_uniform is not None
        if self.gathered:
            rows_sl = self.sharded.time_rows(self.interleaved)
            self.time_index = np.arange(T)[rows_sl]                      # global time index of every local row
            self.times_local = self.times[rows_sl].contiguous()
            self.dt_local = self.dt_uniform * world if self.interleaved else self.dt_uniform
        self.folded_buf = torch.empty_like(self.h_buf) if with_position and not self.gathered else None
        free = kernels.free_bytes()
        row_bytes = self.n_k_local * n * 16
        t_tile = T
        if self.gathered:
            row_bytes = self.n_k_global * n * 16  # a rank evolves T / world time samples of EVERY block
            t_tile = T // world
        t_tile, o_tile = choose_tiles(t_tile, row_bytes, free, with_position, self.gathered)
        self.t_tile, self.o_tile, self.row_bytes = t_tile, o_tile, row_bytes
        self.states = torch.empty((t_tile, row_bytes // 16), dtype=torch.complex128, device="cuda")
        self.pos_out = torch.empty((o_tile, row_bytes // 16), dtype=torch.complex128, device="cuda") if with_position else None
        self.eig_all = (
            torch.empty((world, self.n_k_local * n), dtype=torch.float64, device="cuda")
            if world > 1 and not self.gathered else None
        )
        self.last_tile: tuple[int, int] = (0, 0)  # (first global time index, rows) of what pos_out / states holds

    def step(self, record: list | None = None) -> None:
        """One pass of the hot path.  `record` collects (stage, start_event, end_event) triples."""
        ctx, solver, sharded = self.ctx, self.solver, self.sharded
        T, t_tile = self.T, self.t_tile

        def timed(name, fn):
            if record is None:
                return fn()
            a, b = ctx.ev(), ctx.ev()
            a.record()
            out = fn()
            b.record()
            record.append((name, a, b))
            return out

        timed("build", lambda: solver.build(self.table, out=self.h_buf))

        def eigh():
            solver.diagonalise()
            if sharded is not None:
                sharded.gather_eigenvalues()
            elif ctx.world > 1:
                ctx.dist.all_gather_into_tensor(self.eig_all, solver.eigenvalues.reshape(-1))

        timed("eigh", eigh)
        if self.with_position:
            # eigenvectors -> in-cell position basis + Bloch twiddle, once per eigensolve (K4 on N-point cells)
            timed("fold", (lambda: sharded.fold()) if sharded is not None else (lambda: solver.fold(out=self.folded_buf)))
        c = timed("project", lambda: solver.project(self.psi0))
        if sharded is not None:
            # Position output at N > 1 (pipeline.TimeShardedEvolver): the ranks exchange the folded eigenvectors on a
            # side stream and each evolves its own T/G time samples of EVERY block, then runs the local cell FFT.
            sharded.start_exchange(c, wrap=timed)
            sharded.evolve_position(self.times_local, self.states, self.pos_out, self.dt_local, wrap=timed)
            last_t = (sharded.rows - 1) // t_tile * t_tile          # first row of the last evolve tile
            o_tile = min(self.o_tile, sharded.rows - last_t)
            last0 = last_t + (sharded.rows - last_t - 1) // o_tile * o_tile  # ... of its last position chunk
            self.last_tile = (last0, sharded.rows - last0)  # LOCAL row range; global indices through time_index
            return
        for t0 in range(0, T, t_tile):
            tt = min(t_tile, T - t0)
            if self.with_position:
                timed("evolve", lambda: solver.evolve_cell(c, self.times[t0 : t0 + tt], out=self.states[:tt],
                                                           uniform_dt=self.dt_uniform))
                timed("position", lambda: solver.cell_to_position(self.states[:tt], out=self.pos_out[:tt]))
            else:
                timed("evolve", lambda: solver.evolve_bloch(c, self.times[t0 : t0 + tt], out=self.states[:tt],
                                                            uniform_dt=self.dt_uniform))
            self.last_tile = (t0, tt)

    def measure(self, steps: int, warmup: int, clocks: bool = True) -> dict:
        from slate_quantum_b200 import kernels

        ctx, torch = self.ctx, self.ctx.torch
        for _ in range(warmup):
            self.step()
        ctx.sync_all()
        assert int(self.solver.info.abs().sum().item()) == 0, "eigensolver reported non-convergence"
        records = [[] for _ in range(steps)]
        launches_before = kernels.LAUNCHES
        start, stop = ctx.ev(), ctx.ev()
        sampler = ClockSampler(ctx.local_rank) if clocks else None
        if sampler:
            sampler.__enter__()
        try:
            ctx.sync_all()
            start.record()
            for k in range(steps):
                self.step(records[k])
            stop.record()
            ctx.sync_all()
        finally:
            if sampler:
                sampler.__exit__()
        launches = (kernels.LAUNCHES - launches_before) // max(steps, 1)
        stage_ms = {nm: 0.0 for nm in self.STAGES}
        for rec in records:
            for nm, a_ev, b_ev in rec:
                stage_ms[nm] += a_ev.elapsed_time(b_ev) / steps
        vals = ctx.max_over_ranks([start.elapsed_time(stop) / steps] + [stage_ms[nm] for nm in self.STAGES])
        ms_per_step = vals[0]
        stage_ms = dict(zip(self.STAGES, vals[1:]))
        del torch
        return {
            "ms_per_step": ms_per_step,
            "value": self.n_k_global / (ms_per_step * 1e-3),
            "stage_ms": stage_ms,
            "gpu_launches": int(launches),
            "clocks": sampler.summary() if sampler else None,
        }

    def evolve_route_ms(self, route: str, reps: int = 2) -> float:
        """K3c alone over all T rows through one route of the uniform-grid evolve (SQB_EVOLVE = gemm | nufft), on the
        eigensystem the last step left in the solver; N = 1 only.  Overwrites `states` (not `pos_out`)."""
        ctx, solver = self.ctx, self.solver
        T, t_tile = self.T, self.t_tile
        os.environ["SQB_EVOLVE"] = route
        try:
            c = solver.project(self.psi0)

            def once():
                for t0 in range(0, T, t_tile):
                    tt = min(t_tile, T - t0)
                    if self.with_position:
                        solver.evolve_cell(c, self.times[t0 : t0 + tt], out=self.states[:tt], uniform_dt=self.dt_uniform)
                    else:
                        solver.evolve_bloch(c, self.times[t0 : t0 + tt], out=self.states[:tt], uniform_dt=self.dt_uniform)

            once()
            ctx.sync_all()
            a, b = ctx.ev(), ctx.ev()
            a.record()
            for _ in range(reps):
                once()
            b.record()
            ctx.sync_all()
            return a.elapsed_time(b) / reps
        finally:
            os.environ.pop("SQB_EVOLVE", None)

    def parity_check(self) -> dict:
        """Sampled blocks of the FIRST and LAST time row this rank holds (the last tile of the last step), against
        the CPU oracle; max over ranks.  At N > 1 the sampled blocks lie in several ranks' k-ranges and every rank
        checks its own time rows, so the eigenvector exchange and the time sharding are both covered."""
        ctx = self.ctx
        t0, tt = self.last_tile
        buf = self.pos_out if self.with_position else self.states
        pick = sorted({0, tt - 1})
        rows = buf[pick].cpu().numpy()
        t_idx = [int(self.time_index[t0 + r]) for r in pick] if self.gathered else [t0 + r for r in pick]
        if self.with_position:
            err, blocks = oracle_gate(self.cfg, self.repeat_global, self.psi_np, self.times_np[t_idx], rows, "position")
        else:
            # k-sharded Bloch-order output: this rank's blocks only
            from oracle import bloch_oracle as o

            b0, cnt = self.solver.block_begin, self.solver.block_count
            blocks = sorted({b0, b0 + cnt // 2, b0 + cnt - 1})
            want = o.evolve_sampled_blocks(self.cfg.vectors(), self.cfg.shape, self.cfg.potential_components(),
                                           self.cfg.mass(), self.repeat_global, self.psi_np, self.times_np[t_idx],
                                           blocks, HBAR)
            got = rows.reshape(len(pick), cnt, self.n)[:, [b - b0 for b in blocks], :]
            err = float(np.abs(got - want).max())
        err_all = ctx.max_over_ranks([err])[0]
        return {
            "max_err": err_all,
            "tol": PARITY_TOL,
            "ok": bool(err_all < PARITY_TOL),
            "checked": "sampled k-blocks of the first and last time row of every rank's last time tile vs the CPU "
            "oracle (np.linalg.eigh + np.exp + matmul per block; rows brought back through np.fft + the K5 permutation)",
            "time_rows_rank0": t_idx,
            "blocks_rank0": blocks,
            "basis": "position" if self.with_position else "bloch momentum (k-sharded)",
        }

    def close(self) -> None:
        torch = self.ctx.torch
        for name in ("states", "pos_out", "h_buf", "folded_buf", "eig_all", "sharded", "solver"):
            setattr(self, name, None)
        torch.cuda.empty_cache()


# ----------------------------------------------------------------------------------------------------
def api_objects(cfg: BlochWorkload):
    """Host-side objects of the reference-shaped API for `cfg` (metadata, potential builder, bases)."""
    from slate_quantum_b200 import operator
    from slate_quantum_b200.core import FundamentalBasis, basis, metadata
    from slate_quantum_b200.core.metadata import Domain
    from slate_quantum_b200.metadata import SpacedTimeMetadata, repeat_volume_metadata

    vectors = cfg.vectors()
    cell_meta = metadata.volume.spaced_volume_metadata_from_stacked_delta_x(
        tuple(vectors[i] for i in range(cfg.nd)), cfg.shape
    )
    super_meta = repeat_volume_metadata(cell_meta, cfg.repeat)
    position_basis = basis.from_metadata(super_meta).upcast()
    # "Fourier" spacing = start + delta j / n, i.e. cfg.times() (linspace without the end point)
    times = FundamentalBasis(SpacedTimeMetadata(cfg.n_times, domain=Domain(delta=cfg.t_max_over_hbar * HBAR),
                                                interpolation="Fourier"))
    build_potential = {"cos": operator.build.cos_potential, "fcc": operator.build.fcc_potential}[cfg.potential]
    return cell_meta, super_meta, position_basis, times, build_potential


def e2e_api(ctx: Ctx, cfg: BlochWorkload, steps: int) -> dict:
    """End to end through the reference-shaped calls with HOST buffers, N = 1:

        bloch.build.kinetic_hamiltonian -> dynamics.solve_schrodinger_equation_decomposition ->
        StateList.with_state_basis(position basis)     (reference dynamics/schrodinger.py:59-73, state/_state.py:153-165)

    psi0 comes from pinned host memory; every evolved position-basis state is copied back to pinned host staging
    buffers (host_tiles) inside the timed region."""
    from slate_quantum_b200 import bloch, dynamics, kernels
    from slate_quantum_b200.state import State

    torch = ctx.torch
    cell_meta, _, position_basis, times, build_potential = api_objects(cfg)
    psi_np = cfg.initial_state()
    psi_h = torch.from_numpy(psi_np).pin_memory()
    T, m = cfg.n_times, cfg.m
    want_rows = sorted({0, T // 2 - 1, T - 1})
    kept: dict[int, np.ndarray] = {}
    n_tiles = [0]

    def step(keep: bool) -> None:
        potential = build_potential(cell_meta, cfg.height)
        hamiltonian = bloch.build.kinetic_hamiltonian(potential, cfg.mass(), cfg.repeat)
        psi0 = State(position_basis, psi_h)
        evolution = dynamics.solve_schrodinger_equation_decomposition(psi0, times, hamiltonian)
        in_position = evolution.with_state_basis(position_basis)
        n_tiles[0] = 0
        for t0, rows in in_position.host_tiles():
            n_tiles[0] += 1
            if keep:
                for r in want_rows:
                    if t0 <= r < t0 + rows.shape[0]:
                        kept[r] = rows[r - t0].numpy().copy()

    step(False)
    ctx.sync_all()
    n_e2e = max(1, min(steps, 2))
    launches_before = kernels.LAUNCHES
    a, b = ctx.ev(), ctx.ev()
    a.record()
    for i in range(n_e2e):
        step(i == n_e2e - 1)
    b.record()
    ctx.sync_all()
    ms = a.elapsed_time(b) / n_e2e
    launches = (kernels.LAUNCHES - launches_before) // n_e2e
    rows = np.stack([kept[r] for r in want_rows])
    err, blocks = oracle_gate(cfg, cfg.repeat, psi_np, cfg.times()[want_rows], rows, "position")
    torch.cuda.empty_cache()
    return {
        "value": cfg.n_k / (ms * 1e-3),
        "unit": UNIT,
        "h2d_bytes_per_step": int(psi_h.numel() * 16 + T * 8 + cfg.potential_components().size * 16),
        "d2h_bytes_per_step": int(T * m * 16),
        "ms_per_step": ms,
        "steps": n_e2e,
        "gpu_launches": int(launches),
        "host_chunks_per_step": n_tiles[0],
        "path": "slate_quantum_b200.bloch.build.kinetic_hamiltonian -> dynamics.solve_schrodinger_equation_decomposition"
        " -> StateList.with_state_basis(position basis) -> host_tiles()",
        "note": "host potential + pinned psi0 + times -> the reference-shaped API -> every evolved POSITION-basis state "
        f"copied back to pinned host staging buffers (PCIe-bound: the state list is {T * m * 16 / 1e9:.1f} GB)",
        "parity_check": {
            "max_err": err, "tol": PARITY_TOL, "ok": bool(err < PARITY_TOL), "time_rows": want_rows, "blocks": blocks,
            "checked": "host copies of the API's position-basis states vs the CPU oracle on sampled k-blocks",
        },
    }


def e2e_api_observables(ctx: Ctx, cfg: BlochWorkload, steps: int) -> dict:
    """The same API chain when the caller wants operator.measure observables of every evolved state instead of the
    states: the lazy list is reduced tile by tile on the GPU (K6), [T] doubles per observable come back."""
    from slate_quantum_b200 import bloch, dynamics
    from slate_quantum_b200.operator import measure
    from slate_quantum_b200.state import State

    torch = ctx.torch
    cell_meta, _, position_basis, times, build_potential = api_objects(cfg)
    psi_h = torch.from_numpy(cfg.initial_state()).pin_memory()
    out: dict[str, np.ndarray] = {}

    def step() -> None:
        potential = build_potential(cell_meta, cfg.height)
        hamiltonian = bloch.build.kinetic_hamiltonian(potential, cfg.mass(), cfg.repeat)
        evolution = dynamics.solve_schrodinger_equation_decomposition(State(position_basis, psi_h), times, hamiltonian)
        # the first request walks the lazy list ONCE, tile by tile, and reduces <x>, the periodic <x> and the variance
        # of every axis while each tile is resident; the later requests read the cached [T] results
        for axis in range(cfg.nd):
            out[f"x{axis}"] = measure.all_x(evolution, axis=axis).raw_data
            out[f"width{axis}"] = measure.all_coherent_width(evolution, axis=axis).raw_data

    step()
    ctx.sync_all()
    n_e2e = max(1, min(steps, 2))
    a, b = ctx.ev(), ctx.ev()
    a.record()
    for _ in range(n_e2e):
        step()
    b.record()
    ctx.sync_all()
    ms = a.elapsed_time(b) / n_e2e
    torch.cuda.empty_cache()
    return {
        "value": cfg.n_k / (ms * 1e-3),
        "unit": UNIT,
        "ms_per_step": ms,
        "d2h_bytes_per_step": int(sum(v.size for v in out.values()) * 8),
        "note": "same API chain, but operator.measure.all_x / all_coherent_width of every evolved state along every "
        "axis are returned instead of the states (the lazy list is reduced tile by tile, never materialised)",
        "x0_first_last": [float(np.real(out["x0"][0])), float(np.real(out["x0"][-1]))],
    }


def e2e_sharded(ctx: Ctx, run: PositionRun, steps: int) -> dict:
    """End to end at N > 1 with HOST buffers: the same deliverable as `value` - POSITION-basis states, time-sharded
    (rank r returns time rows [r T/G, (r+1) T/G) of the whole supercell) - plus an observables-only leg."""
    from slate_quantum_b200 import kernels

    torch, cfg, solver, sharded = ctx.torch, run.cfg, run.solver, run.sharded
    n, T = run.n, run.T
    row_bytes = run.row_bytes
    m_global = row_bytes // 16
    rows_local = sharded.rows
    t_tile = max(1, min(rows_local, (8 << 30) // row_bytes))
    host_rows = max(1, min(t_tile, (1 << 30) // row_bytes))
    n_stage = 4
    run.states = run.pos_out = None
    torch.cuda.empty_cache()
    states = torch.empty((t_tile, m_global), dtype=torch.complex128, device="cuda")
    outs = [torch.empty((t_tile, m_global), dtype=torch.complex128, device="cuda") for _ in range(2)]
    stage = [torch.empty((host_rows, m_global), dtype=torch.complex128).pin_memory() for _ in range(n_stage)]
    eig_host = torch.empty((run.n_k_local * n,), dtype=torch.float64).pin_memory()
    copy_stream = torch.cuda.Stream()
    big = solver.supercell
    vec = cfg.vectors()
    lengths = [float(np.linalg.norm(vec[a])) * run.repeat_global[a] for a in range(cfg.nd)]
    mom_host = torch.empty((cfg.nd, rows_local, 5), dtype=torch.float64).pin_memory()
    kept: dict[int, np.ndarray] = {}

    def step(mode: str, keep: bool = False) -> None:
        cur = torch.cuda.current_stream()
        tb = run.table_h.cuda(non_blocking=True)
        ps = run.psi_h.cuda(non_blocking=True)
        tm = run.times_h.cuda(non_blocking=True)
        solver.build(tb, out=run.h_buf)
        solver.diagonalise()
        sharded.gather_eigenvalues()
        eig_host.copy_(solver.eigenvalues.reshape(-1), non_blocking=True)
        sharded.fold()
        c = solver.project(ps)
        sharded.start_exchange(c)
        k = [0]

        def drain(t0: int, tt: int, tile):
            if mode == "observables":
                for a in range(cfg.nd):
                    mom = kernels.measure_moments(tile, big, a, lengths[a] / big[a], lengths[a])
                    mom_host[a, t0 : t0 + tt].copy_(mom, non_blocking=True)
                return None
            ready = torch.cuda.Event()
            ready.record(cur)
            copy_stream.wait_event(ready)
            with torch.cuda.stream(copy_stream):
                for r0 in range(0, tt, host_rows):
                    rr = min(host_rows, tt - r0)
                    stage[k[0] % n_stage][:rr].copy_(tile[r0 : r0 + rr], non_blocking=True)
                    k[0] += 1
                done = torch.cuda.Event()
                done.record(copy_stream)
            if keep and t0 + tt == rows_local:
                done.synchronize()
                last = (k[0] - 1) % n_stage
                rr = tt - (tt - 1) // host_rows * host_rows
                kept[int(run.time_index[rows_local - 1])] = stage[last][rr - 1].numpy().copy()
            return done

        sharded.evolve_position(tm[sharded.time_rows(run.interleaved)].contiguous(), states, outs, run.dt_local,
                                on_tile=drain)
        cur.wait_stream(copy_stream)

    result = {}
    for mode in ("states", "observables"):
        step(mode)
        ctx.sync_all()
        n_e2e = max(1, min(steps, 2))
        a, b = ctx.ev(), ctx.ev()
        a.record()
        for i in range(n_e2e):
            step(mode, keep=(mode == "states" and i == n_e2e - 1))
        b.record()
        ctx.sync_all()
        ms = ctx.max_over_ranks([a.elapsed_time(b) / n_e2e])[0]
        result[mode] = ms
    t_idx = sorted(kept)
    err, blocks = oracle_gate(cfg, run.repeat_global, run.psi_np, run.times_np[t_idx], np.stack([kept[t] for t in t_idx]),
                              "position")
    err = ctx.max_over_ranks([err])[0]
    h2d = run.table_h.numel() * 16 + run.psi_h.numel() * 16 + run.times_h.numel() * 8
    return {
        "value": run.n_k_global / (result["states"] * 1e-3),
        "unit": UNIT,
        "h2d_bytes_per_step": int(h2d),
        "d2h_bytes_per_step": int(eig_host.numel() * 8 + rows_local * row_bytes),
        "ms_per_step": result["states"],
        "basis": "position (supercell), time-sharded: rank r returns time rows "
        + ("r, r + G, r + 2G, ..." if run.interleaved else "[r T/G, (r+1) T/G)") + " - the deliverable `value` measures",
        "numa_node_rank0": ctx.numa_node,
        "note": "pinned host inputs -> C ABI (pipeline.TimeShardedEvolver) -> eigenvalues and this rank's evolved "
        f"position-basis states copied back through {n_stage} x {host_rows * row_bytes / 2**30:.2f} GiB pinned staging "
        f"buffers allocated on the GPU's NUMA node ({rows_local * row_bytes / 1e9:.1f} GB per rank)",
        "parity_check": {"max_err": err, "tol": PARITY_TOL, "ok": bool(err < PARITY_TOL), "blocks": blocks,
                         "checked": "every rank's LAST time row, read from its host staging buffer, vs the CPU oracle"},
        "observables_only": {
            "value": run.n_k_global / (result["observables"] * 1e-3),
            "unit": UNIT,
            "ms_per_step": result["observables"],
            "d2h_bytes_per_step": int(eig_host.numel() * 8 + mom_host.numel() * 8),
            "note": "same host-buffer step, but norm / <x> / <x^2> / scattering phase (K6) of every evolved state "
            "along every axis are returned instead of the states",
        },
    }


# ----------------------------------------------------------------------------------------------------
def run_other_config(ctx: Ctx, cfg: BlochWorkload, steps: int = 1) -> dict:
    """A short k-sharded run of another BASELINE.json config: build + eigh (+ evolution over the config's time batch
    in the Bloch momentum basis), blocks split over the ranks, with the oracle gate on sampled blocks."""
    from slate_quantum_b200 import kernels
    from slate_quantum_b200.pipeline import BlochSolver

    torch = ctx.torch
    begin, count = shard_blocks(cfg.n_k, ctx.rank, ctx.world)
    n, T = cfg.n, cfg.n_times
    solver = BlochSolver(cfg.shape, cfg.repeat, cfg.dk(), cfg.mass(), HBAR, begin, count)
    table = kernels.to_device(cfg.potential_table().ravel().copy(), torch.complex128)
    psi_np = cfg.initial_state()
    psi0 = kernels.to_device(psi_np, torch.complex128)
    times_np = cfg.times()
    times = kernels.to_device(times_np, torch.float64)
    dt = kernels.uniform_step(times_np)
    h_buf = torch.empty((count, n, n), dtype=torch.complex128, device="cuda")
    free = kernels.free_bytes()
    row_bytes = count * n * 16
    t_tile = max(1, min(T, int(free * 0.25) // row_bytes))
    if t_tile >= 512:
        t_tile = t_tile // 512 * 512
    states = torch.empty((t_tile, count * n), dtype=torch.complex128, device="cuda")
    rec: dict[str, float] = {}

    def step(record: bool) -> None:
        evs = [ctx.ev() for _ in range(5)]
        evs[0].record()
        solver.build(table, out=h_buf)
        evs[1].record()
        solver.diagonalise()
        evs[2].record()
        c = solver.project(psi0)
        evs[3].record()
        for t0 in range(0, T, t_tile):
            tt = min(t_tile, T - t0)
            solver.evolve_bloch(c, times[t0 : t0 + tt], out=states[:tt], uniform_dt=dt)
        evs[4].record()
        if record:
            torch.cuda.synchronize()
            for i, nm in enumerate(("build", "eigh", "project", "evolve")):
                rec[nm] = rec.get(nm, 0.0) + evs[i].elapsed_time(evs[i + 1]) / steps
            rec["total"] = rec.get("total", 0.0) + evs[0].elapsed_time(evs[4]) / steps

    if cfg.n_k * n**3 < 1e11:  # cfg1 / cfg2 take a few ms per step: average several (still far below a second)
        steps = max(steps, 5)
    step(False)
    ctx.sync_all()
    for _ in range(steps):
        step(True)
    ctx.sync_all()
    assert int(solver.info.abs().sum().item()) == 0, "eigensolver reported non-convergence"
    names = ("build", "eigh", "project", "evolve", "total")
    vals = dict(zip(names, ctx.max_over_ranks([rec[k] for k in names])))
    # oracle gate: this rank's first / middle / last block at the last tile's first and last time
    from oracle import bloch_oracle as o

    t0 = (T - 1) // t_tile * t_tile
    tt = T - t0
    pick = sorted({0, tt - 1})
    rows = states[pick].cpu().numpy().reshape(len(pick), count, n)
    blocks = sorted({begin, begin + count // 2, begin + count - 1})
    want = o.evolve_sampled_blocks(cfg.vectors(), cfg.shape, cfg.potential_components(), cfg.mass(), cfg.repeat, psi_np,
                                   times_np[[t0 + r for r in pick]], blocks, HBAR)
    err = float(np.abs(rows[:, [b - begin for b in blocks], :] - want).max())
    w_ref = np.linalg.eigvalsh(o.bloch_blocks(cfg.vectors(), cfg.shape, cfg.potential_components(), cfg.mass(),
                                              cfg.repeat, hbar=HBAR, block_begin=blocks[-1], block_count=1))[0]
    w_err = float(np.abs(solver.eigenvalues[blocks[-1] - begin].cpu().numpy() - w_ref).max() / np.abs(w_ref).max())
    err, w_err = ctx.max_over_ranks([err, w_err])
    flops_eigh = (68.0 / 3.0) * n**3 * count
    flops_evolve = 8.0 * n * n * T * count
    out = {
        "workload": cfg.name,
        "k_blocks": cfg.n_k,
        "k_blocks_per_gpu": count,
        "block_size_N": n,
        "n_times": T,
        "value": cfg.n_k / (vals["total"] * 1e-3),
        "unit": UNIT,
        "ms_per_step": vals["total"],
        "steps": steps,
        "stage_ms": {k: vals[k] for k in ("build", "eigh", "project", "evolve")},
        "eigh_tflops": flops_eigh / (vals["eigh"] * 1e-3) / 1e12,
        "evolve_tflops": flops_evolve / (vals["evolve"] * 1e-3) / 1e12,
        "basis": "bloch momentum (k-sharded, no data-path collective)",
        "parity_check": {"max_err": err, "tol": PARITY_TOL, "eigenvalue_rel_err": w_err,
                         "ok": bool(err < PARITY_TOL and w_err < 1e-10), "blocks_rank0": blocks},
    }
    del states, h_buf, solver
    torch.cuda.empty_cache()
    return out


# ----------------------------------------------------------------------------------------------------
def run_ours(args) -> None:
    ctx = Ctx()
    torch, dist, lib = ctx.torch, ctx.dist, ctx.lib
    from slate_quantum_b200 import kernels

    world, rank = ctx.world, ctx.rank
    cfg = CONFIGS[args.config]
    with_position = bool(args.position)
    strong = args.scaling == "strong"
    n, T = cfg.n, cfg.n_times

    # ---- measured FP64 tensor (DMMA) and FMA peaks for the roofline denominators
    sink = torch.zeros(8, dtype=torch.float64, device="cuda")
    fl = ctypes.c_double(0.0)

    def peak(fn, *lead) -> float:
        best = 0.0
        st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        for _ in range(2):
            fn(*lead, 20000, ctypes.c_void_p(sink.data_ptr()), ctypes.byref(fl), st)
        torch.cuda.synchronize()
        for _ in range(4):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn(*lead, 20000, ctypes.c_void_p(sink.data_ptr()), ctypes.byref(fl), st)
            b.record()
            torch.cuda.synchronize()
            best = max(best, fl.value / a.elapsed_time(b) / 1e9)
        return best

    run = PositionRun(ctx, cfg, strong, with_position)
    main = run.measure(args.steps, args.warmup)
    # roofline denominators, measured now that the steps have brought the clocks up
    dmma_variants = [peak(lib.sqb_peak_dmma, v) for v in range(3)]  # 4 / 8 / 16 chains per warp
    dmma_peak = max(dmma_variants)
    dfma_peak = peak(lib.sqb_peak_dfma)
    parity = run.parity_check()
    # the tensor-core GEMM route of the same stage (north_star kernel 3), timed on the same eigensystem: what the NUFFT
    # route replaced on uniform grids, and what still runs for general grids, N > 256 and short tiles
    gemm_route_ms = None
    if world == 1 and run.dt_uniform is not None and not os.environ.get("SQB_EVOLVE"):
        if kernels.evolve_uniform_route(n, run.n_k_local, min(run.t_tile, T)) == "nufft":
            gemm_route_ms = run.evolve_route_ms("gemm")
    # opt-in time-reversal eigensolve (pipeline.BlochSolver.time_reversal: blocks of half the k-grid are diagonalised,
    # the rest mirrored from their partners), timed beside the default on the same run; N = 1 only - at N > 1 a rank's
    # contiguous block range does not contain its partners
    time_reversal = None
    if world == 1 and not args.no_time_reversal:
        run.solver.time_reversal = True
        try:
            m2 = run.measure(min(args.steps, 3), 1, clocks=False)
            if run.solver._time_reversal_applies():  # noqa: SLF001
                time_reversal = {
                    "value": m2["value"],
                    "unit": UNIT,
                    "ms_per_step": m2["ms_per_step"],
                    "stage_ms": m2["stage_ms"],
                    "gpu_launches": m2["gpu_launches"],
                    "parity_check": run.parity_check(),
                    "blocks_diagonalised": (cfg.repeat[0] // 2 + 1) * (cfg.n_k // cfg.repeat[0]),
                    "note": "SQB_TIME_REVERSAL=1 / BlochSolver.time_reversal (off by default): orthogonal lattice + "
                    "Hermitian-symmetric potential table => H_{-k}[pi g, pi g'] = conj(H_k[g, g']); only the blocks of "
                    "half the k-grid go through the eigensolver, the others take eigenvalues and permuted, conjugated "
                    "eigenvectors from their partners (sqb_bloch_time_reversal_mirror)",
                }
        finally:
            run.solver.time_reversal = False
    stage_ms, ms_per_step, value = main["stage_ms"], main["ms_per_step"], main["value"]
    t_tile, n_k_local, n_k_global = run.t_tile, run.n_k_local, run.n_k_global
    run_o_tile = run.o_tile
    run_interleaved = bool(getattr(run, "interleaved", False))
    repeat_global = run.repeat_global
    exchange_mode = run.sharded.exchange_mode if run.sharded is not None else None
    time_sharded = run.sharded is not None

    # ---- K6 (SURVEY 8f "next" row, outside the step): observables of the position-basis states of the last time
    # tile, one pass over the tile; reported beside the step, extrapolated to all T samples
    observables = None
    if with_position and world == 1:
        big = cfg.supercell
        length0 = float(np.linalg.norm(cfg.vectors()[0])) * cfg.repeat[0]
        tt = min(t_tile, T)
        kernels.measure_moments(run.pos_out[:tt], big, 0, length0 / big[0], length0)
        torch.cuda.synchronize()
        a_ev, b_ev = ctx.ev(), ctx.ev()
        a_ev.record()
        mom = kernels.measure_moments(run.pos_out[:tt], big, 0, length0 / big[0], length0)
        b_ev.record()
        torch.cuda.synchronize()
        obs_ms = a_ev.elapsed_time(b_ev)
        obs_gbs = tt * n_k_local * n * 16 / (obs_ms * 1e-3) / 1e9
        observables = {
            "kernel": "measure_moments_kernel (norm, <x>, <x^2>, scattering phase of every state in one pass)",
            "ms_for_all_times": obs_ms * T / tt,
            "bound": "hbm",
            "achieved": obs_gbs,
            "peak": _hbm_peak(),
            "unit": "GB/s",
            "frac": obs_gbs / _hbm_peak(),
            "d2h_bytes": T * 5 * 8,
            "note": "read-only pass: `peak` is the driver's copy figure (a read+write mix), which a pure read can exceed",
            "norm_drift_max": float((mom[:, 0] - mom[0, 0]).abs().max().item()),
        }

    # ---- end to end with HOST buffers
    e2e = None
    if not args.no_e2e:
        if world == 1:
            run.close()
            e2e = e2e_api(ctx, cfg, args.steps)
            e2e["observables_only"] = e2e_api_observables(ctx, cfg, args.steps)
        elif run.sharded is not None:
            e2e = e2e_sharded(ctx, run, args.steps)
    run.close()

    # ---- the OTHER scaling mode in the same line.  Default: `value` is strong scaling of the config itself
    # (north_star: "cfg3 ... scales >= 7x at 8 GPUs") and `weak_scaling` the k-grid enlarged to N x the blocks.
    second_line = None
    if world > 1 and not (args.no_strong or args.no_weak) and cfg.n_k % world == 0 and T % world == 0:
        srun = PositionRun(ctx, cfg, not strong, with_position)
        sm = srun.measure(args.steps, args.warmup, clocks=False)
        sp = srun.parity_check()
        second_line = {
            "value": sm["value"],
            "unit": UNIT,
            "ms_per_step": sm["ms_per_step"],
            "stage_ms": sm["stage_ms"],
            "gpu_launches": sm["gpu_launches"],
            "k_blocks_total": srun.n_k_global,
            "k_blocks_per_gpu": srun.n_k_local,
            "time_rows_per_gpu": T // world,
            "time_tile": srun.t_tile,
            "position_tile": srun.o_tile,
            "parity_check": sp,
            "note": (
                f"{cfg.name}'s own k-grid split over {world} GPUs (same kernels and exchange as the weak run); "
                "speed-up = this value / the 1-GPU value of the same bench"
                if not strong else
                f"every GPU owns {cfg.name}'s k-grid (k-grid enlarged along axis 0 to {world} x the blocks; same kernels "
                "and exchange as the strong run); efficiency = this value / (N x the 1-GPU value); the cell FFT of the "
                "enlarged block grid takes one pass per axis instead of the one-pass cluster kernel of 64 x 64 grids"
            ),
        }
        srun.close()

    # ---- the other BASELINE.json configs, short k-sharded runs with the oracle gate
    others = None
    if not args.no_others:
        others = {}
        for name in sorted(CONFIGS):
            if name == args.config:
                continue
            try:
                others[name] = run_other_config(ctx, CONFIGS[name])
            except Exception as err:  # noqa: BLE001  (reported, never hidden: a failing config shows in the line)
                others[name] = {"error": repr(err)}
                torch.cuda.empty_cache()

    if rank == 0:
        flops_evolve = 8.0 * n * n * T * n_k_local
        flops_eigh = (68.0 / 3.0) * n**3 * n_k_local
        ach = flops_evolve / (stage_ms["evolve"] * 1e-3) / 1e12
        uniform = kernels.uniform_step(cfg.times()) is not None
        rows_launch = min(t_tile, T)
        nufft = uniform and kernels.evolve_uniform_route(n, n_k_local, rows_launch) == "nufft"
        kernel_name = "evolve_nufft_kernel" if nufft else ("evolve_bloch_3m_kernel" if uniform else "evolve_bloch_kernel")
        traffic, traffic_src = ncu_traffic(kernel_name, cfg.name, rows_launch) if world == 1 else (None, None)
        # time rows one rank evolves per step (time-sharded over the ranks when the position stage follows)
        rows_step = T // world if (world > 1 and time_sharded) else T
        n_k_evolved = n_k_global if (world > 1 and time_sharded) else n_k_local
        rows_launch = min(rows_launch, rows_step)
        launches_per_step = -(-rows_step // rows_launch)
        launch_ms = stage_ms["evolve"] / launches_per_step
        if nufft:
            # executed flop of the NUFFT route per output element, counted from csrc/evolve_nufft.cuh at N = 256: gather
            # 1088 (quad, source) pairs x 16 + phases 1536 + radix-16 stage 64 x 258 + radix-8 stage 128 x 98 + pruned
            # last stage 128 x 56 = 55 k flop per column and tile of 512 rows
            flop_per_output = (n * 4.25 * 16 + n * 6 + 64 * 258 + 128 * 98 + 128 * 56) / 512.0
            exec_flops = flop_per_output * n * n_k_evolved * 512.0 * -(-rows_launch // 512)
            out_bytes = 16.0 * n * rows_launch * n_k_evolved
            roofline_kernel = {
                "kernel": "evolve_nufft_kernel (K3c on a uniform time grid as a type-1 non-uniform FFT: window gather of "
                "the N eigenstates onto 1024 grid points + one shared-memory FFT per column and 512-row tile; "
                "csrc/evolve_nufft.cuh)",
                "bound": "tensor",
                "achieved": ach,
                "peak": dmma_peak,
                "unit": "TFLOP/s",
                "frac": ach / dmma_peak,
                "normalisation": "SURVEY 8d useful-work figure 8 N^2 T flop per block (the complex GEMM [T x N] . [N x N] the "
                "reference evaluates) over the measured launch time: the NUFFT route executes ~110 flop per output "
                "element instead of 8 N = 2048, which is why `frac` exceeds 1 - the resources this kernel actually "
                "loads are in `executed` (FP64 pipe) and `hbm` (its remaining bound: the states it must write)",
                "executed": {
                    "flops_per_launch": exec_flops,
                    "flop_per_output_element": flop_per_output,
                    "achieved": exec_flops / (launch_ms * 1e-3) / 1e12,
                    "peak": dfma_peak,
                    "frac": exec_flops / (launch_ms * 1e-3) / 1e12 / dfma_peak,
                    "note": "plain FP64 adds / multiplies / FMAs of the gather and the radix-16/8/8 butterflies (no tensor "
                    "cores: B200's DMMA and DFMA peaks are equal); ncu: shared-memory wavefronts ~55-60 % of peak, FP64 "
                    "pipe ~28 %",
                },
                "hbm": {
                    "bound": "hbm",
                    "achieved": (out_bytes + 16.0 * n * n * n_k_evolved) / (launch_ms * 1e-3) / 1e9,
                    "peak": _hbm_peak(),
                    "unit": "GB/s",
                    "frac": (out_bytes + 16.0 * n * n * n_k_evolved) / (launch_ms * 1e-3) / 1e9 / _hbm_peak(),
                    "note": "algorithmic bytes of SURVEY 8d: 16 N T written per block + the eigenvectors read once",
                },
                "limiter": {
                    "resource": "shared-memory pipe (the grid lines of the gather and of the in-place FFT live in shared "
                    "memory: neither the FP64 pipes nor HBM bound this kernel)",
                    "pct_of_peak": (ncu_metric(kernel_name, cfg.name, rows_launch,
                                               "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed")
                                    if world == 1 else None),
                    "source": "same committed ncu capture as `traffic` (null when no capture matches this launch shape)",
                },
                "traffic": traffic,
                "traffic_source": traffic_src,
                "traffic_unit": "bytes per launch (ncu dram__bytes_read.sum + dram__bytes_write.sum of the committed "
                "--set full capture of this kernel at this config and time tile; null when no capture matches)",
                "algorithmic_bytes_per_launch": out_bytes + 16.0 * n * n * n_k_evolved,
                "launch_ms": launch_ms,
                "peak_source": "best of the FP64 DMMA issue-rate micro-benchmark variants (sqb_peak_dmma, "
                f"4/8/16 chains per warp: {[round(v, 1) for v in dmma_variants]} TFLOP/s) measured in this run after "
                "the timed steps; MEASURED_PEAKS.json carries only bf16 and HBM - the FP64 tensor peak it lacks is "
                "ncu's sm__ops_path_tensor_src_fp64.sum.peak_sustained (18 944 op/cycle x 1.965 GHz = 37.2 TFLOP/s, "
                f"profiles/r01_s2_evolve3m_cfg3_raw.csv), which this micro-benchmark reproduces; plain DFMA {dfma_peak:.1f}; "
                "HBM: MEASURED_PEAKS.json copy figure",
                "algorithmic_flops_per_launch": flops_evolve * rows_launch / T,
            }
            if gemm_route_ms is not None:
                g_ach = flops_evolve / (gemm_route_ms * 1e-3) / 1e12
                roofline_kernel["gemm_route"] = {
                    "kernel": "evolve_bloch_3m_kernel (SQB_EVOLVE=gemm: fused phase x eigenvector complex GEMM, 3M product, "
                    "DMMA.8x8x4; also the route of general grids, N > 256 and tiles under 192 rows)",
                    "stage_ms": gemm_route_ms,
                    "bound": "tensor",
                    "achieved": g_ach,
                    "peak": dmma_peak,
                    "unit": "TFLOP/s",
                    "frac": g_ach / dmma_peak,
                    "executed_frac": 0.75 * g_ach / dmma_peak,
                    "note": "same eigensystem, same run, timed after the steps; the default route is "
                    f"{gemm_route_ms / stage_ms['evolve']:.2f}x faster",
                }
        else:
            roofline_kernel = {
                "kernel": f"{kernel_name} (K3c fused phase x eigenvector complex GEMM"
                + (", 3M product" if uniform else "") + ", DMMA.8x8x4)",
                "bound": "tensor",
                "achieved": ach,
                "peak": dmma_peak,
                "unit": "TFLOP/s",
                "frac": ach / dmma_peak,
                # the 3M complex product executes 6 N^2 T real flop for the 8 N^2 T algorithmic flop of the
                # normalisation, so `frac` (algorithmic / peak) can exceed 1; the tensor-pipe utilisation is this:
                "executed": {
                    "flops_per_launch": (0.75 if uniform else 1.0) * flops_evolve * rows_launch / T,
                    "achieved": (0.75 if uniform else 1.0) * ach,
                    "frac": (0.75 if uniform else 1.0) * ach / dmma_peak,
                    "note": "3 real DMMA products per complex MAC (3M) instead of 4: executed = 0.75 x algorithmic flop",
                },
                "traffic": traffic,
                "traffic_source": traffic_src,
                "traffic_unit": "bytes per launch (ncu dram__bytes_read.sum + dram__bytes_write.sum of the committed "
                "--set full capture of this kernel at this config and time tile; null when no capture matches)",
                # states written + eigenvectors read once (+ their Re+Im plane for the 3M kernel)
                "algorithmic_bytes_per_launch": (16.0 * n * rows_launch + (24.0 if uniform else 16.0) * n * n) * n_k_local,
                "peak_source": "best of the FP64 DMMA issue-rate micro-benchmark variants (sqb_peak_dmma, "
                f"4/8/16 chains per warp: {[round(v, 1) for v in dmma_variants]} TFLOP/s) measured in this run after "
                "the timed steps; MEASURED_PEAKS.json carries only bf16 and HBM - the FP64 tensor peak it lacks is "
                "ncu's sm__ops_path_tensor_src_fp64.sum.peak_sustained (18 944 op/cycle x 1.965 GHz = 37.2 TFLOP/s, "
                f"profiles/r01_s2_evolve3m_cfg3_raw.csv), which this micro-benchmark reproduces; plain DFMA {dfma_peak:.1f}",
                "algorithmic_flops_per_launch": flops_evolve * rows_launch / T,
            }
        line = {
            "metric": METRIC,
            "value": value,
            "unit": UNIT,
            "n_gpus": world,
            "steps": args.steps,
            "warmup": args.warmup,
            "ms_per_step": ms_per_step,
            "higher_is_better": True,
            "scaling": "strong" if strong else "weak",
            "vs_baseline": None,
            "dtype": "c128 (f64)",
            "data": "synthetic",
            # identical (keys and values) to the `config` the reference arm prints for the same flags
            "config": config_dict(cfg, world, strong, with_position),
            "execution": {
                "time_tile": t_tile,
                "position_tile": run_o_tile,
                "time_sharding": (("interleaved: rank r evolves time rows r, r + N, ..." if run_interleaved else
                                   "contiguous: rank r evolves time rows [r T/N, (r+1) T/N)") if world > 1 and time_sharded
                                  else "none"),
                "collectives": ("NCCL all-gather of eigenvalues and coefficients" if world > 1 else "none")
                + (f"; eigenvector exchange ({exchange_mode}: "
                   + ("peer copies over NVLink from symmetric memory" if exchange_mode == "p2p"
                      else "ring of NCCL send/recv")
                   + ") + time-sharded evolve before the local cell FFT" if exchange_mode is not None else ""),
            },
            "blocks_diagonalised_per_s": n_k_global / ((stage_ms["build"] + stage_ms["eigh"]) * 1e-3),
            "state_time_evolutions_per_s": T
            * world
            / ((stage_ms["fold"] + stage_ms["project"] + stage_ms["evolve"] + stage_ms["exchange"] + stage_ms["position"]) * 1e-3),
            "stage_ms": stage_ms,
            "parity_check": parity,
            "roofline": roofline_kernel,
            "roofline_stages": {
                "eigh": {
                    "bound": "tensor (FP64)",
                    "achieved": flops_eigh / (stage_ms["eigh"] * 1e-3) / 1e12,
                    "peak": dmma_peak,
                    "unit": "TFLOP/s",
                    "frac": flops_eigh / (stage_ms["eigh"] * 1e-3) / 1e12 / dmma_peak,
                    "normalisation": "(68/3) N^3 real flop per block (SURVEY 8d)",
                },
                "build": {
                    "bound": "hbm",
                    "achieved": n_k_local * n * n * 16 / (stage_ms["build"] * 1e-3) / 1e9,
                    "peak": _hbm_peak(),
                    "unit": "GB/s",
                    "frac": n_k_local * n * n * 16 / (stage_ms["build"] * 1e-3) / 1e9 / _hbm_peak(),
                },
                "whole_step_fp64": {
                    "achieved": (flops_evolve + flops_eigh) / (ms_per_step * 1e-3) / 1e12,
                    "peak": dmma_peak,
                    "unit": "TFLOP/s",
                    "frac": (flops_evolve + flops_eigh) / (ms_per_step * 1e-3) / 1e12 / dmma_peak,
                },
            },
            "gpu_launches": main["gpu_launches"],
            "clocks": main["clocks"],
        }
        if second_line is not None:
            line["weak_scaling" if strong else "strong_scaling"] = second_line
        if time_reversal is not None:
            line["time_reversal_eigensolve"] = time_reversal
        if observables is not None:
            line["observables"] = observables
        if e2e is not None:
            line["e2e"] = e2e
        if others is not None:
            line["other_configs"] = others
        if world == 1 and not args.no_cpu:
            nb = 64 if n <= 256 else 16
            line["cpu_baseline"] = dict(cpu_sample(cfg, nb, 256, with_position))
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--config", default="cfg3", choices=sorted(CONFIGS))
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--position", type=int, default=1, help="include the cell FFT to the position basis")
    ap.add_argument("--scaling", default="strong", choices=["weak", "strong"],
                    help="what `value` is at N > 1.  strong (default): the config's own k-grid split over the GPUs; weak: "
                    "every GPU owns the config's k-grid.  The line also carries the other figure (`weak_scaling` / "
                    "`strong_scaling`)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-time-reversal", action="store_true", help="skip the extra run with the time-reversal eigensolve")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-strong", action="store_true", help="skip the second scaling mode (same as --no-weak)")
    ap.add_argument("--no-weak", action="store_true", help="skip the second scaling mode")
    ap.add_argument("--no-others", action="store_true", help="skip the short runs of the other BASELINE.json configs")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3  # timing rule: W >= 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
