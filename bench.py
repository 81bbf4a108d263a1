#!/usr/bin/env python
"""bench.py -- one "step" = one pass of the hot path over one synthetic Bloch workload:

    K1 build every k-block  ->  K2 eigh of every block  ->  K4+K5+K3a project psi0  ->
    K3c evolve over T times (fused phase x eigenvector DMMA GEMM)  ->  [N=1: K5+K4 to the position basis]

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config cfg3] [--impl ours|reference]

Prints ONE JSON line (rank 0).  Multi-GPU: one process per GPU (torchrun), k-blocks sharded in contiguous
ranges, weak scaling (each rank owns prod(repeat) blocks of a k-grid enlarged along axis 0), the only
collective is the NCCL all-gather of eigenvalues.  See DESIGN.md section "Measurement".
"""

from __future__ import annotations

import argparse
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

from slate_quantum_b200.configs import CONFIGS, HBAR, BlochWorkload  # noqa: E402

METRIC = "bloch_k_blocks_per_s"
UNIT = "k-blocks/s (each block built, diagonalised and evolved over the config's time batch)"


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


# ----------------------------------------------------------------------------------------------------
# CPU arm: the oracle (numpy restatement of the reference; the reference itself cannot be imported here)
def cpu_sample(cfg: BlochWorkload, n_blocks: int, n_times: int, with_position: bool) -> dict:
    """Time the oracle on a bounded sample of `cfg`; extrapolate linearly to one full step."""
    from oracle import bloch_oracle as o

    try:
        from threadpoolctl import threadpool_info

        blas_threads = max((p.get("num_threads", 1) for p in threadpool_info()), default=1)
    except Exception:  # noqa: BLE001
        blas_threads = 1
    cores = len(os.sched_getaffinity(0))
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
    cfg = CONFIGS[args.config]
    # bounded sample per step so that K+W steps end within a few minutes
    nb = 64 if cfg.n <= 256 else 16
    nt = 256
    for _ in range(args.warmup):
        cpu_sample(cfg, 2, 8, args.position)
    t0 = time.perf_counter()
    vals = [cpu_sample(cfg, nb, nt, args.position) for _ in range(args.steps)]
    wall = time.perf_counter() - t0
    best = max(vals, key=lambda d: d["value"])
    line = {
        "impl": "reference",
        "metric": METRIC,
        "value": best["value"],
        "unit": UNIT,
        "n_gpus": args.gpus,
        "steps": args.steps,
        "warmup": args.warmup,
        "ms_per_step": best["step_s_extrapolated"] * 1e3,
        "higher_is_better": True,
        "scaling": args.scaling,
        "vs_baseline": None,
        "dtype": "c128 (f64)",
        "data": "synthetic",
        "config": {"workload": cfg.name, "position_basis": bool(args.position)},
        "cpu_baseline": {k: best[k] for k in ("value", "unit", "cores", "blas_threads", "kind", "sample")},
        "e2e": {"value": best["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "numpy/LAPACK restatement of the reference (reference needs Python>=3.14 + un-vendored slate_core); "
        f"sample wall {wall:.1f}s",
    }
    print(json.dumps(line))


# ----------------------------------------------------------------------------------------------------
def run_ours(args) -> None:
    import torch
    import torch.distributed as dist

    from slate_quantum_b200 import _lib, kernels
    from slate_quantum_b200.pipeline import BlochSolver, TimeShardedEvolver

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device; slate_quantum_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    lib = _lib.load()
    cfg = CONFIGS[args.config]
    with_position = bool(args.position)

    # weak scaling (default): the k-grid grows along axis 0 with the number of GPUs, rank r owns n_k contiguous
    # blocks.  strong scaling (--scaling strong): the config's own k-grid is split into `world` contiguous ranges.
    strong = args.scaling == "strong"
    if strong and cfg.n_k % world:
        raise SystemExit(f"--scaling strong needs n_k={cfg.n_k} divisible by the number of GPUs")
    repeat_global = tuple(cfg.repeat) if strong else (cfg.repeat[0] * world, *cfg.repeat[1:])
    n, T = cfg.n, cfg.n_times
    n_k_local = cfg.n_k // world if strong else cfg.n_k
    n_k_global = n_k_local * world
    solver = BlochSolver(cfg.shape, repeat_global, cfg.dk(), cfg.mass(), HBAR, rank * n_k_local, n_k_local)
    supercell = solver.supercell
    m_global = solver.m

    # ---- host inputs (pinned) -> device
    table_h = torch.from_numpy(cfg.potential_table().ravel().copy()).pin_memory()
    wl_global = BlochWorkload(cfg.name, cfg.delta_x, cfg.shape, repeat_global, cfg.potential, cfg.height, T, cfg.t_max_over_hbar)
    psi_h = torch.from_numpy(wl_global.initial_state()).pin_memory()
    times_h = torch.from_numpy(cfg.times()).pin_memory()
    table = table_h.cuda(non_blocking=True)
    psi0 = psi_h.cuda(non_blocking=True)
    times = times_h.cuda(non_blocking=True)
    dt_uniform = kernels.uniform_step(cfg.times())  # SpacedTimeMetadata-style uniform grid -> phase recurrence

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

    # ---- buffers
    h_buf = torch.empty((n_k_local, n, n), dtype=torch.complex128, device="cuda")
    ws_bytes = lib.sqb_eigh_workspace_bytes(n, n_k_local)
    solver._eigh_ws = torch.empty(ws_bytes, dtype=torch.uint8, device="cuda")  # noqa: SLF001
    free, _ = torch.cuda.mem_get_info()
    row_bytes = n_k_local * n * 16
    t_tile = T
    # buffers of one time tile: states (+ position output)
    gathered = world > 1 and with_position  # time-sharded evolve over ring-exchanged eigenvectors
    sharded = TimeShardedEvolver(solver, T) if gathered else None  # pipeline.py; DESIGN.md section 6
    folded_buf = torch.empty_like(h_buf) if with_position and not gathered else None
    free, _ = torch.cuda.mem_get_info()
    if gathered:
        row_bytes = n_k_global * n * 16  # a rank evolves T / world time samples of EVERY block
        t_tile = T // world
    budget = int(free * 0.8 / (2 if with_position else 1))
    while t_tile * row_bytes > budget and t_tile > 64:
        t_tile //= 2
    states = torch.empty((t_tile, row_bytes // 16), dtype=torch.complex128, device="cuda")
    pos_out = torch.empty_like(states) if with_position else None
    eig_all = torch.empty((world, n_k_local * n), dtype=torch.float64, device="cuda") if world > 1 and not gathered else None

    ev = lambda: torch.cuda.Event(enable_timing=True)  # noqa: E731
    stage_names = ["build", "eigh", "fold", "project", "evolve", "exchange", "position"]

    def step(record: list | None = None) -> None:
        """One pass of the hot path.  `record` collects (stage, start_event, end_event) triples."""

        def timed(name, fn):
            if record is None:
                return fn()
            a, b = ev(), ev()
            a.record()
            out = fn()
            b.record()
            record.append((name, a, b))
            return out

        timed("build", lambda: solver.build(table, out=h_buf))

        def eigh():
            solver.diagonalise()
            if sharded is not None:
                sharded.gather_eigenvalues()
            elif world > 1:
                dist.all_gather_into_tensor(eig_all, solver.eigenvalues.reshape(-1))

        timed("eigh", eigh)
        if with_position:
            # eigenvectors -> in-cell position basis + Bloch twiddle, once per eigensolve (K4 on N-point cells)
            timed("fold", (lambda: sharded.fold()) if sharded is not None else (lambda: solver.fold(out=folded_buf)))
        c = timed("project", lambda: solver.project(psi0))
        if world > 1 and with_position:
            # Position output at N > 1 (pipeline.TimeShardedEvolver): the ranks ring-exchange the folded eigenvectors
            # on a side stream and each evolves its own T/G time samples of EVERY block, then runs the local cell FFT.
            sharded.start_exchange(c, wrap=timed)
            t_lo = sharded.time_begin
            sharded.evolve_position(times[t_lo : t_lo + sharded.rows], states, pos_out, dt_uniform, wrap=timed)
            return
        for t0 in range(0, T, t_tile):
            tt = min(t_tile, T - t0)
            if with_position:
                timed("evolve", lambda: solver.evolve_cell(c, times[t0 : t0 + tt], out=states[:tt], uniform_dt=dt_uniform))
                timed("position", lambda: solver.cell_to_position(states[:tt], out=pos_out[:tt]))
            else:
                timed("evolve", lambda: solver.evolve_bloch(c, times[t0 : t0 + tt], out=states[:tt], uniform_dt=dt_uniform))

    def sync_all() -> None:
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    sync_all()
    assert int(solver.info.abs().sum().item()) == 0, "eigensolver reported non-convergence"
    # roofline denominators, measured now that the warm-up steps have brought the clocks up
    dmma_variants = [peak(lib.sqb_peak_dmma, v) for v in range(3)]  # 4 / 8 / 16 chains per warp
    dmma_peak = max(dmma_variants)
    dfma_peak = peak(lib.sqb_peak_dfma)

    # ---- device-resident timed region: exactly K steps
    records = [[] for _ in range(args.steps)]
    launches_before = kernels.LAUNCHES
    start, stop = ev(), ev()
    with ClockSampler(local_rank) as clocks:
        sync_all()
        start.record()
        for k in range(args.steps):
            step(records[k])
        stop.record()
        sync_all()
    launches = (kernels.LAUNCHES - launches_before) // max(args.steps, 1)
    total_ms = start.elapsed_time(stop)
    stage_ms = {nm: 0.0 for nm in stage_names}
    for rec in records:
        for nm, a_ev, b_ev in rec:
            stage_ms[nm] += a_ev.elapsed_time(b_ev) / args.steps
    ms_per_step = total_ms / args.steps
    if world > 1:
        tmax = torch.tensor([ms_per_step], dtype=torch.float64, device="cuda")
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        ms_per_step = float(tmax.item())
        st = torch.tensor([stage_ms[nm] for nm in stage_names], dtype=torch.float64, device="cuda")
        dist.all_reduce(st, op=dist.ReduceOp.MAX)
        stage_ms = dict(zip(stage_names, st.tolist()))
    value = n_k_global / (ms_per_step * 1e-3)

    # ---- K6 (SURVEY 8f "next" row, outside the step): observables of the position-basis states of the last time
    # tile, one pass over the tile; reported beside the step, extrapolated to all T samples
    observables = None
    if with_position and world == 1:
        big = cfg.supercell
        length0 = float(np.linalg.norm(cfg.vectors()[0])) * cfg.repeat[0]
        tt = min(t_tile, T)
        kernels.measure_moments(pos_out[:tt], big, 0, length0 / big[0], length0)
        torch.cuda.synchronize()
        a_ev, b_ev = ev(), ev()
        a_ev.record()
        mom = kernels.measure_moments(pos_out[:tt], big, 0, length0 / big[0], length0)
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

    # ---- end-to-end through the C ABI with HOST buffers (H2D of inputs, D2H of eigenvalues and states)
    e2e = None
    if not args.no_e2e:
        copy_stream = torch.cuda.Stream()
        # self-contained buffers (the device-resident leg's big tiles are released first)
        del states, pos_out
        torch.cuda.empty_cache()
        row_e = n_k_local * n * 16
        e_tile = max(1, min(T // 2, (4 << 30) // row_e))          # two device tiles ping-pong
        d2h_rows = max(1, min(e_tile, (1 << 30) // row_e))        # <= 1 GiB pinned staging buffers, ping-pong
        e2e_position = with_position and world == 1  # at N > 1 the e2e leg returns k-sharded momentum-basis states
        e_src = torch.empty((2 * e_tile, n_k_local * n), dtype=torch.complex128, device="cuda")
        pos_tmp = torch.empty((e_tile, n_k_local * n), dtype=torch.complex128, device="cuda") if e2e_position else None
        stage_buf = [torch.empty((d2h_rows, n_k_local * n), dtype=torch.complex128).pin_memory() for _ in range(2)]
        eig_host = torch.empty((n_k_local * n,), dtype=torch.float64).pin_memory()
        tile_free = [torch.cuda.Event(), torch.cuda.Event()]
        h2d = table_h.numel() * 16 + psi_h.numel() * 16 + times_h.numel() * 8
        d2h = eig_host.numel() * 8 + T * row_e

        def e2e_step() -> None:
            cur = torch.cuda.current_stream()
            tb = table_h.cuda(non_blocking=True)
            ps = psi_h.cuda(non_blocking=True)
            tm = times_h.cuda(non_blocking=True)
            solver.build(tb, out=h_buf)
            solver.diagonalise()
            eig_host.copy_(solver.eigenvalues.reshape(-1), non_blocking=True)
            if e2e_position:
                solver.fold(out=folded_buf)
            c = solver.project(ps)
            k = 0
            for i_tile, t0 in enumerate(range(0, T, e_tile)):
                tt = min(e_tile, T - t0)
                buf = e_src[(i_tile & 1) * e_tile : (i_tile & 1) * e_tile + tt]
                cur.wait_event(tile_free[i_tile & 1])  # the D2H of the tile that used this buffer is done
                if e2e_position:
                    solver.evolve_cell(c, tm[t0 : t0 + tt], out=pos_tmp[:tt], uniform_dt=dt_uniform)
                    solver.cell_to_position(pos_tmp[:tt], out=buf)
                else:
                    solver.evolve_bloch(c, tm[t0 : t0 + tt], out=buf, uniform_dt=dt_uniform)
                ready = torch.cuda.Event()
                ready.record(cur)
                copy_stream.wait_event(ready)
                with torch.cuda.stream(copy_stream):
                    for r0 in range(0, tt, d2h_rows):
                        rr = min(d2h_rows, tt - r0)
                        stage_buf[k & 1][:rr].copy_(buf[r0 : r0 + rr], non_blocking=True)
                        k += 1
                    tile_free[i_tile & 1].record(copy_stream)
            cur.wait_stream(copy_stream)

        e2e_step()
        sync_all()
        a, b = ev(), ev()
        a.record()
        n_e2e = max(1, min(args.steps, 2))
        for _ in range(n_e2e):
            e2e_step()
        b.record()
        sync_all()
        e2e_ms = a.elapsed_time(b) / n_e2e
        if world > 1:
            tmax = torch.tensor([e2e_ms], dtype=torch.float64, device="cuda")
            dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
            e2e_ms = float(tmax.item())
        e2e = {
            "value": n_k_global / (e2e_ms * 1e-3),
            "unit": UNIT,
            "h2d_bytes_per_step": int(h2d),
            "d2h_bytes_per_step": int(d2h),
            "ms_per_step": e2e_ms,
            "steps": n_e2e,
            "note": "host potential table + psi0 + times in pinned memory -> C ABI -> eigenvalues and every "
            "evolved state copied back to pinned host staging buffers (PCIe-bound: the state list is "
            f"{T * row_e / 1e9:.1f} GB per rank)",
        }
        if e2e_position:
            # Same step when the caller wants OBSERVABLES of the evolved states instead of the states (what the
            # reference's examples plot): operator.measure's moments (K6) are reduced on the GPU and [T, 5]
            # doubles per axis come back instead of the state list.  Extra figure, not the contract's `e2e`.
            big = cfg.supercell
            lengths = [float(np.linalg.norm(cfg.vectors()[a])) * cfg.repeat[a] for a in range(cfg.nd)]
            mom_host = torch.empty((cfg.nd, T, 5), dtype=torch.float64).pin_memory()

            def e2e_observables_step() -> None:
                tb = table_h.cuda(non_blocking=True)
                ps = psi_h.cuda(non_blocking=True)
                tm = times_h.cuda(non_blocking=True)
                solver.build(tb, out=h_buf)
                solver.diagonalise()
                eig_host.copy_(solver.eigenvalues.reshape(-1), non_blocking=True)
                solver.fold(out=folded_buf)
                c = solver.project(ps)
                for t0 in range(0, T, e_tile):
                    tt = min(e_tile, T - t0)
                    solver.evolve_cell(c, tm[t0 : t0 + tt], out=pos_tmp[:tt], uniform_dt=dt_uniform)
                    pos = solver.cell_to_position(pos_tmp[:tt], out=e_src[:tt])
                    for a in range(cfg.nd):
                        mom = kernels.measure_moments(pos, big, a, lengths[a] / big[a], lengths[a])
                        mom_host[a, t0 : t0 + tt].copy_(mom, non_blocking=True)

            e2e_observables_step()
            sync_all()
            a, b = ev(), ev()
            a.record()
            for _ in range(n_e2e):
                e2e_observables_step()
            b.record()
            sync_all()
            obs_ms = a.elapsed_time(b) / n_e2e
            e2e["observables_only"] = {
                "value": n_k_global / (obs_ms * 1e-3),
                "unit": UNIT,
                "ms_per_step": obs_ms,
                "h2d_bytes_per_step": int(h2d),
                "d2h_bytes_per_step": int(eig_host.numel() * 8 + mom_host.numel() * 8),
                "note": "same host-buffer step, but norm / <x> / <x^2> / scattering phase of every evolved state along "
                "every axis (operator.measure, K6) are returned instead of the states",
            }

    if rank == 0:
        flops_evolve = 8.0 * n * n * T * n_k_local
        flops_eigh = (68.0 / 3.0) * n**3 * n_k_local
        ach = flops_evolve / (stage_ms["evolve"] * 1e-3) / 1e12
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
            "config": {
                "workload": cfg.name,
                "shape": list(cfg.shape),
                "repeat_per_gpu": list(cfg.repeat),
                "repeat_global": list(repeat_global),
                "block_size_N": n,
                "k_blocks_per_gpu": n_k_local,
                "n_times": T,
                "position_basis": with_position,
                "time_tile": t_tile,
                "l2_policy": "inputs larger than L2: every step rebuilds and re-reads the "
                f"{n_k_local * n * n * 16 / 1e9:.2f} GB block array (L2 = 126 MB)",
                "parallelism": f"k-block sharding x{world}"
                + (", NCCL all-gather of eigenvalues" if world > 1 else "")
                + (f", eigenvector exchange ({sharded.exchange_mode}: "
                   + ("peer copies over NVLink from symmetric memory" if sharded.exchange_mode == "p2p"
                      else "ring of NCCL send/recv")
                   + ") + time-sharded evolve before the local cell FFT" if sharded is not None else ""),
            },
            "blocks_diagonalised_per_s": n_k_global / ((stage_ms["build"] + stage_ms["eigh"]) * 1e-3),
            "state_time_evolutions_per_s": T
            * world
            / ((stage_ms["fold"] + stage_ms["project"] + stage_ms["evolve"] + stage_ms["exchange"] + stage_ms["position"]) * 1e-3),
            "stage_ms": stage_ms,
            "roofline": {
                "kernel": "evolve_bloch_3m_kernel (K3c fused phase x eigenvector complex GEMM, 3M product, DMMA.8x8x4)"
                if dt_uniform is not None
                else "evolve_bloch_kernel (K3c fused phase x eigenvector complex GEMM, DMMA.8x8x4)",
                "bound": "tensor",
                "achieved": ach,
                "peak": dmma_peak,
                "unit": "TFLOP/s",
                "frac": ach / dmma_peak,
                # the 3M complex product executes 6 N^2 T real flop for the 8 N^2 T algorithmic flop of the
                # normalisation, so `frac` (algorithmic / peak) can exceed 1; the tensor-pipe utilisation is this:
                "executed": {
                    "flops_per_launch": (0.75 if dt_uniform is not None else 1.0) * flops_evolve * min(t_tile, T) / T,
                    "achieved": (0.75 if dt_uniform is not None else 1.0) * ach,
                    "frac": (0.75 if dt_uniform is not None else 1.0) * ach / dmma_peak,
                    "note": "3 real DMMA products per complex MAC (3M) instead of 4: executed = 0.75 x algorithmic flop",
                },
                # dram__bytes_read.sum + dram__bytes_write.sum of ONE launch from profiles/r01_s2_evolve3m_cfg3_raw.csv
                # (ncu --set full, cfg3, 2048-sample time tile); null for other shapes (not captured)
                "traffic": 41.518867e9 if (cfg.name.startswith("cfg3") and min(t_tile, T) == 2048 and world == 1) else None,
                "traffic_unit": "bytes per launch (ncu dram__bytes_read.sum + dram__bytes_write.sum)",
                # states written + eigenvectors read once (+ their Re+Im plane for the 3M kernel)
                "algorithmic_bytes_per_launch": (16.0 * n * min(t_tile, T) + (24.0 if dt_uniform is not None else 16.0) * n * n)
                * n_k_local,
                "peak_source": "best of the FP64 DMMA issue-rate micro-benchmark variants (sqb_peak_dmma, "
                f"4/8/16 chains per warp: {[round(v, 1) for v in dmma_variants]} TFLOP/s) measured in this run after "
                "warm-up (MEASURED_PEAKS.json carries only bf16 and HBM); plain DFMA peak measured "
                f"{dfma_peak:.1f} TFLOP/s",
                "algorithmic_flops_per_launch": flops_evolve * min(t_tile, T) / T,
            },
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
            "gpu_launches": int(launches),
            "clocks": clocks.summary(),
        }
        if observables is not None:
            line["observables"] = observables
        if e2e is not None:
            line["e2e"] = e2e
        if world == 1 and not args.no_cpu:
            nb = 64 if n <= 256 else 16
            line["cpu_baseline"] = dict(cpu_sample(cfg, nb, 256, with_position))
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def _hbm_peak() -> float:
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"])
        except Exception:  # noqa: BLE001
            pass
    return 6650.0  # B200_PROFILING.md fallback


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--config", default="cfg3", choices=sorted(CONFIGS))
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--position", type=int, default=1, help="include K5+K4 to the position basis (N=1 only)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: every GPU owns the config's k-grid (default); strong: the config is split over the GPUs")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3  # timing rule: W >= 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
