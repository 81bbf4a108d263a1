"""Device-resident driver of the hot path: K1 build -> K2 eigh -> K4/K5/K3a project -> K3c evolve -> K5/K4.

This is the host-side engine the API mirror (slate_quantum_b200.bloch / operator / dynamics) sits on.
It owns nothing but torch CUDA tensors and calls the C ABI for every arithmetic step.  A solver instance
covers a contiguous range of k-blocks, which is how the path shards over GPUs (one process per GPU).
"""

from __future__ import annotations

from typing import Sequence

import numpy as np
import torch

from slate_quantum_b200 import kernels


class BlochSolver:
    """Bloch Hamiltonian blocks [block_begin, block_begin + block_count) of one periodic system.

    Mirrors the reference call chain bloch.build.kinetic_hamiltonian (bloch/build.py:146-167) ->
    operator.into_diagonal_hermitian (operator/_linalg.py:45-83) ->
    dynamics.solve_schrodinger_equation_decomposition (dynamics/schrodinger.py:59-73).
    """

    def __init__(
        self,
        shape: Sequence[int],
        repeat: Sequence[int],
        dk: np.ndarray,
        mass: float,
        hbar: float,
        block_begin: int = 0,
        block_count: int | None = None,
    ) -> None:
        self.shape = tuple(int(s) for s in shape)
        self.repeat = tuple(int(r) for r in repeat)
        self.dk = np.asarray(dk, dtype=np.float64)
        self.mass = float(mass)
        self.hbar = float(hbar)
        self.n = int(np.prod(self.shape))
        self.n_k = int(np.prod(self.repeat))
        self.block_begin = int(block_begin)
        self.block_count = self.n_k - self.block_begin if block_count is None else int(block_count)
        self.supercell = tuple(n * r for n, r in zip(self.shape, self.repeat, strict=True))
        self.m = self.n * self.n_k
        self.hamiltonian: torch.Tensor | None = None
        self.eigenvalues: torch.Tensor | None = None
        self.transform: torch.Tensor | None = None
        self.info: torch.Tensor | None = None
        self._eigh_ws: torch.Tensor | None = None
        self.transform_folded: torch.Tensor | None = None
        self._evolve_ws: torch.Tensor | None = None
        self._evolve_ws_for: tuple | None = None  # (data_ptr of the transform whose Re+Im plane the workspace holds)
        # Opt-in (SQB_TIME_REVERSAL=1 or .time_reversal = True): diagonalise only the blocks of half the k-grid and
        # mirror the rest, see diagonalise().  Needs an orthogonal lattice, the whole k-grid on this solver and a
        # Hermitian-symmetric potential table (checked on the table build() receives).
        import os

        self.time_reversal = os.environ.get("SQB_TIME_REVERSAL", "0") == "1"
        self._table_symmetric: tuple | None = None  # (data_ptr, version, verdict) of the last table checked

    # ------------------------------------------------------------------ K1
    def _time_reversal_applies(self) -> bool:
        off_diagonal = self.dk - np.diag(np.diag(self.dk))
        return (
            self.time_reversal
            and self.block_begin == 0
            and self.block_count == self.n_k
            and self.repeat[0] > 2
            and self.n <= kernels.EIGH_KERNEL_MAX_N
            and float(np.abs(off_diagonal).max()) <= 1e-14 * float(np.abs(self.dk).max())
            and self._table_symmetric is not None
            and self._table_symmetric[2]
        )

    def _check_table(self, table: torch.Tensor) -> None:
        """Is the potential table Hermitian-symmetric, Vt[-q] = conj(Vt[q]) (a real V(x))?  One small device reduction
        and a host read per NEW table; the verdict is cached on (data_ptr, version)."""
        key = (table.data_ptr(), table._version)  # noqa: SLF001
        if self._table_symmetric is not None and self._table_symmetric[:2] == key:
            return
        t = table.reshape(self.shape)
        axes = tuple(range(t.dim()))
        mirrored = torch.roll(torch.flip(t, axes), shifts=(1,) * t.dim(), dims=axes)
        scale = float(t.abs().max().item())
        ok = float((torch.conj_physical(mirrored) - t).abs().max().item()) <= 1e-14 * max(scale, 1e-300)
        self._table_symmetric = (*key, ok)

    def build(self, potential_table: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
        if self.time_reversal:
            self._check_table(potential_table)
        self.hamiltonian = kernels.bloch_build(
            self.shape, self.repeat, self.dk, potential_table, self.mass, self.hbar,
            self.block_begin, self.block_count, out=out,
        )
        return self.hamiltonian

    # ------------------------------------------------------------------ K2
    def diagonalise(self, hamiltonian: torch.Tensor | None = None) -> tuple[torch.Tensor, torch.Tensor]:
        """In place: the Hamiltonian buffer becomes the transform (rows = eigenvectors)."""
        h = self.hamiltonian if hamiltonian is None else hamiltonian
        if h is None:
            msg = "build() first"
            raise RuntimeError(msg)
        if self._eigh_ws is None:
            from slate_quantum_b200 import _lib

            want = _lib.load().sqb_eigh_workspace_bytes(self.n, h.shape[0])
            free = kernels.free_bytes()
            floor = _lib.load().sqb_eigh_workspace_bytes(self.n, 1)
            self._eigh_ws = torch.empty(max(min(want, int(free * 0.5)), floor), dtype=torch.uint8, device="cuda")
        n_rep = (self.repeat[0] // 2 + 1) * (self.n_k // self.repeat[0])
        if hamiltonian is None and self._time_reversal_applies() and n_rep < self.n_k:
            # time reversal: H_{-k}[pi g, pi g'] = conj(H_k[g, g']) (csrc/bloch_build.cu), so the blocks whose first
            # k-index lies in the upper half of its FFT-ordered range take eigenvalues and (permuted, conjugated)
            # eigenvectors from their partners; the partners are the contiguous prefix of the block array
            w_rep, _, info_rep = kernels.eigh_batched(h[:n_rep], workspace=self._eigh_ws)
            w = torch.empty((self.n_k, self.n), dtype=torch.float64, device="cuda")
            w[:n_rep] = w_rep
            info = torch.zeros((self.n_k,), dtype=torch.int32, device="cuda")
            info[:n_rep] = info_rep
            kernels.time_reversal_mirror(self.shape, self.repeat, n_rep, self.n_k - n_rep, h, w)
            v = h
        else:
            w, v, info = kernels.eigh_batched(h, workspace=self._eigh_ws)
        self.eigenvalues, self.transform, self.info = w, v, info
        self.hamiltonian = None
        self.transform_folded = None
        self._evolve_ws_for = None
        return w, v

    def fold(self, out: torch.Tensor | None = None) -> torch.Tensor:
        """Eigenvectors in the in-cell POSITION basis with the Bloch twiddle applied (done once per eigh)."""
        if self.transform_folded is None or out is not None:
            self.transform_folded = kernels.fold_transform(
                self.shape, self.repeat, self.transform, self.block_begin, out=out
            )
            self._evolve_ws_for = None
        return self.transform_folded

    def _evolve(self, transform, coefficients, times, out, uniform_dt):
        """K3c on `transform` with a persistent workspace: the Re+Im plane of the 3M kernel is computed once per
        (eigensolve, transform) and reused by every later time tile."""
        if uniform_dt is None:
            return kernels.evolve_bloch(transform, self.eigenvalues, coefficients, times, self.hbar, out=out)
        need = kernels.evolve_uniform_workspace_bytes(self.n, transform.shape[0], times.numel())
        if self._evolve_ws is None or self._evolve_ws.numel() < need:
            self._evolve_ws = torch.empty(need, dtype=torch.uint8, device="cuda")
            self._evolve_ws_for = None
        key = (transform.data_ptr(),)
        reuse = self._evolve_ws_for == key
        res = kernels.evolve_bloch(
            transform, self.eigenvalues, coefficients, times, self.hbar, out=out, uniform_dt=uniform_dt,
            workspace=self._evolve_ws, reuse_plane=reuse,
        )
        self._evolve_ws_for = key
        return res

    def release_workspace(self) -> None:
        self._eigh_ws = None

    # ------------------------------------------------------------------ (E, V) shard dump / resume
    def _manifest(self) -> dict:
        return {
            "format": "slate_quantum_b200.eigensystem/1",
            "shape": list(self.shape),
            "repeat": list(self.repeat),
            "block_begin": self.block_begin,
            "block_count": self.block_count,
            "block_size": self.n,
            "mass": self.mass,
            "hbar": self.hbar,
            "dk": self.dk.tolist(),
            "layout": "eigenvalues [block_count, N] float64 ascending per block; "
            "transform [block_count, N, N] complex128, rows = eigenvectors (EigenstateBasis transform)",
        }

    def save_eigensystem(self, directory: str, tag: str = "shard") -> str:
        """Write this solver's (E, V) shard: `<tag>.eigenvalues.npy`, `<tag>.transform.npy`, `<tag>.json`.

        SURVEY section 8f rank 2: the reference has no checkpointing; a k-block shard (one per rank) makes the
        eigensolve resumable - evolve / project only need (E, V).  The manifest records the system so that a
        shard cannot be loaded into a different problem, plus a checksum of the eigenvalues.
        """
        import hashlib
        import json
        from pathlib import Path

        if self.eigenvalues is None or self.transform is None:
            msg = "diagonalise() first"
            raise RuntimeError(msg)
        d = Path(directory)
        d.mkdir(parents=True, exist_ok=True)
        w = self.eigenvalues.detach().cpu().numpy()
        np.save(d / f"{tag}.eigenvalues.npy", w)
        np.save(d / f"{tag}.transform.npy", self.transform.detach().cpu().numpy())
        manifest = self._manifest()
        manifest["eigenvalues_sha256"] = hashlib.sha256(np.ascontiguousarray(w).tobytes()).hexdigest()
        (d / f"{tag}.json").write_text(json.dumps(manifest, indent=1))
        return str(d / f"{tag}.json")

    def load_eigensystem(self, directory: str, tag: str = "shard") -> None:
        """Restore (E, V) written by save_eigensystem into this solver (same system and block range)."""
        import hashlib
        import json
        from pathlib import Path

        d = Path(directory)
        manifest = json.loads((d / f"{tag}.json").read_text())
        mine = self._manifest()
        for key in ("format", "shape", "repeat", "block_begin", "block_count", "block_size", "mass", "hbar", "dk"):
            if manifest.get(key) != mine[key]:
                msg = f"eigensystem shard {tag!r} was written for a different problem: {key} differs"
                raise ValueError(msg)
        w = np.load(d / f"{tag}.eigenvalues.npy")
        if hashlib.sha256(np.ascontiguousarray(w).tobytes()).hexdigest() != manifest["eigenvalues_sha256"]:
            msg = f"eigensystem shard {tag!r}: eigenvalue checksum mismatch"
            raise ValueError(msg)
        v = np.load(d / f"{tag}.transform.npy")
        if w.shape != (self.block_count, self.n) or v.shape != (self.block_count, self.n, self.n):
            msg = f"eigensystem shard {tag!r} has the wrong array shapes"
            raise ValueError(msg)
        self.eigenvalues = kernels.to_device(w, torch.float64)
        self.transform = kernels.to_device(np.ascontiguousarray(v), torch.complex128)
        self.info = torch.zeros((self.block_count,), dtype=torch.int32, device="cuda")
        self.hamiltonian = None
        self.transform_folded = None
        self._evolve_ws_for = None

    # ------------------------------------------------------------------ K4 + K5 + K3a
    def to_bloch(self, psi_position: torch.Tensor) -> torch.Tensor:
        """Position-basis state(s) over the whole supercell [.., M] -> Bloch order [.., n_k, N]."""
        flat = psi_position.reshape(-1, self.m)
        psi_k = kernels.fft_c2c(flat, self.supercell, -1)
        return kernels.bloch_permute(self.shape, self.repeat, psi_k, 0).reshape(-1, self.n_k, self.n)

    def project(self, psi_position: torch.Tensor) -> torch.Tensor:
        """c[b, e] for this solver's blocks from ONE position-basis state."""
        psi_b = self.to_bloch(psi_position)[0, self.block_begin : self.block_begin + self.block_count].contiguous()
        return kernels.project(self.transform, psi_b)

    # ------------------------------------------------------------------ K3b / K3c
    def evolve_eigen(self, coefficients: torch.Tensor, times: torch.Tensor) -> torch.Tensor:
        """The reference's return value: [T, count*N] amplitudes in the eigenbasis (schrodinger.py:51-53)."""
        return kernels.evolve_eigen(self.eigenvalues.reshape(-1), coefficients.reshape(-1), times, self.hbar)

    def evolve_bloch(
        self, coefficients: torch.Tensor, times: torch.Tensor, out: torch.Tensor | None = None,
        uniform_dt: float | None = None,
    ) -> torch.Tensor:
        """[T, count*N] states in the Bloch-ordered momentum basis (fused phase x eigenvector GEMM)."""
        return self._evolve(self.transform, coefficients, times, out, uniform_dt)

    def evolve_cell(
        self, coefficients: torch.Tensor, times: torch.Tensor, out: torch.Tensor | None = None,
        uniform_dt: float | None = None,
    ) -> torch.Tensor:
        """[T, count*N] states in the cell layout (block, in-cell POSITION); cell_to_position finishes the job."""
        return self._evolve(self.fold(), coefficients, times, out, uniform_dt)

    def cell_to_position(
        self, states_cell: torch.Tensor, out: torch.Tensor | None = None, ranks: int = 1
    ) -> torch.Tensor:
        """[T, M] cell layout (ALL blocks) -> supercell position basis: R-point FFTs across blocks only.
        `states_cell` is used as scratch.  ranks > 1: `states_cell` is the rank-major all-to-all receive
        buffer [ranks, T, n_k/ranks * N] (see exchange_k_to_time(..., transpose=False))."""
        if states_cell.numel() % self.m:
            msg = "cell_to_position needs every k-block (all-to-all the k-shards into time-shards first)"
            raise ValueError(msg)
        return kernels.cell_fft(self.shape, self.repeat, states_cell, 1, out=out, ranks=ranks)

    def evolve_position(
        self, coefficients: torch.Tensor, times: torch.Tensor, uniform_dt: float | None = None
    ) -> torch.Tensor:
        """Position-basis states [T, M] (needs block_count == n_k)."""
        return self.cell_to_position(self.evolve_cell(coefficients, times, uniform_dt=uniform_dt))

    # ------------------------------------------------------------------ K5 + K4
    def bloch_to_position(self, states_bloch: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
        """[T, M] Bloch order (ALL blocks) -> position basis.  Needs block_count == n_k."""
        if states_bloch.shape[-1] != self.m:
            msg = "bloch_to_position needs every k-block (gather or all-to-all the shards first)"
            raise ValueError(msg)
        states_k = kernels.bloch_permute(self.shape, self.repeat, states_bloch.contiguous(), 1)
        return kernels.fft_c2c(states_k, self.supercell, +1, out=states_k if out is None else out)


def exchange_k_to_time(states: torch.Tensor, group=None, *, transpose: bool = True) -> torch.Tensor:
    """All-to-all re-sharding of evolved states: k-sharded [T, count*N] -> time-sharded [T/G, G*count*N].

    Rank r owns a contiguous range of k-blocks, so concatenating the ranks' block ranges
    in rank order IS the global block order.  After the exchange rank g holds time rows
    [g*T/G, (g+1)*T/G) of EVERY block, which is what the across-block cell FFT (sqb_cell_fft) needs.
    The only data-path collective of the position-basis output; NCCL all_to_all over NVLink (a grouped
    send/recv fallback is used on backends without all_to_all, e.g. gloo in the CPU tests).
    """
    import torch.distributed as dist

    world = dist.get_world_size(group)
    if world == 1:
        return states
    t, width = states.shape
    if t % world:
        msg = f"the time tile ({t}) must be divisible by the number of ranks ({world})"
        raise ValueError(msg)
    rows = t // world
    send = torch.view_as_real(states.contiguous()).reshape(world, rows, width, 2)
    recv = torch.empty_like(send)
    if dist.get_backend(group) == "nccl":
        dist.all_to_all_single(recv, send, group=group)
    else:
        rank = dist.get_rank(group)
        ops = []
        for peer in range(world):
            if peer == rank:
                recv[peer].copy_(send[peer])
            else:
                ops.append(dist.P2POp(dist.isend, send[peer].contiguous(), peer, group))
                ops.append(dist.P2POp(dist.irecv, recv[peer], peer, group))
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    if not transpose:
        # rank-major receive buffer [src rank, rows, width]: sqb_cell_fft_sharded reads it in place
        return torch.view_as_complex(recv)
    # [src rank, rows, width] -> [rows, src rank * width]
    out = torch.view_as_complex(recv.permute(1, 0, 2, 3).contiguous())
    return out.reshape(rows, world * width)


# ---------------------------------------------------------------------------------------------------------------
# Position-basis output on G > 1 GPUs: time-sharded evolve over ring-exchanged eigenvectors
# ---------------------------------------------------------------------------------------------------------------
def ring_segments(rank: int, world: int, blocks_per_rank: int) -> list[tuple[int, int, int]]:
    """(source rank, first global block, block count) in the order rank `rank` evolves them.

    Segment 0 is the rank's own block range (its eigenvectors are local); segment k >= 1 is the range whose
    eigenvectors arrive in step k of ring_exchange, i.e. rank - k (mod world).
    """
    if not 0 <= rank < world:
        msg = f"rank {rank} outside world of {world}"
        raise ValueError(msg)
    return [(((rank - k) % world), ((rank - k) % world) * blocks_per_rank, blocks_per_rank) for k in range(world)]


def ring_exchange(local: torch.Tensor, gathered: torch.Tensor, group=None, on_arrival=None) -> None:
    """gathered[src] <- rank src's `local` for every src != rank, as world-1 paired send/recv steps.

    `gathered` is [world, *local.shape]; row `rank` is NOT written (the caller keeps using `local`).  Step k sends
    `local` to rank + k and receives rank - k's, then calls on_arrival(k) - the GPU path records an event there so
    the evolve of that peer's blocks can start while later steps are still in flight.  Complex tensors travel as
    their real views (NCCL has no complex dtype).
    """
    import torch.distributed as dist

    world, rank = dist.get_world_size(group), dist.get_rank(group)
    if gathered.shape[0] != world or tuple(gathered.shape[1:]) != tuple(local.shape):
        msg = f"gathered must be [{world}, {', '.join(map(str, local.shape))}], got {tuple(gathered.shape)}"
        raise ValueError(msg)
    if local.is_complex():
        local, gathered = torch.view_as_real(local), torch.view_as_real(gathered)
    for k in range(1, world):
        src, dst = (rank - k) % world, (rank + k) % world
        ops = [
            dist.P2POp(dist.isend, local, dst if group is None else dist.get_global_rank(group, dst), group),
            dist.P2POp(dist.irecv, gathered[src], src if group is None else dist.get_global_rank(group, src), group),
        ]
        for req in dist.batch_isend_irecv(ops):
            req.wait()
        if on_arrival is not None:
            on_arrival(k)


def shard_time_rows(rank: int, world: int, n_times: int, interleaved: bool = False) -> slice:
    """Rows of the global time grid that `rank` of `world` evolves (n_times divisible by world): the contiguous range
    [rank T/G, (rank + 1) T/G), or - `interleaved` - rows rank, rank + G, ... (on a uniform grid again a uniform grid,
    step G dt)."""
    if n_times % world:
        msg = f"the number of time samples ({n_times}) must be divisible by the number of ranks ({world})"
        raise ValueError(msg)
    rows = n_times // world
    return slice(rank, n_times, world) if interleaved else slice(rank * rows, (rank + 1) * rows)


class TimeShardedEvolver:
    """Position-basis states at G > 1: every rank evolves T/G time samples of EVERY k-block.

    The across-block cell FFT (K4) needs all blocks of a time sample on one GPU.  Re-sharding the evolved states
    would move T*M*16 B per step; instead the ranks exchange what the states are made FROM - the folded
    eigenvectors (n_k*N^2*16 B, 2x less at cfg3), eigenvalues and coefficients - once per eigensolve, and each
    evolves its own time range for all blocks (same flops per rank).  The exchange runs on a side stream as a
    ring, enqueued before the evolve of the rank's own blocks, and each peer's segment waits only for that
    peer's eigenvectors, so from the second segment on the transfer hides behind the GEMMs.

    `solver` must own block range [rank*c, (rank+1)*c) with c = n_k / world (the weak- and strong-scaling layouts
    of bench.py); `n_times` must be a multiple of world.
    """

    def __init__(self, solver: BlochSolver, n_times: int, group=None, exchange: str | None = None) -> None:
        import torch.distributed as dist

        self.solver, self.group = solver, group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        c, n = solver.block_count, solver.n
        if c * self.world != solver.n_k or solver.block_begin != self.rank * c:
            msg = "TimeShardedEvolver needs equal contiguous block ranges in rank order"
            raise ValueError(msg)
        if n_times % self.world:
            msg = f"the number of time samples ({n_times}) must be divisible by the number of ranks ({self.world})"
            raise ValueError(msg)
        self.rows = n_times // self.world
        self.time_begin = self.rank * self.rows
        self.n_times = n_times
        self.segments = ring_segments(self.rank, self.world, c)
        self.transforms = torch.empty((self.world, c, n, n), dtype=torch.complex128, device="cuda")
        self.coefficients = torch.empty((self.world, c, n), dtype=torch.complex128, device="cuda")
        self.eigenvalues = torch.empty((self.world, c, n), dtype=torch.float64, device="cuda")
        self._ws = [
            torch.empty(kernels.evolve_uniform_workspace_bytes(n, c, self.rows), dtype=torch.uint8, device="cuda")
            for _ in self.segments
        ]
        self.comm_stream = torch.cuda.Stream(priority=-1)
        self._arrived: list = [None] * self.world
        self._local: tuple | None = None
        # Where fold() writes the rank's own folded eigenvectors.  exchange="p2p" (default; env SQB_EXCHANGE): two
        # symmetric-memory buffers (torch.distributed._symmetric_memory: cuMemCreate allocations mapped into every
        # peer's address space over NVLink), so a peer's eigenvectors are fetched with a plain device-to-device copy
        # on the copy engines - no communication kernel competes with the evolve CTAs for SMs.  Two buffers,
        # alternating per eigensolve: the barrier that opens exchange s+1 proves every rank has finished reading
        # the buffers of exchange s, and buffer s%2 is not written again before fold s+2.
        # exchange="nccl": ring of NCCL send/recv pairs (ring_exchange) from an ordinary buffer.
        import os

        want = exchange or os.environ.get("SQB_EXCHANGE", "p2p")
        if want not in ("p2p", "nccl"):
            msg = f"exchange must be 'p2p' or 'nccl', got {want!r}"
            raise ValueError(msg)
        self._symm = None
        self._generation = 0
        if want == "p2p" and self.world > 1:
            try:
                import torch.distributed._symmetric_memory as symm

                dev = torch.device("cuda", torch.cuda.current_device())
                flat = symm.empty((2 * c * n * n * 2,), dtype=torch.float64, device=dev)
                handle = symm.rendezvous(flat, dist.group.WORLD if group is None else group)
                self._symm = (flat, handle)
            except (RuntimeError, ImportError, AttributeError) as err:
                if exchange == "p2p":
                    raise
                import warnings

                self._symm_error = repr(err)
                warnings.warn(  # both transports are GPU paths; say which one runs
                    f"symmetric memory unavailable ({err!r}); TimeShardedEvolver uses the NCCL ring exchange",
                    RuntimeWarning, stacklevel=2,
                )
        self.exchange_mode = "p2p" if self._symm is not None else "nccl"
        if self._symm is not None:
            both = torch.view_as_complex(self._symm[0].view(2, c, n, n, 2))
            self._fold_buffers = [both[0], both[1]]
        else:
            self._fold_buffers = [torch.empty((c, n, n), dtype=torch.complex128, device="cuda")]

    def time_rows(self, interleaved: bool = False) -> slice:
        """Which rows of the global time grid this rank evolves: the contiguous range [time_begin, time_begin + rows),
        or - `interleaved` - every world-th row starting at `rank`.  On a uniform grid the interleaved rows are again a
        uniform grid (step world * dt, pass that as `uniform_dt`); the wider step spreads the angles w dt / hbar of the
        NUFFT evolve over its gather grid when a rank's share is a single 512-row tile (csrc/evolve_nufft.cuh)."""
        return shard_time_rows(self.rank, self.world, self.n_times, interleaved)

    def gather_eigenvalues(self) -> torch.Tensor:
        """All-gather of the eigenvalues ([n_k, N], global block order); part of the eigensolve stage."""
        import torch.distributed as dist

        dist.all_gather_into_tensor(self.eigenvalues, self.solver.eigenvalues.contiguous(), group=self.group)
        return self.eigenvalues.reshape(-1, self.solver.n)

    def fold(self) -> torch.Tensor:
        """solver.fold() into the buffer the exchange reads from (call once per eigensolve, before
        start_exchange)."""
        self._generation += 1
        return self.solver.fold(out=self._fold_buffers[self._generation % len(self._fold_buffers)])

    def start_exchange(self, coefficients: torch.Tensor, wrap=None) -> None:
        """Enqueue the coefficient all-gather and the eigenvector exchange on the side stream (after everything
        already enqueued on the current stream).  `wrap(name, fn)` lets a caller time the enqueue."""
        import torch.distributed as dist

        mine = self._fold_buffers[self._generation % len(self._fold_buffers)]
        folded = self.solver.transform_folded
        if folded is None or folded.data_ptr() != mine.data_ptr():
            msg = "TimeShardedEvolver.fold() first (the exchange reads the evolver's own buffer)"
            raise RuntimeError(msg)
        self._local = (folded, coefficients)
        cur = torch.cuda.current_stream()
        ready = torch.cuda.Event()
        ready.record(cur)
        self.comm_stream.wait_event(ready)

        def arrived(k: int) -> None:
            self._arrived[k] = torch.cuda.Event()
            self._arrived[k].record(self.comm_stream)

        def run() -> None:
            dist.all_gather_into_tensor(
                torch.view_as_real(self.coefficients), torch.view_as_real(coefficients.contiguous()), group=self.group
            )
            if self._symm is None:
                ring_exchange(folded, self.transforms, self.group, arrived)
                return
            _, handle = self._symm
            handle.barrier(channel=0)  # every rank's fold of this generation is complete
            count = folded.numel() * 2
            offset = (self._generation % 2) * count
            for k in range(1, self.world):
                src = (self.rank - k) % self.world
                peer = handle.get_buffer(src, (count,), torch.float64, offset)
                torch.view_as_real(self.transforms[src]).reshape(-1).copy_(peer, non_blocking=True)
                arrived(k)

        with torch.cuda.stream(self.comm_stream):
            run() if wrap is None else wrap("exchange", run)

    def evolve_position(
        self, times_local: torch.Tensor, states: torch.Tensor, out, uniform_dt: float | None,
        on_tile=None, wrap=None,
    ) -> None:
        """Evolve this rank's time samples `times_local` ([rows], = times[self.time_rows(...)], contiguous) for all
        blocks, tile by tile: `states` is a [t_tile, M] scratch buffer (cell layout), `out` a [o_tile <= t_tile, M]
        buffer; on_tile(t0, tt, out[:tt]) is called after each (chunk of a) tile's position-basis states have been
        enqueued.

        `out` may be a LIST of buffers used round-robin (a consumer that drains a tile asynchronously, e.g. a
        device -> host copy on another stream): when on_tile returns a CUDA event, the buffer it was given is not
        written again before that event has completed - only the cell FFT of the later tile waits, its evolve
        (which writes `states`) overlaps the drain."""
        if self._local is None:
            msg = "start_exchange() first"
            raise RuntimeError(msg)
        s, n = self.solver, self.solver.n
        folded, c_local = self._local
        cur = torch.cuda.current_stream()
        call = (lambda _name, fn: fn()) if wrap is None else wrap
        outs = list(out) if isinstance(out, (list, tuple)) else [out]
        busy: list = [None] * len(outs)
        t_tile = states.shape[0]
        i_out = 0
        for i_tile, t0 in enumerate(range(0, times_local.numel(), t_tile)):
            tt = min(t_tile, times_local.numel() - t0)
            tsl = times_local[t0 : t0 + tt]
            for i_seg, (src, b0, cnt) in enumerate(self.segments):
                if i_seg == 0:
                    tr, ww, cc = folded, s.eigenvalues, c_local
                else:
                    if i_tile == 0:
                        cur.wait_event(self._arrived[i_seg])  # that rank's eigenvectors (and coefficients) are in
                    tr, ww, cc = self.transforms[src], self.eigenvalues[src], self.coefficients[src]
                if uniform_dt is None:
                    call("evolve", lambda: kernels.evolve_bloch(
                        tr, ww, cc, tsl, s.hbar, out=states[:tt, b0 * n : (b0 + cnt) * n]))
                else:
                    # the workspace size depends on the route the call takes (a short ragged tile runs the GEMM,
                    # whose Re+Im plane needs far more than the NUFFT plan): grow on demand
                    need = kernels.evolve_uniform_workspace_bytes(n, cnt, tt)
                    reuse = i_tile > 0
                    if self._ws[i_seg].numel() < need:
                        self._ws[i_seg] = torch.empty(need, dtype=torch.uint8, device="cuda")
                        reuse = False
                    ws = self._ws[i_seg]
                    call("evolve", lambda: kernels.evolve_bloch(
                        tr, ww, cc, tsl, s.hbar, out=states[:tt, b0 * n : (b0 + cnt) * n], uniform_dt=uniform_dt,
                        workspace=ws, reuse_plane=reuse))
            # the position-basis buffers may hold fewer rows than the evolve tile (two full tiles of a 512 x 64 block
            # grid do not fit beside the exchanged eigenvectors): the cell FFT then runs chunk by chunk
            o_tile = min(tt, outs[0].shape[0])
            for p0 in range(0, tt, o_tile):
                pt = min(o_tile, tt - p0)
                slot = i_out % len(outs)
                i_out += 1
                if busy[slot] is not None:
                    cur.wait_event(busy[slot])
                    busy[slot] = None
                dst = outs[slot]
                call("position", lambda: s.cell_to_position(states[p0 : p0 + pt], out=dst[:pt]))
                if on_tile is not None:
                    busy[slot] = on_tile(t0 + p0, pt, dst[:pt])
