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

    # ------------------------------------------------------------------ K1
    def build(self, potential_table: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
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
            free, _ = torch.cuda.mem_get_info()
            floor = _lib.load().sqb_eigh_workspace_bytes(self.n, 1)
            self._eigh_ws = torch.empty(max(min(want, int(free * 0.5)), floor), dtype=torch.uint8, device="cuda")
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
