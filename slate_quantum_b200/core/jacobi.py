"""Dense Hermitian eigensolver for ONE large matrix (n > 512): two-sided block Jacobi on the B200 kernels.

SURVEY section 8f rank 3: the reference's single-cell examples (examples/eigenstates.py:16-35,
examples/schrodinger.py:23-35, dynamics/caldeira_leggett/_periodic.py:244-275) call np.linalg.eigh on one dense
[n, n] operator.  The batched eigensolver (K2, sqb_eigh_batched) keeps a whole matrix on one SM and stops at
n = 512; a single larger matrix is reduced to MANY such problems instead:

    split the index range into blocks of b = 256 columns; in one round of a round-robin tournament the blocks are
    paired (p, q) disjointly, every pair's 2b x 2b Hermitian sub-matrix [[A_pp, A_pq], [A_qp, A_qq]] is
    diagonalised by ONE batched K2 call, and the block unitaries are applied to A (both sides) and accumulated
    into V with batched complex GEMMs on the FP64 tensor cores (K7, sqb_zgemm_batched).

Every round annihilates the paired off-diagonal blocks exactly; sweeps converge quadratically (6 - 9 sweeps to
1e-14 of the Frobenius norm).  Sub-problems return ascending eigenvalues, which is the "sorting" variant of the
method.  Padding columns (n not a multiple of 2b) carry distinct diagonal entries far above the spectrum and stay
decoupled.  All arithmetic runs through the C ABI; torch only gathers / scatters blocks and sorts the result.
"""

from __future__ import annotations

import torch

BLOCK = 256
MAX_SWEEPS = 30


def _round_robin(n_blocks: int) -> list[list[tuple[int, int]]]:
    """Rounds of disjoint pairs covering every pair of `n_blocks` (even) blocks exactly once per sweep."""
    ring = list(range(n_blocks))
    rounds = []
    for _ in range(n_blocks - 1):
        rounds.append([(min(ring[i], ring[-1 - i]), max(ring[i], ring[-1 - i])) for i in range(n_blocks // 2)])
        ring = [ring[0], ring[-1], *ring[1:-1]]
    return rounds


def eigh_block_jacobi(a: torch.Tensor, tol: float = 1e-15) -> tuple[torch.Tensor, torch.Tensor, int]:
    """Eigen-decomposition of one Hermitian [n, n] CUDA complex128 matrix.

    Returns (w [n] float64 ascending, transform [n, n] with ROWS = eigenvectors (the K2 layout), sweeps used).
    """
    from slate_quantum_b200 import kernels

    n = a.shape[0]
    b = BLOCK
    n_blocks = -(-n // b)
    n_blocks += n_blocks & 1
    m = n_blocks * b
    scale = float(a.abs().max().item()) or 1.0
    work = torch.zeros((m, m), dtype=torch.complex128, device=a.device)
    work[:n, :n] = a
    if m > n:  # decoupled padding: distinct diagonal entries well above the spectrum (|lambda| <= n * max|a|)
        pad = torch.arange(1, m - n + 1, dtype=torch.float64, device=a.device)
        work[range(n, m), range(n, m)] = ((2.0 * n + 10.0) * scale * (1.0 + pad / (m - n))).to(torch.complex128)
    v = torch.eye(m, dtype=torch.complex128, device=a.device)
    rounds = _round_robin(n_blocks)
    total = float(torch.linalg.matrix_norm(work[:n, :n]).item()) or 1.0
    idx = torch.arange(m, device=a.device).reshape(n_blocks, b)
    sweeps = 0
    for sweeps in range(1, MAX_SWEEPS + 1):
        for pairs in rounds:
            p = torch.tensor([pq[0] for pq in pairs], device=a.device)
            q = torch.tensor([pq[1] for pq in pairs], device=a.device)
            cols = torch.cat([idx[p], idx[q]], dim=1)  # [pairs, 2b] global indices of every pair
            # sub-matrices [pairs, 2b, 2b]
            sub = work[cols[:, :, None], cols[:, None, :]].contiguous()
            sub = 0.5 * (sub + sub.mH)  # exactly Hermitian input for K2 (rounding of the two-sided updates)
            _, u, info = kernels.eigh_batched(sub.contiguous())
            if int(info.abs().sum().item()):
                msg = "block Jacobi: a sub-problem did not converge"
                raise RuntimeError(msg)
            # rows of u are eigenvectors: S = Q diag Q^H with Q = u^T.  A <- Q^H A Q on the paired rows / columns.
            qmat = u.mT  # [pairs, 2b, 2b] strided view, passed to the GEMM as it is
            flat = cols.reshape(-1)
            # columns: A[:, cols] <- A[:, cols] Q
            a_cols = work[:, flat].reshape(m, len(pairs), 2 * b).permute(1, 0, 2).contiguous()
            work[:, flat] = kernels.zgemm(a_cols, qmat).permute(1, 0, 2).reshape(m, -1)
            # rows: A[cols, :] <- Q^H A[cols, :]
            a_rows = work[flat, :].reshape(len(pairs), 2 * b, m)
            work[flat, :] = kernels.zgemm(u.conj(), a_rows).reshape(-1, m)
            # V[:, cols] <- V[:, cols] Q
            v_cols = v[:, flat].reshape(m, len(pairs), 2 * b).permute(1, 0, 2).contiguous()
            v[:, flat] = kernels.zgemm(v_cols, qmat).permute(1, 0, 2).reshape(m, -1)
        diag = torch.diagonal(work).real
        off = torch.linalg.matrix_norm(work[:n, :n] - torch.diag_embed(torch.diagonal(work[:n, :n])))
        if float(off.item()) <= tol * total * n**0.5:
            break
    else:
        msg = f"block Jacobi did not converge in {MAX_SWEEPS} sweeps"
        raise RuntimeError(msg)
    diag = torch.diagonal(work).real[:n] if m == n else torch.diagonal(work).real
    if m > n:
        # padding eigenvalues are the m - n largest and their eigenvectors live in the padding coordinates
        order = torch.argsort(diag)[:n]
    else:
        order = torch.argsort(diag)
    w = diag[order].contiguous()
    transform = v[:n, order].mT.contiguous()  # rows = eigenvectors
    return w, transform, sweeps
