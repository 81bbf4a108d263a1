"""Numpy model of the GPU eigensolver's algorithm (test helper, runs on CPU).

Mirrors slate_quantum_b200/csrc/eigh.cu step by step: (1) Householder tridiagonalisation in the
column-major view of the row-major input (i.e. of conj(H)) with the deferred rank-2 update,
(2) implicit-shift QL on the real tridiagonal with recorded Givens rotations, (3) rotation replay on
the identity, (4) Householder back-transformation, final conjugation so rows = eigenvectors of H.
"""
from __future__ import annotations

import numpy as np


def tridiagonalise(h: np.ndarray):
    n = h.shape[0]
    a = np.array(h.T, dtype=np.complex128)  # col-major view of row-major memory: A = H^T = conj(H)
    d = np.zeros(n)
    e = np.zeros(n)
    tau = np.zeros(n, dtype=np.complex128)
    y = np.zeros((n, n), dtype=np.complex128)  # y[i, r] reflector i
    v_old = np.zeros(n, dtype=np.complex128)
    w_old = np.zeros(n, dtype=np.complex128)
    for i in range(n - 1):
        # phase A: materialise column i
        col = a[i:, i] - v_old[i:] * np.conj(w_old[i]) - w_old[i:] * np.conj(v_old[i])
        d[i] = col[0].real
        alpha = col[1]
        x = col[2:]
        xnorm = np.linalg.norm(x)
        v_new = np.zeros(n, dtype=np.complex128)
        if xnorm == 0 and alpha.imag == 0:
            t = 0.0
            beta = alpha.real
            v_new[i + 1] = 1
        else:
            beta = -np.copysign(np.sqrt(alpha.real**2 + alpha.imag**2 + xnorm**2), alpha.real)
            t = complex((beta - alpha.real) / beta, -alpha.imag / beta)
            v_new[i + 1] = 1
            v_new[i + 2 :] = x / (alpha - beta)
        e[i] = beta
        tau[i] = t
        y[i] = v_new
        # phase B: materialise trailing block and accumulate A22 v
        blk = a[i + 1 :, i + 1 :]
        blk -= np.outer(v_old[i + 1 :], np.conj(w_old[i + 1 :])) + np.outer(w_old[i + 1 :], np.conj(v_old[i + 1 :]))
        p = t * (blk @ v_new[i + 1 :])
        alpha2 = -0.5 * t * np.vdot(p, v_new[i + 1 :])
        w_new = np.zeros(n, dtype=np.complex128)
        w_new[i + 1 :] = p + alpha2 * v_new[i + 1 :]
        v_old, w_old = v_new, w_new
    last = a[n - 1, n - 1] - v_old[n - 1] * np.conj(w_old[n - 1]) - w_old[n - 1] * np.conj(v_old[n - 1])
    d[n - 1] = last.real
    return d, e, tau, y


def tridiagonalise_blocked(h: np.ndarray, nb: int = 16):
    """Panel-blocked variant of `tridiagonalise` (LAPACK zhetrd / zlatrd, lower triangle), the design candidate for
    the next version of tridiag_kernel.

    Inside a panel of `nb` columns the trailing matrix is only READ: column i is materialised from the panel's
    pending rank-2j update, the matrix-vector product uses the matrix as it was at the panel start and is
    corrected with the panel's V and W (two thin products), and the rank-2nb update A -= V W^H + W V^H is applied
    once per panel (GEMM shaped: the DMMA path).  Same conventions and outputs as `tridiagonalise`.
    Returns (d, e, tau, y, sweeps) with sweeps = (read sweeps, read+write sweeps) over the trailing matrix.
    """
    n = h.shape[0]
    a = np.array(h.T, dtype=np.complex128)
    d = np.zeros(n)
    e = np.zeros(n)
    tau = np.zeros(n, dtype=np.complex128)
    y = np.zeros((n, n), dtype=np.complex128)
    reads = writes = 0
    i0 = 0
    while i0 < n - 1:
        width = min(nb, n - 1 - i0)
        v_panel = np.zeros((n, width), dtype=np.complex128)
        w_panel = np.zeros((n, width), dtype=np.complex128)
        for j in range(width):
            i = i0 + j
            col = a[i:, i] - v_panel[i:, :j] @ np.conj(w_panel[i, :j]) - w_panel[i:, :j] @ np.conj(v_panel[i, :j])
            d[i] = col[0].real
            alpha = col[1]
            x = col[2:]
            xnorm = np.linalg.norm(x)
            v = np.zeros(n, dtype=np.complex128)
            v[i + 1] = 1
            if xnorm == 0 and alpha.imag == 0:
                t = 0.0
                beta = alpha.real
            else:
                beta = -np.copysign(np.sqrt(alpha.real**2 + alpha.imag**2 + xnorm**2), alpha.real)
                t = complex((beta - alpha.real) / beta, -alpha.imag / beta)
                v[i + 2 :] = x / (alpha - beta)
            e[i] = beta
            tau[i] = t
            y[i] = v
            vt, lo = v[i + 1 :], i + 1
            p = a[lo:, lo:] @ vt  # the matrix of the panel start: read only
            reads += 1
            p -= v_panel[lo:, :j] @ (np.conj(w_panel[lo:, :j]).T @ vt) + w_panel[lo:, :j] @ (np.conj(v_panel[lo:, :j]).T @ vt)
            p *= t
            w = p + (-0.5 * t * np.vdot(p, vt)) * vt
            v_panel[:, j] = v
            w_panel[lo:, j] = w
        lo = i0 + width
        a[lo:, lo:] -= v_panel[lo:] @ np.conj(w_panel[lo:]).T + w_panel[lo:] @ np.conj(v_panel[lo:]).T
        writes += 1
        i0 += width
    d[n - 1] = a[n - 1, n - 1].real
    return d, e, tau, y, (reads, writes)


def ql_record(d: np.ndarray, e: np.ndarray, max_iter: int = 60):
    """Implicit QL (EISPACK imtql2 / NR tqli); returns eigenvalues (unsorted), sweeps [(l, m, [(c,s)...])], fails."""
    n = len(d)
    d = d.copy()
    e = e.copy()
    e[n - 1] = 0.0
    sweeps = []
    fails = 0
    eps = np.finfo(np.float64).eps
    for l in range(n):
        it = 0
        while True:
            m = l
            while m < n - 1:
                dd = abs(d[m]) + abs(d[m + 1])
                if abs(e[m]) <= eps * dd:
                    break
                m += 1
            if m == l:
                break
            if it == max_iter:
                fails += 1
                break
            it += 1
            g = (d[l + 1] - d[l]) / (2.0 * e[l])
            r = np.hypot(g, 1.0)
            g = d[m] - d[l] + e[l] / (g + np.copysign(r, g))
            s = c = 1.0
            p = 0.0
            rots = []
            i = m - 1
            underflow = False
            while i >= l:
                f = s * e[i]
                b = c * e[i]
                r = np.hypot(f, g)
                e[i + 1] = r
                if r == 0.0:
                    d[i + 1] -= p
                    e[m] = 0.0
                    underflow = True
                    break
                s = f / r
                c = g / r
                g = d[i + 1] - p
                r = (d[i] - g) * s + 2.0 * c * b
                p = s * r
                d[i + 1] = g + p
                g = c * r - b
                rots.append((c, s))
                i -= 1
            sweeps.append((l, m, rots))
            if underflow:
                continue
            d[l] -= p
            e[l] = g
            e[m] = 0.0
    return d, sweeps, fails


def replay(n: int, sweeps) -> np.ndarray:
    z = np.eye(n)  # z[row, col]; columns converge to eigenvectors of T
    for (l, m, rots) in sweeps:
        carry = z[:, m].copy()
        i = m - 1
        for (c, s) in rots:
            zi = z[:, i].copy()
            z[:, i + 1] = s * zi + c * carry
            carry = c * zi - s * carry
            i -= 1
        z[:, i + 1] = carry
    return z


def back_transform(z: np.ndarray, tau: np.ndarray, y: np.ndarray) -> np.ndarray:
    n = z.shape[0]
    u = z.astype(np.complex128)
    for i in range(n - 2, -1, -1):
        v = y[i]
        s = np.conj(v) @ u
        u -= tau[i] * np.outer(v, s)
    return u


def wy_tfactor(v: np.ndarray, tau: np.ndarray) -> np.ndarray:
    """T of H_0 ... H_{k-1} = I - V T V^H (LAPACK zlarft, forward / columnwise), as csrc/eigh_wy.cuh builds it."""
    k = v.shape[1]
    t = np.zeros((k, k), dtype=np.complex128)
    g = v.conj().T @ v
    for j in range(k):
        t[j, j] = tau[j]
        if j > 0:
            t[:j, j] = -tau[j] * (t[:j, :j] @ g[:j, j])
    return t


def back_transform_wy(z: np.ndarray, tau: np.ndarray, y: np.ndarray, nb: int = 32) -> np.ndarray:
    """Blocked back-transformation in the row layout of csrc/eigh_wy.cuh: W = Z conj(V), W2 = W T^T, Z -= W2 V^T."""
    n = z.shape[0]
    zr = z.T.astype(np.complex128)  # zr[e, k]
    for i0 in reversed(range(0, n - 1, nb)):
        i1 = min(i0 + nb, n - 1)
        v = np.array([y[i] for i in range(i0, i1)]).T
        t = wy_tfactor(v, tau[i0:i1])
        w2 = (zr @ v.conj()) @ t.T
        zr = zr - w2 @ v.T
    return zr.T


def eigh_model(h: np.ndarray):
    n = h.shape[0]
    d, e, tau, y = tridiagonalise(h)
    lam, sweeps, fails = ql_record(d, e)
    z = replay(n, sweeps)
    order = np.argsort(lam, kind="stable")
    u = back_transform(z[:, order], tau, y)
    transform = np.conj(u).T  # rows = eigenvectors of H
    nrot = sum(len(r) for (_, _, r) in sweeps)
    return lam[order], transform, fails, nrot, len(sweeps)
