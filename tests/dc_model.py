"""Numpy model of the GPU divide-and-conquer eigensolver for the real symmetric tridiagonal stage of K2.

Test helper (runs on CPU).  Mirrors slate_quantum_b200/csrc/eigh_dc.cu step by step:

  * the matrix is torn at every node boundary of a balanced binary tree (all leaves at the same depth, leaf
    size <= 32): T = diag(T1', T2') + |beta| w w^T (Cuppen), the tears of every level applied up front;
  * leaves are solved by implicit QL with the rotations applied to the leaf's eigenvector rows;
  * a merge sorts the children's eigenvalues, deflates (negligible z components and close pole pairs by a
    Givens rotation of two eigenvector rows - the criteria of LAPACK dlaed2), solves the secular equation
    1/rho + sum z_i^2 / (d_i - lambda) = 0 root by root in the shifted variable tau = lambda - d_origin
    ("middle way" two-pole rational iteration, bracket-safeguarded), recomputes z from the roots
    (Gu / Eisenstat, so that the eigenvectors are orthogonal to working precision), and forms the merged
    eigenvectors as ONE dense product  Qe_new = U^T . Qe  (rows = eigenvectors) - the FP64 tensor-core GEMM
    of the CUDA version.

Reference call site this whole stage replaces: np.linalg.eigh inside
slate_core.linalg.into_diagonal_hermitian (slate_quantum/operator/_linalg.py:56).
"""
from __future__ import annotations

import numpy as np

EPS = 2.220446049250313e-16
LEAF = 32
MAX_SECULAR_ITER = 40


def tree_depth(n: int, leaf: int = LEAF) -> int:
    depth = 0
    while ((n + (1 << depth) - 1) >> depth) > leaf:
        depth += 1
    return depth


def node_start(n: int, level: int, i: int) -> int:
    return (i * n) >> level


def leaf_ql(d: np.ndarray, e: np.ndarray, max_iter: int = 60):
    """Implicit QL with vectors on one leaf; returns (eigenvalues ascending, rows = eigenvectors, fails)."""
    n = len(d)
    d = d.copy()
    e = np.concatenate([e[: n - 1], [0.0]]).astype(float)
    z = np.eye(n)  # z[row, col]; rotations act on columns
    fails = 0
    for l in range(n):
        it = 0
        while True:
            m = n - 1
            for idx in range(l, n - 1):
                if abs(e[idx]) <= EPS * (abs(d[idx]) + abs(d[idx + 1])):
                    m = idx
                    break
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
            underflow = False
            for i in range(m - 1, l - 1, -1):
                f = s * e[i]
                b = c * e[i]
                r = np.sqrt(f * f + g * g)
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
                zi = z[:, i].copy()
                z[:, i] = c * zi - s * z[:, i + 1]
                z[:, i + 1] = s * zi + c * z[:, i + 1]
            if not underflow:
                d[l] -= p
                e[l] = g
                e[m] = 0.0
    order = np.argsort(d, kind="stable")
    return d[order], z.T[order].copy(), fails


def _iterate(j_lo, j_hi, split, origin, delta, z2, rhoinv, tau, lo, hi, last):
    """Two-pole rational ("middle way") iteration on g(tau) = rhoinv + sum z2 / (delta - tau), bracketed by (lo, hi).

    Poles j_lo / j_hi are the two the interpolant keeps explicit; terms [0, split) form psi, [split, k) phi.
    """
    for it in range(MAX_SECULAR_ITER):
        dm = delta - tau
        t = z2 / dm
        psi = np.sum(t[:split])
        phi = np.sum(t[split:])
        t2 = t / dm
        dpsi = np.sum(t2[:split])
        dphi = np.sum(t2[split:])
        w = rhoinv + psi + phi
        erretm = 8.0 * (abs(phi) + abs(psi)) + 2.0 * rhoinv + abs(tau) * (dpsi + dphi)
        if abs(w) <= EPS * erretm:
            return origin, tau, it
        if w > 0.0:
            hi = tau
        else:
            lo = tau
        dl, du = dm[j_lo], dm[j_hi]
        c = w - dl * dpsi - du * dphi
        a = (dl + du) * w - dl * du * (dpsi + dphi)
        b = dl * du * w
        if last:
            if c < 0.0:
                c = abs(c)
            if c == 0.0:
                eta = hi - tau
            else:
                disc = np.sqrt(abs(a * a - 4.0 * b * c))
                eta = (a + disc) / (2.0 * c) if a >= 0.0 else 2.0 * b / (a - disc)
        else:
            if c == 0.0:
                eta = b / a if a != 0.0 else 0.0
            else:
                disc = np.sqrt(abs(a * a - 4.0 * b * c))
                eta = (a - disc) / (2.0 * c) if a <= 0.0 else 2.0 * b / (a + disc)
        if w * eta >= 0.0:
            eta = -w / (dpsi + dphi)
        new = tau + eta
        if not (lo < new < hi):
            new = 0.5 * (lo + hi)
        if new == tau or new == lo or new == hi:
            return origin, tau, it
        tau = new
    return origin, tau, MAX_SECULAR_ITER


def secular_root(j: int, dk: np.ndarray, z2: np.ndarray, rho: float):
    """Root j of 1/rho + sum z2_i / (dk_i - lambda); returns (origin, tau, iterations), lambda = dk[origin] + tau.

    The initial guess keeps the two poles next to the root exact and freezes the rest at the interval's
    midpoint (the dlaed4 start), which is what makes roots that hug a pole (tiny z_j) converge at once.
    """
    k = len(dk)
    rhoinv = 1.0 / rho
    if k == 1:
        return 0, rho * z2[0], 0
    if j < k - 1:
        gap = dk[j + 1] - dk[j]
        dmid = (dk - dk[j]) - 0.5 * gap
        c = rhoinv + np.sum(z2[:j] / dmid[:j]) + np.sum(z2[j + 2 :] / dmid[j + 2 :])
        w = c + z2[j] / dmid[j] + z2[j + 1] / dmid[j + 1]
        if w > 0.0:  # root in the left half: shift to pole j
            origin, lo, hi = j, 0.0, 0.5 * gap
            a = c * gap + z2[j] + z2[j + 1]
            b = z2[j] * gap
            den = a + np.sqrt(abs(a * a - 4.0 * b * c))
            tau = 2.0 * b / den if den > 0.0 else hi
        else:
            origin, lo, hi = j + 1, -0.5 * gap, 0.0
            a = -c * gap + z2[j] + z2[j + 1]
            b = z2[j + 1] * gap
            den = a + np.sqrt(abs(a * a + 4.0 * b * c))
            tau = -2.0 * b / den if den > 0.0 else lo
        if not (lo < tau < hi):
            tau = 0.5 * (lo + hi)
        return _iterate(j, j + 1, j + 1, origin, dk - dk[origin], z2, rhoinv, tau, lo, hi, False)
    # last root: to the right of every pole, lambda in (dk[k-1], dk[k-1] + rho * sum z2]
    origin = k - 1
    delta = dk - dk[k - 1]
    lo = 0.0
    hi = rho * float(np.sum(z2)) * (1.0 + 8.0 * EPS) + np.finfo(float).tiny
    mid = 0.5 * hi
    dmid = delta - mid
    c = rhoinv + np.sum(z2[: k - 2] / dmid[: k - 2])
    w = c + z2[k - 2] / dmid[k - 2] + z2[k - 1] / dmid[k - 1]
    gap = dk[k - 1] - dk[k - 2]
    a = -c * gap + z2[k - 2] + z2[k - 1]
    b = z2[k - 1] * gap
    if c > 0.0:
        disc = np.sqrt(abs(a * a + 4.0 * b * c))
        tau = 2.0 * b / (disc - a) if a < 0.0 else (a + disc) / (2.0 * c)
    else:
        tau = hi
    if w > 0.0:
        hi = mid
    else:
        lo = mid
    if not (lo < tau < hi):
        tau = 0.5 * (lo + hi)
    return _iterate(k - 2, k - 1, k - 1, origin, delta, z2, rhoinv, tau, lo, hi, True)


def merge(dl: np.ndarray, ql: np.ndarray, dr: np.ndarray, qr: np.ndarray, beta: float, stats: dict | None = None):
    """Merge two solved halves (rows of ql / qr = eigenvectors) coupled by the off-diagonal beta."""
    n1, n2 = len(dl), len(dr)
    n = n1 + n2
    qe = np.zeros((n, n))
    qe[:n1, :n1] = ql
    qe[n1:, n1:] = qr
    d = np.concatenate([dl, dr]).astype(float)
    sgn = 1.0 if beta >= 0 else -1.0
    z = np.concatenate([ql[:, n1 - 1], sgn * qr[:, 0]]) / np.sqrt(2.0)
    rho = 2.0 * abs(beta)
    order = np.argsort(d, kind="stable")
    tol = 8.0 * EPS * max(np.abs(d).max(), np.abs(z).max())
    nd_idx: list[int] = []
    defl: list[int] = []
    pj = -1
    for j in order:
        if rho * abs(z[j]) <= tol:
            defl.append(j)
            continue
        if pj < 0:
            pj = j
            continue
        s, c = z[pj], z[j]
        tau = np.hypot(c, s)
        t = d[j] - d[pj]
        c /= tau
        s = -s / tau
        if abs(t * c * s) <= tol:
            z[j] = tau
            z[pj] = 0.0
            rp, rj = qe[pj].copy(), qe[j].copy()
            qe[pj] = c * rp + s * rj
            qe[j] = -s * rp + c * rj
            t2 = d[pj] * c * c + d[j] * s * s
            d[j] = d[pj] * s * s + d[j] * c * c
            d[pj] = t2
            defl.append(pj)
            pj = j
        else:
            nd_idx.append(pj)
            pj = j
    if pj >= 0:
        nd_idx.append(pj)
    k = len(nd_idx)
    if stats is not None:
        stats.setdefault("k", []).append((n, k))
    lam_all = np.empty(n)
    q_all = np.empty((n, n))
    if k > 0:
        dk = d[nd_idx]
        zk = z[nd_idx]
        z2 = zk * zk
        origin = np.zeros(k, dtype=int)
        mu = np.zeros(k)
        for j in range(k):
            origin[j], mu[j], its = secular_root(j, dk, z2, rho)
            if stats is not None:
                stats.setdefault("iters", []).append(its)
        # dmat[i, j] = dk_i - lambda_j, accurately
        dmat = (dk[:, None] - dk[origin][None, :]) - mu[None, :]
        # Gu / Eisenstat: zhat_i^2 ~ prod_j (dk_i - lam_j) / prod_{j != i} (dk_i - dk_j)
        zhat = np.empty(k)
        for i in range(k):
            w = dmat[i, i]
            for j in range(k):
                if j != i:
                    w *= dmat[i, j] / (dk[i] - dk[j])
            zhat[i] = np.copysign(np.sqrt(abs(w)), zk[i])
        u = zhat[:, None] / dmat
        u /= np.linalg.norm(u, axis=0)[None, :]
        lam_all[:k] = dk[origin] + mu
        q_all[:k] = u.T @ qe[nd_idx]
    lam_all[k:] = d[defl]
    q_all[k:] = qe[defl]
    order2 = np.argsort(lam_all, kind="stable")
    return lam_all[order2], q_all[order2]


def dc_eigh_tridiagonal(d: np.ndarray, e: np.ndarray, leaf: int = LEAF, stats: dict | None = None):
    """Eigen-decomposition of the symmetric tridiagonal (d, e[:n-1]); returns (w ascending, rows = eigenvectors)."""
    n = len(d)
    depth = tree_depth(n, leaf)
    d = np.array(d, dtype=float)
    e = np.array(e, dtype=float)
    # scale to unit max-norm (dstedc does the same): the deflation tolerance compares |d| with |z| <= 1
    scale = max(np.abs(d).max(), np.abs(e[: n - 1]).max() if n > 1 else 0.0)
    if scale > 0.0:
        d /= scale
        e /= scale
    else:
        scale = 1.0
    for level in range(1, depth + 1):
        for i in range(1, 1 << level, 2):  # odd nodes start at a NEW tear of this level
            s = node_start(n, level, i)
            d[s - 1] -= abs(e[s - 1])
            d[s] -= abs(e[s - 1])
    nodes = []
    for i in range(1 << depth):
        s0, s1 = node_start(n, depth, i), node_start(n, depth, i + 1)
        w, q, _ = leaf_ql(d[s0:s1], e[s0 : s1 - 1] if s1 - s0 > 1 else np.zeros(0))
        nodes.append((w, q))
    for level in range(depth - 1, -1, -1):
        merged = []
        for i in range(1 << level):
            s_mid = node_start(n, level + 1, 2 * i + 1)
            (wl, ql), (wr, qr) = nodes[2 * i], nodes[2 * i + 1]
            merged.append(merge(wl, ql, wr, qr, e[s_mid - 1], stats))
        nodes = merged
    w, q = nodes[0]
    return w * scale, q
