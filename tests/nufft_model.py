"""Numpy model of the uniform-grid NUFFT evolve kernel (test helper, runs on CPU).

Mirrors slate_quantum_b200/csrc/evolve_nufft.cuh step by step.  For a uniform time grid t_r = t_0 + r dt the
eigenbasis evolution of a k-block (dynamics/schrodinger.py:45-56 of the reference followed by the
StateList.with_state_basis back-transformation, state/_state.py:153-165)

    out[r, j] = sum_e c_e exp(-1j w_e t_r / hbar) V[e, j]

is, per tile of TT = 512 rows, a type-1 non-uniform DFT with N sources at the angles theta_e = (w_e dt / hbar) mod 2 pi
shared by every column j of the block:

    out[r0 + m, j] = sum_e a_e[j] exp(-1j (m - TT/2) theta_e),   a_e[j] = c_e exp(-1j w_e t_{r0 + TT/2} / hbar) V[e, j]

(interleaved tiles: rows r0 + K m with the angles K theta_e, see `interleave`)

evaluated as (1) a gather of the sources onto an oversampled grid of NG = 2 TT points with the "exponential of
semicircle" window of width W (Barnett, Magland, af Klinteberg 2019), (2) an in-place decimation-in-frequency FFT of
NG points (radix 16, 8, 8) and (3) a division by the window's Fourier transform on the TT wanted frequencies.
"""
from __future__ import annotations

import numpy as np

TT = 512
NG = 1024
PAD = 3  # zero weights either side of a source's window: a thread gathers 4 consecutive grid points


def es_window(z: np.ndarray, beta: float) -> np.ndarray:
    return np.exp(beta * (np.sqrt(np.maximum(0.0, 1.0 - z * z)) - 1.0))


def reduce_angle(x: np.ndarray) -> np.ndarray:
    """x mod 2 pi in [0, 2 pi) with a two-term (Cody-Waite) reduction, as the plan kernel does."""
    two_pi_hi = 6.283185307179586
    two_pi_lo = 2.4492935982947064e-16
    q = np.rint(x / two_pi_hi)
    th = (x - q * two_pi_hi) - q * two_pi_lo
    return np.where(th < 0.0, th + two_pi_hi, th)


def interleave(w_all: np.ndarray, dt: float, hbar: float, tiles: int) -> int:
    """K of the kernel's nu_interleave: a tile takes every K-th row of a super-tile of K * TT rows, so that its angles
    K w dt / hbar spread over the circle even when the time step is small against the bandwidth of the spectrum."""
    x = w_all * dt / hbar
    span = float(x.max() - x.min())
    k_max = min(tiles & (-tiles), 64)
    k = 1
    while 2 * k <= k_max and 2.0 * k * span < 0.8 * 2 * np.pi:
        k *= 2
    return k


def plan(w: np.ndarray, dt: float, hbar: float, width: int, beta: float, k: int = 1):
    """Per block: sources sorted by the first grid point of their window.  Returns (src, l0s, wts, pc)."""
    n = len(w)
    u = reduce_angle(k * (w * dt / hbar)) * (NG / (2 * np.pi))
    l0 = np.ceil(u - width / 2).astype(np.int64)
    l0w = np.mod(l0, NG)
    order = np.lexsort((np.arange(n), l0w))
    src = order
    l0s = l0w[order]
    wts = np.zeros((n, width + 2 * PAD))
    for s, e in enumerate(order):
        kk = np.arange(width)
        wts[s, PAD : PAD + width] = es_window((l0[e] + kk - u[e]) / (width / 2), beta)
    # pc[x + 1] = number of sources with l0s <= x for x in [-1, NG + width + 3): pc[0] = 0
    xs = np.arange(-1, NG + width + 3)
    pc = np.searchsorted(l0s, xs, side="right")
    return src, l0s, wts, pc


def deconvolution(width: int, beta: float, points: int = 256) -> np.ndarray:
    """1 / (window transform) on the frequencies m' = m - TT/2, m = 0..TT-1, by the trapezoid rule in z = sin(t)."""
    t = -np.pi / 2 + (np.arange(points) + 0.5) * (np.pi / points)
    mp = np.arange(TT) - TT // 2
    alpha = mp * (2 * np.pi / NG) * (width / 2)
    f = np.exp(beta * (np.cos(t) - 1.0)) * np.cos(t)
    integral = (np.cos(np.outer(alpha, np.sin(t))) * f).sum(axis=1) * (np.pi / points)
    return 1.0 / ((width / 2) * integral)


def gather(src, l0s, wts, pc, a: np.ndarray, width: int) -> np.ndarray:
    """g[l, j] = sum over the sources whose window covers l; a thread owns the 4 grid points of a quad."""
    n, nj = a.shape
    g = np.zeros((NG, nj), dtype=np.complex128)

    def pcx(x):  # sources with l0s <= x
        return 0 if x < -1 else int(pc[min(x, NG + width + 2) + 1])

    for quad in range(NG // 4):
        l = 4 * quad
        acc = np.zeros((4, nj), dtype=np.complex128)
        for base in (l, l + NG) if l < width + 3 else (l,):
            for s in range(pcx(base - width), pcx(base + 3)):
                k0 = base - int(l0s[s])
                for i in range(4):
                    acc[i] += wts[s, PAD + k0 + i] * a[src[s]]
        g[l : l + 4] = acc
    return g


def fft_dif(g: np.ndarray) -> np.ndarray:
    """In-place DIF stages of the kernel; returns the TT wanted rows in output order m = 0..TT-1."""
    x = g.copy()
    nj = x.shape[1]
    tw = np.exp(-2j * np.pi * np.arange(NG) / NG)
    # stage 1: radix 16 over n1 (stride 64), twiddle w_1024^(n2 k1)
    for n2 in range(64):
        idx = 64 * np.arange(16) + n2
        y = np.fft.fft(x[idx], axis=0)
        x[idx] = y * tw[(n2 * np.arange(16)) % NG][:, None]
    # stage 2: radix 8 over n2a (stride 8) inside every block of 64, twiddle w_64^(n2b k2)
    for k1 in range(16):
        for n2b in range(8):
            idx = 64 * k1 + 8 * np.arange(8) + n2b
            y = np.fft.fft(x[idx], axis=0)
            x[idx] = y * tw[(16 * n2b * np.arange(8)) % NG][:, None]
    # stage 3: radix 8 over n2b (stride 1); frequency k = k1 + 16 k2 + 128 k3, only k3 in {0, 1, 6, 7} is kept
    out = np.zeros((TT, nj), dtype=np.complex128)
    for k1 in range(16):
        for k2 in range(8):
            idx = 64 * k1 + 8 * k2 + np.arange(8)
            z = np.fft.fft(x[idx], axis=0)
            for k3 in (0, 1):
                out[k1 + 16 * k2 + 128 * k3 + TT // 2] = z[k3]
            for k3 in (6, 7):
                out[k1 + 16 * k2 + 128 * (k3 - 6)] = z[k3]
    return out


def evolve_uniform(w, c, v, t0: float, dt: float, t_count: int, hbar: float, width: int = 14, beta_per_w: float = 2.3,
                   w_all=None):
    """out[r, j] for r < t_count from eigenvalues w [N], coefficients c [N], eigenvector rows v [N, NJ]; `w_all` =
    the eigenvalues of every block of the call (they fix the interleave factor; default: this block's)."""
    beta = beta_per_w * width
    tiles = -(-t_count // TT)
    k = interleave(w if w_all is None else w_all, dt, hbar, tiles)
    src, l0s, wts, pc = plan(w, dt, hbar, width, beta, k)
    dec = deconvolution(width, beta)
    out = np.zeros((t_count, v.shape[1]), dtype=np.complex128)
    for ti in range(tiles):
        r0 = (ti // k) * k * TT + ti % k  # rows r0 + k m
        tc = (t0 + r0 * dt) + (k * (TT // 2)) * dt
        a = (c * np.exp(-1j * w * tc / hbar))[:, None] * v
        f = fft_dif(gather(src, l0s, wts, pc, a, width)) * dec[:, None]
        rows = np.arange(r0, t_count, k)[:TT]
        out[rows] = f[: len(rows)]
    return out
