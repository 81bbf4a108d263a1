"""numpy model of the index maps of csrc/fft.cu's Stockham pass (test infrastructure; CPU only).

`stockham(x, radices)` walks the stages exactly like `stage_fixed` (out of place, ping-pong buffers);
`stockham_inplace_mid(x, radices)` runs every middle stage like `stage_fixed_inplace`: all butterflies READ their
inputs, then all butterflies WRITE their outputs into the same buffer.  Both must equal np.fft.
"""

from __future__ import annotations

import numpy as np


def factorize(length: int, allow16: bool = False) -> list[int]:
    """launch_pass / factorize: 8s, then 4s, then 2s, then odd primes; with `allow16`, power-of-two lengths of 128
    points and more lead with the fewest radix-16 stages that reach the minimum stage count."""
    out, rem = [], length
    if allow16 and length >= 128 and length & (length - 1) == 0:
        k = length.bit_length() - 1
        best16, best = 0, (k + 2) // 3
        for n16 in range(1, k // 4 + 1):
            st = n16 + (k - 4 * n16 + 2) // 3
            if st < best:
                best, best16 = st, n16
        out += [16] * best16
        rem //= 16**best16
    for r in (8, 4, 2):
        while rem % r == 0:
            out.append(r)
            rem //= r
    p = 3
    while rem > 1:
        while rem % p == 0:
            out.append(p)
            rem //= p
        p += 2
    return out


def _stage(src: np.ndarray, radix: int, ns: int, sign: int) -> np.ndarray:
    """One stage: butterfly j reads src[j + t * per_line], writes dst[(j - k) * R + k + t * ns], k = j % ns."""
    n = src.size
    per_line = n // radix
    dst = np.empty_like(src)
    t = np.arange(radix)
    dft = np.exp(sign * 2j * np.pi * np.outer(t, t) / radix)
    for j in range(per_line):
        k = j % ns
        x = src[j + t * per_line] * np.exp(sign * 2j * np.pi * t * k / (ns * radix))
        dst[(j - k) * radix + k + t * ns] = dft @ x
    return dst


def stockham(x: np.ndarray, radices: list[int], sign: int = -1) -> np.ndarray:
    buf, ns = np.asarray(x, dtype=np.complex128).copy(), 1
    for r in radices:
        buf = _stage(buf, r, ns, sign)
        ns *= r
    return buf


def stockham_inplace_mid(x: np.ndarray, radices: list[int], sign: int = -1) -> np.ndarray:
    """First and last stage out of place (global <-> shared), the middle ones in place: phase 1 gathers every
    butterfly's inputs and computes its outputs, phase 2 (after the barrier) scatters them into the SAME buffer."""
    buf, ns = np.asarray(x, dtype=np.complex128).copy(), 1
    for s, r in enumerate(radices):
        if s in (0, len(radices) - 1):
            buf = _stage(buf, r, ns, sign)
        else:
            n = buf.size
            per_line = n // r
            t = np.arange(r)
            dft = np.exp(sign * 2j * np.pi * np.outer(t, t) / r)
            held = []
            for j in range(per_line):  # phase 1: reads only
                k = j % ns
                held.append(dft @ (buf[j + t * per_line] * np.exp(sign * 2j * np.pi * t * k / (ns * r))))
            for j in range(per_line):  # phase 2: writes only
                k = j % ns
                buf[(j - k) * r + k + t * ns] = held[j]
        ns *= r
    return buf
