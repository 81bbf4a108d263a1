"""CPU: the Stockham index maps of csrc/fft.cu (out-of-place stages and the in-place middle stage of the
one-buffer variant) reproduce np.fft; and the launch rule picks the in-place variant exactly where intended."""

from __future__ import annotations

import numpy as np
import pytest

from fft_model import factorize, stockham, stockham_inplace_mid


@pytest.mark.parametrize("length", [8, 64, 128, 256, 512, 1024, 96, 60, 343])
@pytest.mark.parametrize("sign", [-1, 1])
def test_stockham_stages_equal_numpy(length, sign):
    rng = np.random.default_rng(length)
    x = rng.normal(size=length) + 1j * rng.normal(size=length)
    want = np.fft.fft(x) if sign < 0 else np.fft.ifft(x) * length
    np.testing.assert_allclose(stockham(x, factorize(length), sign), want, atol=1e-11)


@pytest.mark.parametrize("length", [128, 256, 512, 1024, 2048, 4096])
@pytest.mark.parametrize("sign", [-1, 1])
def test_radix16_stage_plan_equals_numpy(length, sign):
    """Radix-16 stages (dft_small<16>): 128 = 16*8 and 256 = 16*16 run in two stages, 1024..4096 in three; 512 keeps
    8*8*8 (no stage saved)."""
    radices = factorize(length, allow16=True)
    want_stages = {128: 2, 256: 2, 512: 3, 1024: 3, 2048: 3, 4096: 3}[length]
    assert len(radices) == want_stages and int(np.prod(radices)) == length
    assert (16 in radices) == (length != 512)
    rng = np.random.default_rng(length + 7)
    x = rng.normal(size=length) + 1j * rng.normal(size=length)
    want = np.fft.fft(x) if sign < 0 else np.fft.ifft(x) * length
    np.testing.assert_allclose(stockham(x, radices, sign), want, atol=1e-10)


def test_radix16_butterfly_as_4x4():
    """The register butterfly of csrc/fft.cu: t = 4 t1 + t2, k = k1 + 4 k2, twiddles W16^(t2 k1)."""
    rng = np.random.default_rng(16)
    x = rng.normal(size=16) + 1j * rng.normal(size=16)
    for sign in (-1, 1):
        f4 = np.exp(sign * 2j * np.pi * np.outer(np.arange(4), np.arange(4)) / 4)
        y = np.empty((4, 4), complex)  # y[t2][k1]
        for t2 in range(4):
            y[t2] = f4 @ x[t2::4]
        y *= np.exp(sign * 2j * np.pi * np.outer(np.arange(4), np.arange(4)) / 16)
        out = np.empty(16, complex)
        for k1 in range(4):
            out[k1::4] = f4 @ y[:, k1]
        want = np.fft.fft(x) if sign < 0 else np.fft.ifft(x) * 16
        np.testing.assert_allclose(out, want, atol=1e-12)


@pytest.mark.parametrize("length", [256, 512])
@pytest.mark.parametrize("sign", [-1, 1])
def test_inplace_middle_stage_equals_numpy(length, sign):
    """256 = 8*8*4 and 512 = 8*8*8: the three-stage lengths that take the one-buffer variant (launch_pass)."""
    radices = factorize(length)
    assert len(radices) == 3
    rng = np.random.default_rng(length + 1)
    x = rng.normal(size=length) + 1j * rng.normal(size=length)
    want = np.fft.fft(x) if sign < 0 else np.fft.ifft(x) * length
    np.testing.assert_allclose(stockham_inplace_mid(x, radices, sign), want, atol=1e-11)


def test_inplace_variant_thread_budget():
    """One butterfly per thread in the middle stage: lines * (len / R_mid) == 512 threads, buffer <= 100 KB."""
    for length, lines in ((256, 16), (512, 8)):
        r_mid = factorize(length)[1]
        assert lines * (length // r_mid) == 512
        assert (lines + 1) * length * 16 <= 100 * 1024
        assert lines * 16 >= 128  # bytes of a coalesced run across the CTA's adjacent lines
