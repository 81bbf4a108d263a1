"""BlochShiftedBasis (reference: slate_quantum/bloch/_shifted_basis.py:13-180), CUDA-backed.

Re-orders one supercell-momentum axis (FFT order, k_tot = k_inner*R + k_bloch) into (inner_index,
block_index) order.  The arithmetic-free permutation is the 1-axis case of the K5 kernel.
"""

from __future__ import annotations

import torch

from slate_quantum_b200 import kernels
from slate_quantum_b200.core.basis import Basis, WrappedBasis, _moved


class BlochShiftedBasis(WrappedBasis):
    def __init__(self, inner: Basis) -> None:
        super().__init__(inner)

    @property
    def size(self) -> int:
        return self.inner.size

    def dual_basis(self) -> "BlochShiftedBasis":
        return BlochShiftedBasis(self.inner.dual_basis())

    def _dims(self) -> tuple[int, int]:
        meta = self.metadata()
        return meta.inner.fundamental_size, meta.n_repeats

    def into_inner(self, t: torch.Tensor, axis: int = -1) -> torch.Tensor:
        """(inner_index, block_index) order -> supercell FFT order (reference :62-84)."""
        n_states, n_repeats = self._dims()

        def fn(v: torch.Tensor) -> torch.Tensor:
            blk_major = v.reshape(*v.shape[:-1], n_states, n_repeats).transpose(-1, -2).contiguous()
            return kernels.bloch_permute((n_states,), (n_repeats,), blk_major.reshape(v.shape), 1)

        return _moved(t, axis, fn)

    def from_inner(self, t: torch.Tensor, axis: int = -1) -> torch.Tensor:
        """supercell FFT order -> (inner_index, block_index) order (reference :87-111)."""
        n_states, n_repeats = self._dims()

        def fn(v: torch.Tensor) -> torch.Tensor:
            blk_major = kernels.bloch_permute((n_states,), (n_repeats,), v, 0)
            return blk_major.reshape(*v.shape[:-1], n_repeats, n_states).transpose(-1, -2).reshape(v.shape)

        return _moved(t, axis, fn)

    def _key(self) -> tuple:
        return ("BlochShifted", self.inner._key())
