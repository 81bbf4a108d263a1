"""BlochTransposedBasis (reference: slate_quantum/bloch/_transposed_basis.py:23-263), CUDA-backed.

inner order  [(inner_index, block_index), (inner_index, block_index), ...]
this order   [(block_index, block_index, ...), (inner_index, inner_index, ...)]

When the inner basis is the standard tuple of BlochShiftedBasis(TransformedBasis(...)) children (what
`bloch.build.basis_from_metadata` constructs) conversions to the supercell momentum / position basis are
fused: ONE K5 permutation kernel (+ ONE multi-axis K4 FFT) instead of 2*nd+1 numpy copies.
"""

from __future__ import annotations

import itertools

import torch

from slate_quantum_b200 import kernels
from slate_quantum_b200.bloch._shifted_basis import BlochShiftedBasis
from slate_quantum_b200.core import basis as _basis
from slate_quantum_b200.core.basis import Basis, FundamentalBasis, TransformedBasis, TupleBasis, WrappedBasis, _moved, _strip


class BlochTransposedBasis(WrappedBasis):
    def __init__(self, inner: TupleBasis) -> None:
        super().__init__(inner)

    def metadata(self):
        return self.inner.metadata()

    @property
    def size(self) -> int:
        return self.inner.size

    def dual_basis(self) -> "BlochTransposedBasis":
        return BlochTransposedBasis(self.inner.dual_basis())

    def _dims(self) -> tuple[tuple[int, ...], tuple[int, ...]]:
        children = self.metadata().children
        return tuple(c.inner.fundamental_size for c in children), tuple(c.n_repeats for c in children)

    # ---- generic transposes (reference :102-133 / :136-167) ---------------------------------
    def into_inner(self, t: torch.Tensor, axis: int = -1) -> torch.Tensor:
        n_states, n_repeats = self._dims()
        nd = len(n_states)

        def fn(v: torch.Tensor) -> torch.Tensor:
            lead = v.ndim - 1
            stacked = v.reshape(*v.shape[:-1], *n_repeats, *n_states)
            order = list(range(lead)) + list(itertools.chain(*((lead + i + nd, lead + i) for i in range(nd))))
            return stacked.permute(order).reshape(v.shape)

        return _moved(t, axis, fn)

    def from_inner(self, t: torch.Tensor, axis: int = -1) -> torch.Tensor:
        n_states, n_repeats = self._dims()
        nd = len(n_states)

        def fn(v: torch.Tensor) -> torch.Tensor:
            lead = v.ndim - 1
            stacked = v.reshape(*v.shape[:-1], *itertools.chain(*zip(n_states, n_repeats, strict=True)))
            order = (
                list(range(lead))
                + list(range(lead + 1, lead + 2 * nd, 2))
                + list(range(lead, lead + 2 * nd, 2))
            )
            return stacked.permute(order).reshape(v.shape)

        return _moved(t, axis, fn)

    # ---- fused fast paths --------------------------------------------------------------------
    def _standard_children(self) -> bool:
        for c in _basis.as_tuple(self.inner).children:
            s = _strip(c)
            if not isinstance(s, BlochShiftedBasis):
                return False
            tr = _strip(s.inner)
            if not (isinstance(tr, TransformedBasis) and isinstance(_strip(tr.inner), FundamentalBasis)):
                return False
        return True

    def _is_dual_any(self) -> bool:
        return any(bool(_strip(_strip(_strip(c).inner).inner).is_dual) for c in _basis.as_tuple(self.inner).children)

    def supercell_momentum_basis(self) -> TupleBasis:
        """The TupleBasis(TransformedBasis...) this Bloch basis is a re-ordering of."""
        tup = _basis.as_tuple(self.inner)
        return TupleBasis(tuple(_strip(_strip(c).inner) for c in tup.children), tup.extra)

    def to_supercell_momentum(self, t: torch.Tensor, axis: int = -1) -> torch.Tensor:
        n_states, n_repeats = self._dims()
        return _moved(t, axis, lambda v: kernels.bloch_permute(n_states, n_repeats, v, 1))

    def from_supercell_momentum(self, t: torch.Tensor, axis: int = -1) -> torch.Tensor:
        n_states, n_repeats = self._dims()
        return _moved(t, axis, lambda v: kernels.bloch_permute(n_states, n_repeats, v, 0))

    def _cell_path(self) -> bool:
        """The factorised route (in-cell FFT + twiddle + R-point FFTs across blocks) needs
        every repeat to be a smooth length the shared-memory FFT takes in one pass."""
        _, n_repeats = self._dims()
        return all(r <= 4096 and kernels.is_smooth(r) for r in n_repeats)

    def into_fundamental(self, t: torch.Tensor, axis: int = -1) -> torch.Tensor:
        """Bloch order -> supercell position basis (reference chain :102-133 + _shifted_basis.py:62-84 + the
        TransformedBasis leaves): N-point FFTs inside the blocks, the Bloch twiddle, R-point FFTs across blocks -
        no K5 gather and no supercell-sized FFT lines.  Bra (dual) indices use the conjugate transform."""
        if not self._standard_children():
            return super().into_fundamental(t, axis)
        n_states, n_repeats = self._dims()
        big = tuple(n * r for n, r in zip(n_states, n_repeats, strict=True))
        dual = self._is_dual_any()

        def fn(v: torch.Tensor) -> torch.Tensor:
            if not self._cell_path():
                vk = kernels.bloch_permute(n_states, n_repeats, v, 1)
                return kernels.fft_c2c(vk, big, -1 if dual else +1, out=vk)
            x = torch.conj_physical(v) if dual else v
            cell = kernels.bloch_cell_vectors(n_states, n_repeats, x, 1)
            pos = kernels.cell_fft(n_states, n_repeats, cell, 1, out=cell if len(n_states) == 1 else None)
            return torch.conj_physical(pos).reshape(v.shape) if dual else pos.reshape(v.shape)

        return _moved(t, axis, fn)

    def from_fundamental(self, t: torch.Tensor, axis: int = -1) -> torch.Tensor:
        if not self._standard_children():
            return super().from_fundamental(t, axis)
        n_states, n_repeats = self._dims()
        big = tuple(n * r for n, r in zip(n_states, n_repeats, strict=True))
        dual = self._is_dual_any()

        def fn(v: torch.Tensor) -> torch.Tensor:
            if not self._cell_path():
                vk = kernels.fft_c2c(v, big, +1 if dual else -1)
                return kernels.bloch_permute(n_states, n_repeats, vk, 0)
            x = torch.conj_physical(v) if dual else v
            cell = kernels.cell_fft(n_states, n_repeats, x, 0)
            out = kernels.bloch_cell_vectors(n_states, n_repeats, cell, 0, out=cell)
            return torch.conj_physical(out).reshape(v.shape) if dual else out.reshape(v.shape)

        return _moved(t, axis, fn)

    def convert_vector_into(self, t: torch.Tensor, other: Basis, axis: int = -1) -> torch.Tensor:
        if self == other:
            return t
        if self._standard_children() and _strip(other) == self.supercell_momentum_basis():
            return self.to_supercell_momentum(t, axis)
        return super().convert_vector_into(t, other, axis)

    def _key(self) -> tuple:
        return ("BlochTransposed", self.inner._key())
