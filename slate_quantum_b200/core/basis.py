"""Minimal slate_core.basis surface for the Bloch hot path (fresh implementation, CUDA-backed).

A *basis* describes how a vector of ``size`` numbers represents a vector of ``fundamental_size``
numbers on the "fundamental" (position-like) grid of its metadata.  The protocol mirrors slate_core's
(reference call sites: slate_quantum/bloch/_shifted_basis.py:62-111, _transposed_basis.py:102-167):

    basis.__into_fundamental__(vectors, axis).ok()     basis.__from_fundamental__(vectors, axis).ok()
    wrapped.__into_inner__(vectors, axis).ok()          wrapped.__from_inner__(vectors, axis).ok()
    basis.__convert_vector_into__(vectors, other, axis).ok()

All conversions run on the GPU: `vectors` may be numpy arrays (uploaded) or torch CUDA tensors, the
result is a torch CUDA complex128 tensor.  Arithmetic goes through the C ABI (FFT, Bloch permutation,
explicit-basis products); pure re-indexing (pad / crop / diagonal embedding / transposes) uses torch
tensor indexing as device-memory plumbing.  There is no CPU path.
"""

from __future__ import annotations

import itertools
import uuid
from typing import Any, Callable, Sequence

import numpy as np
import torch

from slate_quantum_b200 import kernels
from slate_quantum_b200.core.metadata import BasisMetadata, SimpleMetadata, TupleMetadata


class BasisConversion:
    """Lazy conversion thunk, evaluated by ``.ok()`` (slate_core.basis.BasisConversion)."""

    def __init__(self, fn: Callable[[], torch.Tensor]) -> None:
        self._fn = fn

    def ok(self) -> torch.Tensor:
        return self._fn()


def as_device(vectors: Any) -> torch.Tensor:
    """numpy / torch -> contiguous CUDA complex128 tensor."""
    if isinstance(vectors, torch.Tensor):
        return vectors.to(device="cuda", dtype=torch.complex128)
    return kernels.to_device(np.asarray(vectors, dtype=np.complex128), torch.complex128)


def _moved(t: torch.Tensor, axis: int, fn: Callable[[torch.Tensor], torch.Tensor]) -> torch.Tensor:
    """Apply `fn` (acting on the LAST axis of a contiguous tensor) along `axis`."""
    axis %= t.ndim
    t = t.resolve_conj().resolve_neg()  # the C ABI reads raw memory: lazy conj / neg view bits must be materialised
    if axis == t.ndim - 1:
        return fn(t.contiguous())
    return fn(t.movedim(axis, -1).contiguous()).movedim(-1, axis).contiguous()


# ======================================================================================== Basis
class Basis:
    def __init__(self, metadata: BasisMetadata, *, is_dual: bool = False) -> None:
        self._metadata = metadata
        self._is_dual = is_dual

    def metadata(self) -> BasisMetadata:
        return self._metadata

    @property
    def size(self) -> int:
        raise NotImplementedError

    @property
    def fundamental_size(self) -> int:
        return self.metadata().fundamental_size

    @property
    def is_dual(self) -> Any:
        return self._is_dual

    @property
    def ctype(self) -> Any:
        return _ANY_CTYPE

    def resolve_ctype(self) -> "Basis":
        return self

    def upcast(self) -> "AsUpcast":
        return AsUpcast(self, self.metadata())

    def dual_basis(self) -> "Basis":
        raise NotImplementedError

    @property
    def features(self) -> set[str]:
        return {"ADD", "MUL", "SUB", "LINEAR_MAP"}

    # ---- conversions (device tensors) -------------------------------------------------------
    def into_fundamental(self, t: torch.Tensor, axis: int = -1) -> torch.Tensor:
        raise NotImplementedError

    def from_fundamental(self, t: torch.Tensor, axis: int = -1) -> torch.Tensor:
        raise NotImplementedError

    def convert_vector_into(self, t: torch.Tensor, other: "Basis", axis: int = -1) -> torch.Tensor:
        if self == other:
            return t
        return other.from_fundamental(self.into_fundamental(t, axis), axis)

    # ---- the slate_core protocol names ------------------------------------------------------
    def __into_fundamental__(self, vectors: Any, axis: int = -1) -> BasisConversion:
        return BasisConversion(lambda: self.into_fundamental(as_device(vectors), axis))

    def __from_fundamental__(self, vectors: Any, axis: int = -1) -> BasisConversion:
        return BasisConversion(lambda: self.from_fundamental(as_device(vectors), axis))

    def __convert_vector_into__(self, vectors: Any, basis: "Basis", axis: int = -1) -> BasisConversion:
        return BasisConversion(lambda: self.convert_vector_into(as_device(vectors), basis, axis))

    def _key(self) -> tuple:
        raise NotImplementedError

    def __eq__(self, other: object) -> bool:
        return isinstance(other, Basis) and _strip(self)._key() == _strip(other)._key()

    def __hash__(self) -> int:
        return hash(_strip(self)._key())

    def __repr__(self) -> str:
        return f"{type(self).__name__}(size={self.size})"


class _AnyCtype:
    def assert_supports_dtype(self, _dtype: Any) -> None:
        return None


_ANY_CTYPE = _AnyCtype()
Ctype = _AnyCtype


def _strip(b: Basis) -> Basis:
    while isinstance(b, AsUpcast):
        b = b.inner
    return b


class FundamentalBasis(Basis):
    @staticmethod
    def from_size(size: int, *, is_dual: bool = False) -> "FundamentalBasis":
        return FundamentalBasis(SimpleMetadata(size), is_dual=is_dual)

    @property
    def size(self) -> int:
        return self.metadata().fundamental_size

    @property
    def points(self) -> np.ndarray:
        return np.arange(self.size)

    def dual_basis(self) -> "FundamentalBasis":
        return FundamentalBasis(self._metadata, is_dual=not self._is_dual)

    def into_fundamental(self, t: torch.Tensor, axis: int = -1) -> torch.Tensor:
        return t

    def from_fundamental(self, t: torch.Tensor, axis: int = -1) -> torch.Tensor:
        return t

    def _key(self) -> tuple:
        return ("F", self._metadata, self._is_dual)


class WrappedBasis(Basis):
    """A basis defined by a map into an inner basis."""

    def __init__(self, inner: Basis) -> None:
        self._inner = inner

    @property
    def inner(self) -> Basis:
        return self._inner

    def metadata(self) -> BasisMetadata:
        return self._inner.metadata()

    @property
    def is_dual(self) -> Any:
        return self._inner.is_dual

    def into_inner(self, t: torch.Tensor, axis: int = -1) -> torch.Tensor:
        raise NotImplementedError

    def from_inner(self, t: torch.Tensor, axis: int = -1) -> torch.Tensor:
        raise NotImplementedError

    def __into_inner__(self, vectors: Any, axis: int = -1) -> BasisConversion:
        return BasisConversion(lambda: self.into_inner(as_device(vectors), axis))

    def __from_inner__(self, vectors: Any, axis: int = -1) -> BasisConversion:
        return BasisConversion(lambda: self.from_inner(as_device(vectors), axis))

    def into_fundamental(self, t: torch.Tensor, axis: int = -1) -> torch.Tensor:
        return self._inner.into_fundamental(self.into_inner(t, axis), axis)

    def from_fundamental(self, t: torch.Tensor, axis: int = -1) -> torch.Tensor:
        return self.from_inner(self._inner.from_fundamental(t, axis), axis)

    def convert_vector_into(self, t: torch.Tensor, other: Basis, axis: int = -1) -> torch.Tensor:
        if self == other:
            return t
        # shortcut: the target wraps (or is) my inner basis -> no round trip through the fundamental grid
        if self._inner == other:
            return self.into_inner(t, axis)
        o = _strip(other)
        if isinstance(o, WrappedBasis) and not isinstance(o, AsUpcast) and o.inner == self:
            return o.from_inner(t, axis)
        return super().convert_vector_into(t, other, axis)


class AsUpcast(WrappedBasis):
    """Type-level re-labelling of a basis' metadata; numerically transparent."""

    def __init__(self, inner: Basis, metadata: BasisMetadata) -> None:
        super().__init__(inner)
        self._upcast_metadata = metadata

    def metadata(self) -> BasisMetadata:
        return self._upcast_metadata

    @property
    def size(self) -> int:
        return self._inner.size

    def upcast(self) -> "AsUpcast":
        return self

    def dual_basis(self) -> "AsUpcast":
        return AsUpcast(self._inner.dual_basis(), self._upcast_metadata)

    def into_inner(self, t: torch.Tensor, axis: int = -1) -> torch.Tensor:
        return t

    def from_inner(self, t: torch.Tensor, axis: int = -1) -> torch.Tensor:
        return t

    def into_fundamental(self, t: torch.Tensor, axis: int = -1) -> torch.Tensor:
        return self._inner.into_fundamental(t, axis)

    def from_fundamental(self, t: torch.Tensor, axis: int = -1) -> torch.Tensor:
        return self._inner.from_fundamental(t, axis)

    def convert_vector_into(self, t: torch.Tensor, other: Basis, axis: int = -1) -> torch.Tensor:
        return self._inner.convert_vector_into(t, _strip(other), axis)

    def __getattr__(self, name: str) -> Any:  # children / shape / transform ... of the wrapped basis
        if name.startswith("__"):
            raise AttributeError(name)
        return getattr(self._inner, name)

    def _key(self) -> tuple:
        return self._inner._key()


class TransformedBasis(WrappedBasis):
    """Momentum-like basis: data = FFT of the inner data.

    ket index: transformed -> inner is ifft(norm="ortho"); a dual (bra) index uses the opposite direction
    (convention pinned by the reference's tests/test_operator.py:197-204).
    """

    def __init__(self, inner: Basis) -> None:
        super().__init__(inner)

    @property
    def size(self) -> int:
        return self._inner.size

    def dual_basis(self) -> "TransformedBasis":
        return TransformedBasis(self._inner.dual_basis())

    def _fft(self, t: torch.Tensor, axis: int, sign: int) -> torch.Tensor:
        axis %= t.ndim
        n = t.shape[axis]
        outer = int(np.prod(t.shape[:axis])) if axis else 1
        inner = int(np.prod(t.shape[axis + 1 :])) if axis + 1 < t.ndim else 1
        return kernels.fft_axis(t.contiguous(), outer, n, inner, sign).reshape(t.shape)

    def into_inner(self, t: torch.Tensor, axis: int = -1) -> torch.Tensor:
        return self._fft(t, axis, -1 if self.is_dual else +1)

    def from_inner(self, t: torch.Tensor, axis: int = -1) -> torch.Tensor:
        return self._fft(t, axis, +1 if self.is_dual else -1)

    def _key(self) -> tuple:
        return ("T", self._inner._key())


class CroppedBasis(WrappedBasis):
    """Keeps the `size` lowest |frequency| entries of an FFT-ordered inner basis."""

    def __init__(self, size: int, inner: Basis) -> None:
        super().__init__(inner)
        self._size = int(size)

    @property
    def size(self) -> int:
        return self._size

    def dual_basis(self) -> "CroppedBasis":
        return CroppedBasis(self._size, self._inner.dual_basis())

    def _index(self) -> torch.Tensor:
        from slate_quantum_b200.core.metadata import fundamental_nk_points

        idx = fundamental_nk_points(self._size) % self._inner.size
        return torch.as_tensor(idx, device="cuda", dtype=torch.long)

    def into_inner(self, t: torch.Tensor, axis: int = -1) -> torch.Tensor:
        axis %= t.ndim
        shape = list(t.shape)
        shape[axis] = self._inner.size
        out = torch.zeros(shape, dtype=t.dtype, device=t.device)
        out.index_copy_(axis, self._index(), t)
        return out

    def from_inner(self, t: torch.Tensor, axis: int = -1) -> torch.Tensor:
        return t.index_select(axis % t.ndim, self._index())

    def _key(self) -> tuple:
        return ("C", self._size, self._inner._key())


class Truncation:
    def __init__(self, n: int, step: int, offset: int) -> None:
        self.n, self.step, self.offset = int(n), int(step), int(offset)

    def _key(self) -> tuple:
        return (self.n, self.step, self.offset)


class TruncatedBasis(WrappedBasis):
    """Entry j of the data sits at inner index offset + j*step."""

    def __init__(self, truncation: Truncation, inner: Basis) -> None:
        super().__init__(inner)
        self.truncation = truncation

    @property
    def size(self) -> int:
        return self.truncation.n

    def dual_basis(self) -> "TruncatedBasis":
        return TruncatedBasis(self.truncation, self._inner.dual_basis())

    def _index(self) -> torch.Tensor:
        tr = self.truncation
        idx = (tr.offset + tr.step * np.arange(tr.n)) % self._inner.size
        return torch.as_tensor(idx, device="cuda", dtype=torch.long)

    def into_inner(self, t: torch.Tensor, axis: int = -1) -> torch.Tensor:
        axis %= t.ndim
        shape = list(t.shape)
        shape[axis] = self._inner.size
        out = torch.zeros(shape, dtype=t.dtype, device=t.device)
        out.index_copy_(axis, self._index(), t)
        return out

    def from_inner(self, t: torch.Tensor, axis: int = -1) -> torch.Tensor:
        return t.index_select(axis % t.ndim, self._index())

    def _key(self) -> tuple:
        return ("Tr", self.truncation._key(), self._inner._key())


# ======================================================================================== tuples
class TupleBasis(Basis):
    def __init__(self, children: Sequence[Basis], extra: Any = None) -> None:
        self.children = tuple(children)
        self.extra = extra

    def metadata(self) -> TupleMetadata:
        return TupleMetadata(tuple(c.metadata() for c in self.children), self.extra)

    @property
    def shape(self) -> tuple[int, ...]:
        return tuple(c.size for c in self.children)

    @property
    def size(self) -> int:
        return int(np.prod(self.shape)) if self.children else 1

    @property
    def is_dual(self) -> tuple:
        return tuple(c.is_dual for c in self.children)

    def dual_basis(self) -> "TupleBasis":
        return TupleBasis(tuple(c.dual_basis() for c in self.children), self.extra)

    def _stack(self, t: torch.Tensor, axis: int, shape: Sequence[int]) -> torch.Tensor:
        return t.reshape(*t.shape[:axis], *shape, *t.shape[axis + 1 :])

    def into_fundamental(self, t: torch.Tensor, axis: int = -1) -> torch.Tensor:
        axis %= t.ndim
        lead, tail = t.shape[:axis], t.shape[axis + 1 :]
        s = self._stack(t, axis, self.shape)
        for i, c in enumerate(self.children):
            s = c.into_fundamental(s, axis + i)
        return s.reshape(*lead, -1, *tail)

    def from_fundamental(self, t: torch.Tensor, axis: int = -1) -> torch.Tensor:
        axis %= t.ndim
        lead, tail = t.shape[:axis], t.shape[axis + 1 :]
        s = self._stack(t, axis, tuple(c.fundamental_size for c in self.children))
        for i, c in enumerate(self.children):
            s = c.from_fundamental(s, axis + i)
        return s.reshape(*lead, -1, *tail)

    def convert_vector_into(self, t: torch.Tensor, other: Basis, axis: int = -1) -> torch.Tensor:
        o = _strip(other)
        if self == o:
            return t
        if isinstance(o, TupleBasis) and len(o.children) == len(self.children):
            axis %= t.ndim
            lead, tail = t.shape[:axis], t.shape[axis + 1 :]
            s = self._stack(t, axis, self.shape)
            for i, (c, oc) in enumerate(zip(self.children, o.children, strict=True)):
                s = c.convert_vector_into(s, oc, axis + i)
            return s.reshape(*lead, -1, *tail)
        return super().convert_vector_into(t, other, axis)

    def _key(self) -> tuple:
        return ("Tu", tuple(c._key() for c in (_strip(c) for c in self.children)), _extra_key(self.extra))


def _extra_key(extra: Any) -> Any:
    from slate_quantum_b200.core.metadata import AxisDirections

    if isinstance(extra, AxisDirections):
        return ("dirs", tuple(tuple(round(x, 12) for x in v) for v in extra.vectors))
    return extra


TupleBasis2D = TupleBasis
TupleBasisLike = Basis
TupleBasisLike2D = Basis


def as_tuple(basis: Basis) -> TupleBasis:
    """The TupleBasis a basis is (a wrapper of); wrappers such as BlockDiagonalBasis are looked through."""
    b = _strip(basis)
    while not isinstance(b, TupleBasis) and isinstance(b, WrappedBasis):
        b = _strip(b.inner)
    if isinstance(b, TupleBasis):
        return b
    msg = f"{type(b).__name__} is not a tuple basis"
    raise TypeError(msg)


def from_metadata(metadata: BasisMetadata, *, is_dual: Any = None) -> Basis:
    """Fundamental basis (tuple of fundamental bases for tuple metadata)."""
    if isinstance(metadata, TupleMetadata):
        duals = is_dual if isinstance(is_dual, tuple) else (is_dual,) * len(metadata.children)
        return TupleBasis(
            tuple(from_metadata(c, is_dual=d) for c, d in zip(metadata.children, duals, strict=True)), metadata.extra
        )
    return FundamentalBasis(metadata, is_dual=bool(is_dual))


def transformed_from_metadata(metadata: BasisMetadata, *, is_dual: Any = None) -> Basis:
    """TransformedBasis of every leaf (momentum basis)."""
    if isinstance(metadata, TupleMetadata):
        duals = is_dual if isinstance(is_dual, tuple) else (is_dual,) * len(metadata.children)
        return TupleBasis(
            tuple(transformed_from_metadata(c, is_dual=d) for c, d in zip(metadata.children, duals, strict=True)),
            metadata.extra,
        )
    return TransformedBasis(FundamentalBasis(metadata, is_dual=bool(is_dual)))


def with_modified_children(basis: TupleBasis, fn: Callable[[int, Basis], Basis]) -> TupleBasis:
    b = as_tuple(basis)
    return TupleBasis(tuple(fn(i, c) for i, c in enumerate(b.children)), b.extra)


def are_dual_shapes(lhs: Any, rhs: Any) -> bool:
    if isinstance(lhs, tuple) and isinstance(rhs, tuple):
        return len(lhs) == len(rhs) and all(are_dual_shapes(a, b) for a, b in zip(lhs, rhs, strict=True))
    if isinstance(lhs, tuple) or isinstance(rhs, tuple):
        return False
    return bool(lhs) != bool(rhs)


# ======================================================================================== operator-shaped
class DiagonalBasis(WrappedBasis):
    """Data = the diagonal of an operator over inner = TupleBasis((b0, b1))."""

    def __init__(self, inner: Basis) -> None:
        super().__init__(inner)
        assert len(as_tuple(inner).children) == 2  # noqa: PLR2004

    @property
    def size(self) -> int:
        return as_tuple(self._inner).children[0].size

    def dual_basis(self) -> "DiagonalBasis":
        return DiagonalBasis(self._inner.dual_basis())

    def into_inner(self, t: torch.Tensor, axis: int = -1) -> torch.Tensor:
        def fn(v: torch.Tensor) -> torch.Tensor:
            return torch.diag_embed(v).reshape(*v.shape[:-1], -1)

        return _moved(t, axis, fn)

    def from_inner(self, t: torch.Tensor, axis: int = -1) -> torch.Tensor:
        n = self.size

        def fn(v: torch.Tensor) -> torch.Tensor:
            return torch.diagonal(v.reshape(*v.shape[:-1], n, n), dim1=-2, dim2=-1).contiguous()

        return _moved(t, axis, fn)

    def _key(self) -> tuple:
        return ("D", self._inner._key())


class RecastBasis(WrappedBasis):
    """Data is stored as a vector in `outer_recast`; as a vector in `inner_recast` it is the inner's data."""

    def __init__(self, inner: Basis, inner_recast: Basis, outer_recast: Basis) -> None:
        super().__init__(inner)
        self.inner_recast = inner_recast
        self.outer_recast = outer_recast
        assert inner.size == inner_recast.size

    @property
    def size(self) -> int:
        return self.outer_recast.size

    def dual_basis(self) -> "RecastBasis":
        return RecastBasis(self._inner.dual_basis(), self.inner_recast.dual_basis(), self.outer_recast.dual_basis())

    def into_inner(self, t: torch.Tensor, axis: int = -1) -> torch.Tensor:
        return self.outer_recast.convert_vector_into(t, self.inner_recast, axis)

    def from_inner(self, t: torch.Tensor, axis: int = -1) -> torch.Tensor:
        return self.inner_recast.convert_vector_into(t, self.outer_recast, axis)

    def _key(self) -> tuple:
        return ("R", self._inner._key(), _strip(self.inner_recast)._key(), _strip(self.outer_recast)._key())


class SplitBasis(Basis):
    """Data = concatenate(lhs data, rhs data); the represented vector is the SUM of both parts."""

    def __init__(self, lhs: Basis, rhs: Basis) -> None:
        self.lhs, self.rhs = lhs, rhs

    def metadata(self) -> BasisMetadata:
        return self.lhs.metadata()

    @property
    def size(self) -> int:
        return self.lhs.size + self.rhs.size

    @property
    def is_dual(self) -> Any:
        return self.lhs.is_dual

    def dual_basis(self) -> "SplitBasis":
        return SplitBasis(self.lhs.dual_basis(), self.rhs.dual_basis())

    def into_fundamental(self, t: torch.Tensor, axis: int = -1) -> torch.Tensor:
        axis %= t.ndim
        a, b = t.narrow(axis, 0, self.lhs.size), t.narrow(axis, self.lhs.size, self.rhs.size)
        return self.lhs.into_fundamental(a.contiguous(), axis) + self.rhs.into_fundamental(b.contiguous(), axis)

    def convert_vector_into(self, t: torch.Tensor, other: Basis, axis: int = -1) -> torch.Tensor:
        if self == other:
            return t
        axis %= t.ndim
        a, b = t.narrow(axis, 0, self.lhs.size), t.narrow(axis, self.lhs.size, self.rhs.size)
        return self.lhs.convert_vector_into(a.contiguous(), other, axis) + self.rhs.convert_vector_into(
            b.contiguous(), other, axis
        )

    def from_fundamental(self, t: torch.Tensor, axis: int = -1) -> torch.Tensor:
        msg = "a SplitBasis has no unique from_fundamental"
        raise NotImplementedError(msg)

    def _key(self) -> tuple:
        return ("S", _strip(self.lhs)._key(), _strip(self.rhs)._key())


class BlockDiagonalBasis(WrappedBasis):
    """Operator over inner = TupleBasis((b0, b1)) whose only non-zero entries are diagonal blocks.

    Data layout [n_blocks, block_shape[0], block_shape[1]] (reference: bloch/build.py:131-135).
    """

    def __init__(self, inner: Basis, block_shape: Sequence[int]) -> None:
        super().__init__(inner)
        self.block_shape = tuple(int(s) for s in block_shape)
        c = as_tuple(inner).children
        assert c[0].size % self.block_shape[0] == 0
        assert c[1].size % self.block_shape[1] == 0

    @property
    def n_blocks(self) -> int:
        return as_tuple(self._inner).children[0].size // self.block_shape[0]

    @property
    def size(self) -> int:
        return self.n_blocks * self.block_shape[0] * self.block_shape[1]

    def dual_basis(self) -> "BlockDiagonalBasis":
        return BlockDiagonalBasis(self._inner.dual_basis(), self.block_shape)

    def into_inner(self, t: torch.Tensor, axis: int = -1) -> torch.Tensor:
        nb, (b0, b1) = self.n_blocks, self.block_shape

        def fn(v: torch.Tensor) -> torch.Tensor:
            blocks = v.reshape(-1, nb, b0, b1)
            dense = torch.zeros((blocks.shape[0], nb, b0, nb, b1), dtype=v.dtype, device=v.device)
            idx = torch.arange(nb, device=v.device)
            dense[:, idx, :, idx, :] = blocks.transpose(0, 1)
            return dense.reshape(*v.shape[:-1], nb * b0 * nb * b1)

        return _moved(t, axis, fn)

    def from_inner(self, t: torch.Tensor, axis: int = -1) -> torch.Tensor:
        nb, (b0, b1) = self.n_blocks, self.block_shape

        def fn(v: torch.Tensor) -> torch.Tensor:
            dense = v.reshape(-1, nb, b0, nb, b1)
            idx = torch.arange(nb, device=v.device)
            blocks = dense[:, idx, :, idx, :].transpose(0, 1)
            return blocks.reshape(*v.shape[:-1], nb * b0 * b1)

        return _moved(t, axis, fn)

    def _key(self) -> tuple:
        return ("B", self.block_shape, self._inner._key())


# ======================================================================================== explicit bases
class BasisStateMetadata(SimpleMetadata):
    """Metadata of 'the states of `basis`' (slate_core.basis.BasisStateMetadata)."""

    def __init__(self, basis: Basis) -> None:
        super().__init__(basis.size)
        self.basis = basis

    def _key(self) -> tuple:
        return ("BasisState", self._fundamental_size, _strip(self.basis)._key())


class ExplicitUnitaryBasis(WrappedBasis):
    """A basis whose states are the rows of a unitary `matrix` over (state index, inner basis).

    direction="forward": rows of `matrix` are the basis states written in the inner basis, so
    explicit -> inner is  v = sum_e a_e matrix[e, :]  and inner -> explicit is  a_e = conj(matrix[e, :]) . v
    (reference: state/_basis.py:25-111, slate_core ExplicitUnitaryBasis; assumption U6 of SURVEY.md).
    `matrix` may be dense ([n, n]) or block diagonal ([n_blocks, nb, nb] over a BlockDiagonalBasis).
    """

    def __init__(
        self,
        matrix: Any,
        *,
        direction: str = "forward",
        data_id: uuid.UUID | None = None,
        assert_unitary: bool = False,
    ) -> None:
        state_meta = matrix.basis.metadata().children[1]
        if not isinstance(state_meta, BasisStateMetadata):
            msg = "the second index of an explicit-basis matrix must carry BasisStateMetadata"
            raise TypeError(msg)
        super().__init__(state_meta.basis)
        if direction != "forward":
            msg = "only direction='forward' is implemented"
            raise NotImplementedError(msg)
        self._matrix = matrix
        self._direction = direction
        self._data_id = data_id or uuid.uuid4()
        self._dual = False
        del assert_unitary

    @property
    def data_id(self) -> uuid.UUID:
        return self._data_id

    def transform(self) -> Any:
        return self._matrix

    @property
    def size(self) -> int:
        return self._blocks().shape[0] * self._blocks().shape[1]

    @property
    def is_dual(self) -> Any:
        return self._dual

    def dual_basis(self) -> "ExplicitUnitaryBasis":
        out = type(self)(self._matrix, direction=self._direction, data_id=self._data_id)
        out._dual = not self._dual
        # a bra index converts through the DUAL of the inner basis (opposite FFT direction further down the chain)
        out._inner = self._inner.dual_basis()
        return out

    def _blocks(self) -> torch.Tensor:
        """transform as a [n_blocks, nb, nb] CUDA tensor (dense = one block)."""
        data = self._matrix.device_data
        b = _strip(self._matrix.basis)
        if isinstance(b, BlockDiagonalBasis):
            return data.reshape(b.n_blocks, *b.block_shape)
        n = int(round(np.sqrt(data.numel())))
        return data.reshape(1, n, n)

    def into_inner(self, t: torch.Tensor, axis: int = -1) -> torch.Tensor:
        v = self._blocks()
        if self._dual:
            # bra index: conj(V) acts, i.e. conj(apply(V, conj(x))).  conj_physical MATERIALISES the conjugate - a lazy
            # .conj() only sets a view bit that the raw-pointer C ABI cannot see
            return _moved(t, axis, lambda x: torch.conj_physical(kernels.apply_transform(v, torch.conj_physical(x), 0)))
        return _moved(t, axis, lambda x: kernels.apply_transform(v, x, 0))

    def from_inner(self, t: torch.Tensor, axis: int = -1) -> torch.Tensor:
        v = self._blocks()
        if self._dual:
            return _moved(t, axis, lambda x: torch.conj_physical(kernels.apply_transform(v, torch.conj_physical(x), 1)))
        return _moved(t, axis, lambda x: kernels.apply_transform(v, x, 1))

    def eigenvectors(self) -> Any:
        """The basis states as a list over (state index, inner basis): slate_core recasts the
        BasisStateMetadata index of `transform` back to the inner basis (reference state/_basis.py:82-111)."""
        from slate_quantum_b200.core.array import Array

        blocks = self._blocks()
        n_blocks, n, _ = blocks.shape
        tup = TupleBasis((FundamentalBasis(SimpleMetadata(n_blocks * n)), self.inner))
        out_basis = BlockDiagonalBasis(tup, (n, n)).upcast() if n_blocks > 1 else tup.upcast()
        return Array(out_basis, blocks.reshape(-1))

    def _key(self) -> tuple:
        return ("E", self._data_id, self._dual)


def _chain_int(*xs: Sequence[int]) -> tuple[int, ...]:
    return tuple(itertools.chain(*xs))
