"""Minimal slate_core.metadata surface for the Bloch hot path (written fresh, Python 3.12 compatible).

Only the classes / helpers that slate_quantum's hot path touches exist here.
Everything is host-side bookkeeping: no arithmetic on state data happens in this module.
"""

from __future__ import annotations

from dataclasses import dataclass
from typing import Any, Literal, Sequence

import numpy as np


class BasisMetadata:
    """Base class: something with a (possibly nested) fundamental shape."""

    @property
    def fundamental_shape(self) -> Any:
        raise NotImplementedError

    @property
    def fundamental_size(self) -> int:
        return size_from_nested_shape(self.fundamental_shape)

    @property
    def features(self) -> set[str]:
        return {SIMPLE_FEATURE}


SIMPLE_FEATURE = "SIMPLE"


def size_from_nested_shape(shape: Any) -> int:
    if isinstance(shape, (int, np.integer)):
        return int(shape)
    return int(np.prod([size_from_nested_shape(s) for s in shape])) if len(shape) else 1


class SimpleMetadata(BasisMetadata):
    def __init__(self, fundamental_size: int) -> None:
        self._fundamental_size = int(fundamental_size)

    @property
    def fundamental_shape(self) -> int:
        return self._fundamental_size

    @property
    def fundamental_size(self) -> int:
        return self._fundamental_size

    def _key(self) -> tuple:
        return (type(self).__name__, self._fundamental_size)

    def __eq__(self, other: object) -> bool:
        return isinstance(other, SimpleMetadata) and self._key() == other._key()

    def __hash__(self) -> int:
        return hash(self._key())

    def __repr__(self) -> str:
        return f"{type(self).__name__}({self._fundamental_size})"


class LabeledMetadata(SimpleMetadata):
    """Metadata whose points carry values."""

    @property
    def values(self) -> np.ndarray:
        raise NotImplementedError


class ExplicitLabeledMetadata(LabeledMetadata):
    def __init__(self, values: np.ndarray) -> None:
        self._values = np.asarray(values)
        super().__init__(self._values.size)

    @property
    def values(self) -> np.ndarray:
        return self._values

    def _key(self) -> tuple:
        return (type(self).__name__, self._fundamental_size, self._values.tobytes())


@dataclass(frozen=True)
class Domain:
    delta: float
    start: float = 0.0


class SpacedMetadata(LabeledMetadata):
    """Metadata with a domain."""


class EvenlySpacedMetadata(SpacedMetadata):
    def __init__(
        self, fundamental_size: int, *, domain: Domain, interpolation: Literal["Fourier", "DST"] = "Fourier"
    ) -> None:
        super().__init__(fundamental_size)
        self.domain = domain
        self.interpolation = interpolation

    @property
    def values(self) -> np.ndarray:
        """Sample positions.

        "Fourier" (periodic): start + delta * j / n.  "DST" (hard wall): the interior points
        start + delta * (j + 1) / (n + 1).  slate_core's own definition is not readable here
        (assumption U8 in SURVEY.md); the dynamics API always takes the values from this object.
        """
        n = self.fundamental_size
        if self.interpolation == "Fourier":
            return self.domain.start + self.domain.delta * np.arange(n) / n
        return self.domain.start + self.domain.delta * (np.arange(n) + 1) / (n + 1)

    @property
    def delta(self) -> float:
        return self.domain.delta

    def _key(self) -> tuple:
        return (
            "EvenlySpaced",
            self._fundamental_size,
            float(self.domain.start),
            float(self.domain.delta),
            self.interpolation,
        )


class LengthMetadata(SpacedMetadata):
    """Metadata of an axis with units of length."""


class EvenlySpacedLengthMetadata(EvenlySpacedMetadata, LengthMetadata):
    pass


@dataclass(frozen=True)
class AxisDirections:
    """Unit vectors of the axes of a volume (the `extra` of a volume TupleMetadata)."""

    vectors: tuple[tuple[float, ...], ...]

    @staticmethod
    def from_vectors(vectors: Sequence[np.ndarray]) -> "AxisDirections":
        return AxisDirections(tuple(tuple(float(x) for x in v) for v in vectors))

    def as_array(self) -> np.ndarray:
        return np.asarray(self.vectors, dtype=np.float64)


class TupleMetadata(BasisMetadata):
    def __init__(self, children: Sequence[BasisMetadata], extra: Any = None) -> None:
        self.children = tuple(children)
        self.extra = extra

    @property
    def fundamental_shape(self) -> tuple:
        return tuple(c.fundamental_shape for c in self.children)

    @property
    def shape(self) -> tuple[int, ...]:
        return tuple(c.fundamental_size for c in self.children)

    def __eq__(self, other: object) -> bool:
        return (
            isinstance(other, TupleMetadata)
            and self.children == other.children
            and _extra_eq(self.extra, other.extra)
        )

    def __hash__(self) -> int:
        return hash((self.children, _extra_hash(self.extra)))

    def __repr__(self) -> str:
        return f"TupleMetadata({self.children!r}, {self.extra!r})"


def _extra_eq(a: Any, b: Any) -> bool:
    if isinstance(a, AxisDirections) and isinstance(b, AxisDirections):
        return np.allclose(a.as_array(), b.as_array(), rtol=1e-12, atol=0)
    return a == b


def _extra_hash(a: Any) -> int:
    return hash(len(a.vectors)) if isinstance(a, AxisDirections) else hash(a)


EvenlySpacedVolumeMetadata = TupleMetadata
VolumeMetadata = TupleMetadata


def fundamental_nk_points(metadata: SimpleMetadata | int) -> np.ndarray:
    """FFT-ordered integer frequencies (n=5 -> [0,1,2,-2,-1]; pinned at reference tests/test_operator.py:195)."""
    n = metadata if isinstance(metadata, (int, np.integer)) else metadata.fundamental_size
    return np.fft.ifftshift(np.arange((-n + 1) // 2, (n + 1) // 2))


def fundamental_nx_points(metadata: SimpleMetadata | int) -> np.ndarray:
    n = metadata if isinstance(metadata, (int, np.integer)) else metadata.fundamental_size
    return np.arange(n)


def spaced_volume_metadata_from_stacked_delta_x(
    vectors: Sequence[np.ndarray], shape: Sequence[int], *, interpolation: Literal["Fourier", "DST"] = "Fourier"
) -> TupleMetadata:
    """Volume metadata from the FULL cell vectors `vectors[i]` sampled by shape[i] points.

    (reference call sites: tests/test_bloch.py:24-27, tests/test_operator.py:22-26.)
    Vectors shorter/longer than the number of axes are kept as given (test_operator.py:305 uses a
    2-component vector for a 1-axis volume): only the first len(shape) components span the cell.
    """
    vecs = [np.asarray(v, dtype=np.float64) for v in vectors]
    lengths = [float(np.linalg.norm(v)) for v in vecs]
    children = tuple(
        EvenlySpacedLengthMetadata(int(n), domain=Domain(delta=length), interpolation=interpolation)
        for n, length in zip(shape, lengths, strict=True)
    )
    directions = AxisDirections.from_vectors([v / length for v, length in zip(vecs, lengths, strict=True)])
    return TupleMetadata(children, directions)


class volume:  # noqa: N801 - mirrors the module name slate_core.metadata.volume
    """Namespace mirroring ``slate_core.metadata.volume``."""

    spaced_volume_metadata_from_stacked_delta_x = staticmethod(spaced_volume_metadata_from_stacked_delta_x)

    @staticmethod
    def fundamental_stacked_delta_x(metadata: TupleMetadata) -> np.ndarray:
        """Row i = full cell vector a_i."""
        dirs = metadata.extra.as_array()
        return np.array([c.domain.delta * d for c, d in zip(metadata.children, dirs, strict=True)])

    @staticmethod
    def fundamental_stacked_dx(metadata: TupleMetadata) -> np.ndarray:
        """Row i = grid step along axis i, a_i / N_i."""
        shape = np.asarray(metadata.shape, dtype=np.float64)
        return volume.fundamental_stacked_delta_x(metadata) / shape[:, np.newaxis]

    @staticmethod
    def fundamental_stacked_delta_k(metadata: TupleMetadata) -> np.ndarray:
        """Row i = N_i * dk_i."""
        dk = volume.fundamental_stacked_dk(metadata)
        return dk * np.asarray(metadata.shape, dtype=np.float64)[:, np.newaxis]

    @staticmethod
    def fundamental_stacked_dk(metadata: TupleMetadata) -> np.ndarray:
        """Reciprocal vectors, row i = dk_i with dk_i . a_j = 2 pi delta_ij (square cells only)."""
        a = volume.fundamental_stacked_delta_x(metadata)
        nd = len(metadata.children)
        a = a[:, :nd]
        return 2 * np.pi * np.linalg.inv(a).T

    @staticmethod
    def fundamental_stacked_x_points(
        metadata: TupleMetadata, *, offset: Sequence[float] | None = None, wrapped: bool = False
    ) -> tuple[np.ndarray, ...]:
        """Cartesian coordinates of every grid point, C-order ravel.

        ``offset`` is a Cartesian displacement (the reference passes lengths, tests/test_measure.py:40-42);
        ``wrapped`` folds the fractional coordinates into [-1/2, 1/2).
        """
        a = volume.fundamental_stacked_delta_x(metadata)
        nd = len(metadata.children)
        a = a[:nd, :nd]
        shape = np.array([c.fundamental_size for c in metadata.children], dtype=np.float64)
        idx = np.array(
            [g.ravel() for g in np.meshgrid(*(np.arange(int(n)) for n in shape), indexing="ij")], dtype=np.float64
        )
        points = np.einsum("ij,ik->jk", a / shape[:, None], idx)
        if offset is not None:
            points = points + np.asarray(offset, dtype=np.float64)[:nd, None]
        if wrapped:
            frac = np.linalg.solve(a.T, points)
            points = points - np.einsum("ij,ik->jk", a, np.floor(frac + 0.5))
        return tuple(points)

    @staticmethod
    def fundamental_stacked_k_points(
        metadata: TupleMetadata, *, offset: Sequence[float] | None = None, wrapped: bool = False
    ) -> tuple[np.ndarray, ...]:
        """Cartesian wavevector of every momentum-basis index (FFT order per axis), C-order ravel; ``offset`` is a
        Cartesian displacement, ``wrapped`` folds the points into the first zone [-K/2, K/2), K_a = N_a dk_a."""
        dk = volume.fundamental_stacked_dk(metadata)
        nd = len(metadata.children)
        nk = np.array(
            [g.ravel() for g in np.meshgrid(*(fundamental_nk_points(c) for c in metadata.children), indexing="ij")],
            dtype=np.float64,
        )
        points = np.einsum("ij,ik->jk", dk, nk)
        if offset is not None:
            points = points + np.asarray(offset, dtype=np.float64)[:nd, None]
        if wrapped:
            big = volume.fundamental_stacked_delta_k(metadata)
            frac = np.linalg.solve(big.T, points)
            points = points - np.einsum("ij,ik->jk", big, np.floor(frac + 0.5))
        return tuple(points)


fundamental_stacked_dk = volume.fundamental_stacked_dk
fundamental_stacked_delta_k = volume.fundamental_stacked_delta_k
