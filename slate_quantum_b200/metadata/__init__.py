"""Mirror of slate_quantum.metadata (reference: slate_quantum/metadata/_label.py, _repeat.py)."""

from __future__ import annotations

from typing import Literal

import numpy as np

from slate_quantum_b200.core.basis import FundamentalBasis
from slate_quantum_b200.core.metadata import (
    AxisDirections,
    Domain,
    EvenlySpacedLengthMetadata,
    EvenlySpacedMetadata,
    ExplicitLabeledMetadata,
    SpacedMetadata,
    TupleMetadata,
)


class TimeMetadata(SpacedMetadata):
    """Metadata of a time axis (reference metadata/_label.py:13-14)."""


class SpacedTimeMetadata(EvenlySpacedMetadata, TimeMetadata):
    """Evenly spaced times; not periodic, so "DST" by default (reference metadata/_label.py:17-28)."""

    def __init__(
        self, fundamental_size: int, *, domain: Domain, interpolation: Literal["Fourier", "DST"] = "DST"
    ) -> None:
        super().__init__(fundamental_size, domain=domain, interpolation=interpolation)


class MomentumMetadata(SpacedMetadata):
    pass


class EvenlySpacedMomentumMetadata(EvenlySpacedMetadata, MomentumMetadata):
    pass


class EigenvalueMetadata(ExplicitLabeledMetadata):
    """Metadata labelled by eigenvalues (reference metadata/_label.py:39-40)."""


def eigenvalue_basis(values: np.ndarray) -> FundamentalBasis:
    return FundamentalBasis(EigenvalueMetadata(values))


class RepeatedMetadata(EvenlySpacedLengthMetadata):
    """A unit-cell axis repeated n_repeats times (reference metadata/_repeat.py:9-25)."""

    def __init__(self, inner: EvenlySpacedLengthMetadata, n_repeats: int) -> None:
        self._inner = inner
        self.n_repeats = int(n_repeats)
        super().__init__(
            inner.fundamental_size * self.n_repeats,
            domain=Domain(start=inner.domain.start, delta=self.n_repeats * inner.domain.delta),
            interpolation=inner.interpolation,
        )

    @property
    def inner(self) -> EvenlySpacedLengthMetadata:
        return self._inner


RepeatedVolumeMetadata = TupleMetadata


def repeat_volume_metadata(metadata: TupleMetadata, shape: tuple[int, ...]) -> TupleMetadata:
    """reference metadata/_repeat.py:34-45."""
    return TupleMetadata(
        tuple(RepeatedMetadata(d, s) for (s, d) in zip(shape, metadata.children, strict=True)), metadata.extra
    )


def unit_cell_metadata(metadata: TupleMetadata) -> TupleMetadata:
    """reference metadata/_repeat.py:48-51."""
    return TupleMetadata(tuple(c.inner for c in metadata.children), metadata.extra)


__all__ = [
    "AxisDirections",
    "EigenvalueMetadata",
    "EvenlySpacedMomentumMetadata",
    "MomentumMetadata",
    "RepeatedMetadata",
    "RepeatedVolumeMetadata",
    "SpacedTimeMetadata",
    "TimeMetadata",
    "eigenvalue_basis",
    "repeat_volume_metadata",
    "unit_cell_metadata",
]
