"""Minimal slate_core.Array: (basis, raw_data) with GPU-backed basis conversion.

`raw_data` is exposed as a numpy array exactly like the reference (flat, basis ordered); internally the
data may live on the GPU (`device_data`, a torch CUDA tensor) and is only copied to the host on demand,
so chained hot-path calls (build -> into_diagonal -> evolve) never leave HBM.
"""

from __future__ import annotations

from typing import Any

import numpy as np
import torch

from slate_quantum_b200 import kernels
from slate_quantum_b200.core import basis as _basis
from slate_quantum_b200.core.basis import Basis


class Array:
    def __init__(self, basis: Basis, data: Any) -> None:
        self._basis = basis
        self._host: np.ndarray | None = None
        self._device: torch.Tensor | None = None
        self._host_tensor: torch.Tensor | None = None
        if isinstance(data, torch.Tensor) and not data.is_cuda:
            # a HOST torch tensor (e.g. pinned memory): kept as it is and uploaded asynchronously on first use
            self._host_tensor = data
            self._host = data.detach().numpy()
            n = data.numel()
        elif isinstance(data, torch.Tensor):
            self._device = data
            n = data.numel()
        else:
            self._host = np.asarray(data)
            n = self._host.size
        if n != basis.size:
            msg = f"data has {n} entries but the basis has size {basis.size}"
            raise ValueError(msg)

    @property
    def basis(self) -> Basis:
        return self._basis

    @property
    def raw_data(self) -> np.ndarray:
        if self._host is None:
            self._host = self._device.detach().cpu().numpy()
        return self._host

    @property
    def device_data(self) -> torch.Tensor:
        """The data as a CUDA tensor (uploaded once if the array was created from numpy)."""
        if self._device is None and self._host_tensor is not None:
            dtype = torch.complex128 if self._host_tensor.is_complex() else torch.float64
            self._device = self._host_tensor.to(device="cuda", dtype=dtype, non_blocking=True).contiguous()
        if self._device is None:
            dtype = torch.complex128 if np.iscomplexobj(self._host) else torch.float64
            self._device = kernels.to_device(np.ascontiguousarray(self._host).astype(
                np.complex128 if dtype == torch.complex128 else np.float64), dtype)
        return self._device

    @property
    def is_on_device(self) -> bool:
        return self._device is not None

    @property
    def dtype(self) -> np.dtype:
        return self._host.dtype if self._host is not None else np.dtype(np.complex128)

    @property
    def fundamental_shape(self) -> Any:
        return self._basis.metadata().fundamental_shape

    def _wrap(self, basis: Basis, data: Any) -> "Array":
        return type(self)(basis, data)

    def with_basis(self, basis: Basis) -> "Array":
        assert basis.metadata() == self.basis.metadata()
        data = self.basis.convert_vector_into(_complex(self.device_data).reshape(-1), basis, 0)
        return self._wrap(basis, data)

    def as_array(self) -> np.ndarray:
        """Data in the fundamental basis, shaped like the metadata's (flattened-per-child) shape."""
        meta = self.basis.metadata()
        fund = self.basis.into_fundamental(_complex(self.device_data).reshape(-1), 0)
        out = fund.detach().cpu().numpy()
        shape = _flat_shape(meta)
        return out.reshape(shape)

    @staticmethod
    def from_array(data: np.ndarray) -> "Array":
        data = np.asarray(data)
        b = _basis.TupleBasis(tuple(_basis.FundamentalBasis.from_size(s) for s in data.shape))
        return Array(b, data)

    # ---- arithmetic (device) ---------------------------------------------------------------
    def _binary(self, other: Any, op: str) -> "Array":
        if isinstance(other, Array):
            rhs = other.basis.convert_vector_into(_complex(other.device_data).reshape(-1), self.basis, 0)
        else:
            msg = "only Array +- Array is implemented on the hot path"
            raise NotImplementedError(msg)
        lhs = _complex(self.device_data).reshape(-1)
        return self._wrap(self.basis, lhs + rhs if op == "+" else lhs - rhs)

    def __add__(self, other: Any) -> "Array":
        return self._binary(other, "+")

    def __sub__(self, other: Any) -> "Array":
        return self._binary(other, "-")

    def __mul__(self, other: complex) -> "Array":
        return self._wrap(self.basis, _complex(self.device_data) * complex(other))

    __rmul__ = __mul__


def _complex(t: torch.Tensor) -> torch.Tensor:
    return t if t.is_complex() else t.to(torch.complex128)


def _flat_shape(meta: Any) -> tuple[int, ...]:
    children = getattr(meta, "children", None)
    if children is None:
        return (meta.fundamental_size,)
    return tuple(c.fundamental_size for c in children)


def cast_basis(array: Array, basis: Basis) -> Array:
    """Re-label the basis without touching the data (slate_core.array.cast_basis)."""
    assert basis.size == array.basis.size
    data = array.device_data if array.is_on_device else array.raw_data
    return type(array)(basis, data)


def as_tuple_basis(array: Array) -> Array:
    b = _basis._strip(array.basis)  # noqa: SLF001
    if isinstance(b, _basis.TupleBasis):
        return array
    meta = array.basis.metadata()
    return array.with_basis(_basis.from_metadata(meta))
