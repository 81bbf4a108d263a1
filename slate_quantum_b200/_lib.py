"""ctypes binding of libslate_b200.so (the C ABI declared in include/slate_b200.h).

There is NO CPU fallback: if the shared library is missing or no CUDA device is present the calls
raise.  PyTorch is used only to own device buffers and to provide the current stream.
"""

from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_char_p, c_double, c_int, c_int64, c_size_t, c_void_p
from pathlib import Path

_HERE = Path(__file__).resolve().parent
_LIB_PATH = Path(os.environ.get("SLATE_B200_LIB", _HERE / "libslate_b200.so"))

EXPORTED_SYMBOLS = (
    "sqb_abi_version",
    "sqb_error_string",
    "sqb_bloch_build",
    "sqb_bloch_kinetic",
    "sqb_kinetic_energy",
    "sqb_bloch_permute",
    "sqb_fft_workspace_bytes",
    "sqb_fft_c2c",
    "sqb_fft_axis",
    "sqb_cell_fft_workspace_bytes",
    "sqb_cell_fft",
    "sqb_cell_fft_sharded",
    "sqb_cell_fft_moments_workspace_bytes",
    "sqb_cell_fft_moments",
    "sqb_fold_transform",
    "sqb_bloch_cell_vectors",
    "sqb_bloch_time_reversal_mirror",
    "sqb_eigh_workspace_bytes",
    "sqb_eigh_batched",
    "sqb_tridiag_eigh_workspace_bytes",
    "sqb_tridiag_eigh_batched",
    "sqb_project",
    "sqb_apply_transform",
    "sqb_evolve_eigen",
    "sqb_evolve_bloch",
    "sqb_evolve_uniform_route",
    "sqb_evolve_uniform_workspace_bytes",
    "sqb_evolve_bloch_uniform",
    "sqb_measure_moments",
    "sqb_measure_moments_lattice",
    "sqb_zgemm_batched",
    "sqb_rowwise_vdot",
    "sqb_rowwise_weighted_abs2",
    "sqb_scale_columns",
    "sqb_abs2",
    "sqb_normalize_rows",
    "sqb_peak_dmma",
    "sqb_peak_dfma",
)


class SlateB200Error(RuntimeError):
    """Raised when the native library is missing or a C-ABI call returns a non-zero status."""


_lib: ctypes.CDLL | None = None


def load() -> ctypes.CDLL:
    """Load (once) and type the shared library.  Raises SlateB200Error when it is not built."""
    global _lib  # noqa: PLW0603
    if _lib is not None:
        return _lib
    if not _LIB_PATH.exists():
        msg = (
            f"{_LIB_PATH} not found: build it with `python -m slate_quantum_b200.build` "
            "(nvcc, sm_100a). slate_quantum_b200 has no CPU fallback."
        )
        raise SlateB200Error(msg)
    lib = ctypes.CDLL(str(_LIB_PATH))
    ip = POINTER(c_int)
    dp = POINTER(c_double)
    lib.sqb_abi_version.restype = c_int
    lib.sqb_abi_version.argtypes = []
    lib.sqb_error_string.restype = c_char_p
    lib.sqb_error_string.argtypes = [c_int]
    lib.sqb_bloch_build.restype = c_int
    lib.sqb_bloch_build.argtypes = [c_int, ip, ip, dp, c_void_p, c_double, c_double, c_int64, c_int64, c_void_p, c_void_p]
    lib.sqb_bloch_kinetic.restype = c_int
    lib.sqb_bloch_kinetic.argtypes = [c_int, ip, ip, dp, c_double, c_double, c_int64, c_int64, c_void_p, c_void_p]
    lib.sqb_kinetic_energy.restype = c_int
    lib.sqb_kinetic_energy.argtypes = [c_int, ip, dp, dp, c_double, c_double, c_void_p, c_void_p]
    lib.sqb_bloch_permute.restype = c_int
    lib.sqb_bloch_permute.argtypes = [c_int, ip, ip, c_int, c_int64, c_void_p, c_void_p, c_void_p]
    lib.sqb_fft_workspace_bytes.restype = c_size_t
    lib.sqb_fft_workspace_bytes.argtypes = [c_int, ip, c_int64]
    lib.sqb_fft_c2c.restype = c_int
    lib.sqb_fft_c2c.argtypes = [c_int, ip, c_int64, c_int, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]
    lib.sqb_fft_axis.restype = c_int
    lib.sqb_fft_axis.argtypes = [c_int64, c_int, c_int64, c_int, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]
    lib.sqb_cell_fft_workspace_bytes.restype = c_size_t
    lib.sqb_cell_fft_workspace_bytes.argtypes = [c_int, ip]
    lib.sqb_cell_fft.restype = c_int
    lib.sqb_cell_fft.argtypes = [c_int, ip, ip, c_int, c_int64, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]
    lib.sqb_cell_fft_sharded.restype = c_int
    lib.sqb_cell_fft_sharded.argtypes = [c_int, ip, ip, c_int64, c_int, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]
    lib.sqb_cell_fft_moments_workspace_bytes.restype = c_size_t
    lib.sqb_cell_fft_moments_workspace_bytes.argtypes = [c_int, ip, ip, c_int64]
    lib.sqb_cell_fft_moments.restype = c_int
    lib.sqb_cell_fft_moments.argtypes = [
        c_int, ip, ip, c_int64, c_int, c_void_p, dp, dp, c_void_p, c_int, c_void_p, c_void_p, c_size_t, c_void_p,
    ]
    lib.sqb_bloch_cell_vectors.restype = c_int
    lib.sqb_bloch_cell_vectors.argtypes = [c_int, ip, ip, c_int, c_int, c_int64, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]
    lib.sqb_fold_transform.restype = c_int
    lib.sqb_fold_transform.argtypes = [c_int, ip, ip, c_int64, c_int64, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]
    lib.sqb_eigh_workspace_bytes.restype = c_size_t
    lib.sqb_eigh_workspace_bytes.argtypes = [c_int, c_int64]
    lib.sqb_eigh_batched.restype = c_int
    lib.sqb_eigh_batched.argtypes = [c_int, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]
    lib.sqb_measure_moments.restype = c_int
    lib.sqb_measure_moments.argtypes = [
        c_int, c_void_p, c_int64, c_void_p, c_int, c_double, c_double, c_void_p, c_int, c_void_p, c_void_p,
    ]
    lib.sqb_measure_moments_lattice.restype = c_int
    lib.sqb_measure_moments_lattice.argtypes = [
        c_int, c_void_p, c_int64, c_void_p, c_int, dp, c_void_p, c_int, c_void_p, c_void_p,
    ]
    lib.sqb_zgemm_batched.restype = c_int
    lib.sqb_zgemm_batched.argtypes = [
        c_int, c_int, c_int, c_int64, c_void_p, c_int64, c_int64, c_int64, c_int, c_void_p, c_int64, c_int64, c_int64,
        c_int, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_void_p,
    ]
    lib.sqb_rowwise_vdot.restype = c_int
    lib.sqb_rowwise_vdot.argtypes = [c_int64, c_int64, c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_void_p]
    lib.sqb_rowwise_weighted_abs2.restype = c_int
    lib.sqb_rowwise_weighted_abs2.argtypes = [c_int64, c_int64, c_void_p, c_int64, c_void_p, c_void_p, c_void_p]
    lib.sqb_scale_columns.restype = c_int
    lib.sqb_scale_columns.argtypes = [c_int64, c_int64, c_void_p, c_int64, c_void_p, c_void_p, c_void_p]
    lib.sqb_abs2.restype = c_int
    lib.sqb_abs2.argtypes = [c_int64, c_void_p, c_void_p, c_void_p]
    lib.sqb_normalize_rows.restype = c_int
    lib.sqb_normalize_rows.argtypes = [c_int64, c_int64, c_void_p, c_void_p, c_void_p, c_void_p]
    lib.sqb_tridiag_eigh_workspace_bytes.restype = c_size_t
    lib.sqb_tridiag_eigh_workspace_bytes.argtypes = [c_int, c_int64]
    lib.sqb_tridiag_eigh_batched.restype = c_int
    lib.sqb_tridiag_eigh_batched.argtypes = [
        c_int, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p,
    ]
    lib.sqb_project.restype = c_int
    lib.sqb_project.argtypes = [c_int, c_int64, c_void_p, c_void_p, c_void_p, c_void_p]
    lib.sqb_apply_transform.restype = c_int
    lib.sqb_apply_transform.argtypes = [c_int, c_int64, c_int64, c_int, c_void_p, c_void_p, c_void_p, c_void_p]
    lib.sqb_evolve_eigen.restype = c_int
    lib.sqb_evolve_eigen.argtypes = [c_int64, c_void_p, c_void_p, c_void_p, c_int64, c_double, c_void_p, c_void_p]
    lib.sqb_evolve_bloch.restype = c_int
    lib.sqb_evolve_bloch.argtypes = [
        c_int, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_double, c_void_p, c_int64, c_void_p,
    ]
    lib.sqb_bloch_time_reversal_mirror.restype = c_int
    lib.sqb_bloch_time_reversal_mirror.argtypes = [c_int, ip, ip, c_int64, c_int64, c_void_p, c_void_p, c_void_p]
    lib.sqb_evolve_uniform_route.restype = c_int
    lib.sqb_evolve_uniform_route.argtypes = [c_int, c_int64, c_int64]
    lib.sqb_evolve_uniform_workspace_bytes.restype = c_size_t
    lib.sqb_evolve_uniform_workspace_bytes.argtypes = [c_int, c_int64, c_int64]
    lib.sqb_evolve_bloch_uniform.restype = c_int
    lib.sqb_evolve_bloch_uniform.argtypes = [
        c_int, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_double, c_double, c_void_p, c_int64,
        c_void_p, c_size_t, c_int, c_void_p,
    ]
    lib.sqb_peak_dmma.restype = c_int
    lib.sqb_peak_dmma.argtypes = [c_int, c_int64, c_void_p, dp, c_void_p]
    lib.sqb_peak_dfma.restype = c_int
    lib.sqb_peak_dfma.argtypes = [c_int64, c_void_p, dp, c_void_p]
    _lib = lib
    return lib


def lib_path() -> Path:
    return _LIB_PATH


def check(code: int, what: str) -> None:
    if code != 0:
        msg = load().sqb_error_string(code).decode()
        raise SlateB200Error(f"{what} failed with status {code}: {msg}")


def int_array(values) -> ctypes.Array:
    vals = [int(v) for v in values]
    return (c_int * len(vals))(*vals)


def double_array(values) -> ctypes.Array:
    vals = [float(v) for v in values]
    return (c_double * len(vals))(*vals)
