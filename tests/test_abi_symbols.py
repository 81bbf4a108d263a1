"""CPU: the C-ABI library builds, loads and exports every symbol include/slate_b200.h declares."""

import ctypes
import re
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def _declared() -> set[str]:
    text = (ROOT / "include" / "slate_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return set(re.findall(r"\b(sqb_[a-z0-9_]+)\s*\(", text))


def test_library_builds_and_exports_header_symbols():
    from slate_quantum_b200 import _lib, build

    lib_path = build.build()
    assert lib_path.exists()
    lib = ctypes.CDLL(str(lib_path))
    declared = _declared()
    assert declared, "no declarations parsed"
    for sym in sorted(declared):
        assert hasattr(lib, sym), f"{sym} declared in include/slate_b200.h but not exported"
    assert declared == set(_lib.EXPORTED_SYMBOLS), declared ^ set(_lib.EXPORTED_SYMBOLS)
    handle = _lib.load()
    assert handle.sqb_abi_version() == 1
    assert handle.sqb_error_string(-1).decode().startswith("bad argument")
    # argument validation happens before any CUDA call, so it is testable without a GPU
    assert handle.sqb_eigh_workspace_bytes(600, 4) == 0
    assert handle.sqb_eigh_workspace_bytes(256, 4) > 4 * 256 * 256 * 16
    assert handle.sqb_bloch_permute(0, None, None, 0, 1, None, None, None) == -1


def test_sass_uses_fp64_tensor_cores_and_bulk_copies():
    """The hot GEMM must compile to DMMA and its B operand must come through bulk (TMA-unit) copies."""
    import shutil
    import subprocess

    from slate_quantum_b200 import build

    if shutil.which("cuobjdump") is None:
        return
    sass = subprocess.run(
        ["cuobjdump", "-sass", str(build.build())], capture_output=True, text=True, check=True
    ).stdout
    assert "DMMA.8x8x4" in sass
    assert "UBLKCP" in sass


def test_no_cpu_fallback_without_cuda():
    import pytest
    import torch

    from slate_quantum_b200 import _lib, kernels

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(_lib.SlateB200Error):
        kernels.bloch_kinetic((5,), (1,), [[1.0]], 1.0, 1.0)
