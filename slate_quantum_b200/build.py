"""Build libslate_b200.so in-tree with nvcc for sm_100a (no JIT cache: the .so travels with the repo)."""

from __future__ import annotations

import os
import subprocess
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
CSRC = HERE / "csrc"
LIB = HERE / "libslate_b200.so"
SOURCES = ["abi.cu", "bloch_build.cu", "fft.cu", "eigh.cu", "evolve.cu", "measure.cu", "zgemm.cu", "peaks.cu"]
NVCC_FLAGS = [
    "-gencode",
    "arch=compute_100a,code=sm_100a",
    "-O3",
    "-std=c++17",
    "-lineinfo",
    "-Xcompiler",
    "-fPIC",
    "-Xptxas=-v",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (Path(cand).exists() or cand == "nvcc"):
            return cand
    return "nvcc"


def needs_build() -> bool:
    if not LIB.exists():
        return True
    mtime = LIB.stat().st_mtime
    deps = [CSRC / s for s in SOURCES] + sorted(CSRC.glob("*.cuh")) + [HERE.parent / "include" / "slate_b200.h"]
    return any(d.stat().st_mtime > mtime for d in deps)


def build(force: bool = False, verbose: bool = False) -> Path:
    """Compile every .cu for sm_100a into one shared library. Raises on failure."""
    if not force and not needs_build():
        return LIB
    objs = []
    procs = []
    build_dir = HERE / "csrc" / "build"
    build_dir.mkdir(exist_ok=True)
    for src in SOURCES:
        obj = build_dir / (src.replace(".cu", ".o"))
        extra = os.environ.get("SQB_NVCC_EXTRA", "").split()  # developer switches, e.g. -DSQB_PANEL_PROFILE
        cmd = [_nvcc(), *NVCC_FLAGS, *extra, "-c", str(CSRC / src), "-o", str(obj)]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(str(obj))
    log = []
    for src, p in procs:
        out, _ = p.communicate()
        log.append(f"== {src}\n{out}")
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{out}")
    # --cudart=shared: the process already holds torch's libcudart.so.12; a second, statically linked runtime would
    # only duplicate state (and drags the whole runtime's symbol table into the shipped artefact)
    cmd = [_nvcc(), "-shared", "--cudart=shared", "-o", str(LIB), *objs, "-gencode", "arch=compute_100a,code=sm_100a"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    (build_dir / "ptxas.log").write_text("\n".join(log))
    if verbose:
        print("\n".join(log))
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
