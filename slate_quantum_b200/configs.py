"""The five BASELINE.json configs as concrete synthetic workloads.

Pure host-side descriptions (numpy only); shared by bench.py, the tests and the examples.
"""

from __future__ import annotations

from dataclasses import dataclass

import numpy as np

HBAR = 1.054571817e-34  # scipy.constants.hbar, the reference's default (dynamics/schrodinger.py:4)


@dataclass(frozen=True)
class BlochWorkload:
    name: str
    delta_x: tuple[tuple[float, ...], ...]  # full unit-cell vectors, row i = a_i
    shape: tuple[int, ...]                  # unit-cell grid
    repeat: tuple[int, ...]                 # k-points per axis
    potential: str                          # "cos" | "fcc"
    height: float
    n_times: int
    t_max_over_hbar: float                  # times = linspace(0, t_max, n_times, endpoint=False)
    mass_over_hbar2: float = 1.0            # mass = hbar**2 as in the reference tests

    @property
    def nd(self) -> int:
        return len(self.shape)

    @property
    def n(self) -> int:
        return int(np.prod(self.shape))

    @property
    def n_k(self) -> int:
        return int(np.prod(self.repeat))

    @property
    def supercell(self) -> tuple[int, ...]:
        return tuple(n * r for n, r in zip(self.shape, self.repeat, strict=True))

    @property
    def m(self) -> int:
        return self.n * self.n_k

    def vectors(self) -> np.ndarray:
        return np.asarray(self.delta_x, dtype=np.float64)

    def dk(self) -> np.ndarray:
        """Reciprocal vectors, row i = dk_i, dk_i . a_j = 2 pi delta_ij."""
        return 2 * np.pi * np.linalg.inv(self.vectors()).T

    def mass(self) -> float:
        return self.mass_over_hbar2 * HBAR**2

    def times(self) -> np.ndarray:
        return np.linspace(0.0, self.t_max_over_hbar * HBAR, self.n_times, endpoint=False)

    def potential_components(self) -> np.ndarray:
        """Cropped Fourier components as the reference's builders produce them
        (operator/_build/_potential.py:86-108 cos, :167-190 fcc)."""
        nd = self.nd
        if self.potential == "cos":
            axis = np.array([2.0, 1.0, 1.0], dtype=np.complex128)
            grids = np.meshgrid(*(axis,) * nd, indexing="ij")
            return 0.25**nd * self.height * np.prod(grids, axis=0) * np.sqrt(self.n)
        if self.potential == "fcc":
            data = np.array([[3, 1, 1], [1, 1, 0], [1, 0, 1]]).astype(np.complex128)
            return (1 / 3) ** 2 * data * self.height * np.sqrt(self.n)
        raise ValueError(self.potential)

    def potential_table(self) -> np.ndarray:
        """Components zero-padded to `shape` in FFT order (what sqb_bloch_build takes)."""
        comps = self.potential_components()
        full = np.zeros(self.shape, dtype=np.complex128)
        idx = [np.array([0, 1, -1][: c]) % n for c, n in zip(comps.shape, self.shape, strict=True)]
        full[np.ix_(*idx)] = comps
        return full

    def initial_state(self) -> np.ndarray:
        """Gaussian wavepacket centred in the supercell, one unit cell wide, position basis, normalised."""
        grids = np.meshgrid(*(np.arange(s) / s - 0.5 for s in self.supercell), indexing="ij")
        arg = sum((g * r) ** 2 for g, r in zip(grids, self.repeat, strict=True))
        psi = np.exp(-0.5 * arg).astype(np.complex128).ravel()
        return psi / np.linalg.norm(psi)


_TWO_PI = 2 * np.pi
_HEX = ((_TWO_PI * np.sin(np.pi / 3), _TWO_PI * np.cos(np.pi / 3)), (0.0, _TWO_PI))

CONFIGS: dict[str, BlochWorkload] = {
    # configs[0]: examples/blochs_theorem.py-style 1D cosine, 3 Fourier comps, 100 k-points, 64 states/block
    "cfg1": BlochWorkload("cfg1-1d-cos-64x100", ((_TWO_PI,),), (64,), (100,), "cos", 1.0, 64, 8 * np.pi),
    # configs[1]: 1D periodic potential eigenstates + evolution: 1024 k-points, N=128, 1024 times
    "cfg2": BlochWorkload("cfg2-1d-cos-128x1024-T1024", ((_TWO_PI,),), (128,), (1024,), "cos", 10.0, 1024, 8 * np.pi),
    # configs[2]: 2D square lattice 64x64 k-points, N=256, eigh + 4096-time evolution  (north_star target)
    "cfg3": BlochWorkload(
        "cfg3-2d-square-16x16-64x64k-T4096", ((_TWO_PI, 0.0), (0.0, _TWO_PI)), (16, 16), (64, 64), "cos", 10.0, 4096,
        8 * np.pi,
    ),
    # configs[3]: 2D hexagonal 128x128 k-points, N=512 (large blocks; eigh focus)
    "cfg4": BlochWorkload("cfg4-2d-hex-32x16-128x128k", _HEX, (32, 16), (128, 128), "fcc", 10.0, 64, 8 * np.pi),
    # configs[4]: 3D cubic 32^3 k-points, N=343, sharded over 8 GPUs
    "cfg5": BlochWorkload(
        "cfg5-3d-cubic-7x7x7-32x32x32k-T512",
        ((_TWO_PI, 0.0, 0.0), (0.0, _TWO_PI, 0.0), (0.0, 0.0, _TWO_PI)),
        (7, 7, 7), (32, 32, 32), "cos", 10.0, 512, 8 * np.pi,
    ),
}


def shard_blocks(n_k: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous block range [begin, begin+count) owned by `rank`."""
    base, rem = divmod(n_k, world)
    begin = rank * base + min(rank, rem)
    return begin, base + (1 if rank < rem else 0)
