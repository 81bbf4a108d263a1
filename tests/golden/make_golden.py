"""Generate tests/golden/bloch_golden.npz from the CPU oracle (committed together with its output).

The reference itself cannot run in the build container (needs Python >= 3.14 and the un-vendored
slate_core), so these vectors are ORACLE outputs on small seeded cases, not reference outputs.  The
reference's own literal known-answer vectors are in tests/golden/reference_literals.json (hand-transcribed
with file:line) and tests/test_oracle_golden.py checks the oracle against both.

    python tests/golden/make_golden.py
"""

import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle import bloch_oracle as o  # noqa: E402

HB = o.HBAR_SI
out = {}
cases = {
    "sq1d": dict(dx=2 * np.pi * np.eye(1), shape=(6,), repeat=(4,), comps=o.cos_potential_components((6,), 3.0)),
    "sq2d": dict(dx=2 * np.pi * np.eye(2), shape=(4, 3), repeat=(3, 2), comps=o.sin_potential_components((4, 3), 2.0, 0.25)),
    "hex2d": dict(
        dx=2 * np.pi * np.array([[np.sin(np.pi / 3), np.cos(np.pi / 3)], [0.0, 1.0]]),
        shape=(4, 4), repeat=(2, 3), comps=o.fcc_potential_components((4, 4), 5.0),
    ),
    "cub3d": dict(dx=2 * np.pi * np.eye(3), shape=(3, 3, 3), repeat=(2, 2, 2), comps=o.cos_potential_components((3, 3, 3), 6.0)),
}
rng = np.random.default_rng(20261019)
for name, c in cases.items():
    blocks = o.bloch_blocks(c["dx"], c["shape"], c["comps"], HB**2, c["repeat"])
    w, t = o.eigh_blocks(blocks)
    m = int(np.prod(c["shape"]) * np.prod(c["repeat"]))
    psi = rng.normal(size=m) + 1j * rng.normal(size=m)
    psi /= np.linalg.norm(psi)
    times = np.linspace(0.0, 2.5 * HB, 5)
    res = o.evolve_pipeline(psi, c["shape"], c["repeat"], w, t, times, HB)
    out[f"{name}/dx"] = c["dx"]
    out[f"{name}/shape"] = np.array(c["shape"])
    out[f"{name}/repeat"] = np.array(c["repeat"])
    out[f"{name}/comps"] = c["comps"]
    out[f"{name}/blocks"] = blocks
    out[f"{name}/eigenvalues"] = w
    out[f"{name}/psi0"] = psi
    out[f"{name}/times"] = times
    out[f"{name}/eigen_amplitudes"] = res["eigen"]
    out[f"{name}/position"] = res["position"]
    out[f"{name}/permutation"] = o.bloch_permutation(c["shape"], c["repeat"])
np.savez_compressed(Path(__file__).with_name("bloch_golden.npz"), **out)
print("wrote", Path(__file__).with_name("bloch_golden.npz"))
