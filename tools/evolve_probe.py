"""Time the uniform-grid evolve routes at a config's block shape (random eigensystems, CUDA events).

    python tools/evolve_probe.py [n] [batch] [rows]          (SQB_EVOLVE=gemm|nufft selects the route)
"""
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from slate_quantum_b200 import kernels  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
rows = int(sys.argv[3]) if len(sys.argv) > 3 else 2048
hbar = 1.0545718e-34
torch.manual_seed(0)
v = torch.randn(batch, n, n, dtype=torch.complex128, device="cuda") / np.sqrt(n)
w = torch.sort(torch.rand(batch, n, dtype=torch.float64, device="cuda") * 40.0 * hbar * 1e12, dim=1).values
c = torch.randn(batch, n, dtype=torch.complex128, device="cuda") / np.sqrt(n)
times = torch.arange(rows, dtype=torch.float64, device="cuda") * 0.37e-12
out = torch.empty(rows, batch * n, dtype=torch.complex128, device="cuda")
ws = torch.empty(kernels.evolve_uniform_workspace_bytes(n, batch, rows), dtype=torch.uint8, device="cuda")
route = kernels.evolve_uniform_route(n, batch, rows)
for i in range(2):
    kernels.evolve_bloch(v, w, c, times, hbar, out=out, uniform_dt=0.37e-12, workspace=ws, reuse_plane=i > 0)
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
reps = 3
ev[0].record()
for _ in range(reps):
    kernels.evolve_bloch(v, w, c, times, hbar, out=out, uniform_dt=0.37e-12, workspace=ws, reuse_plane=True)
ev[1].record()
torch.cuda.synchronize()
ms = ev[0].elapsed_time(ev[1]) / reps
# spot check of a few entries against the direct sum
rr = [0, rows // 2 - 1, rows - 1]
bb = [0, batch // 2, batch - 1]
err = 0.0
for b in bb:
    ph = torch.exp(-1j * w[b][None, :] * times[rr][:, None] / hbar) * c[b][None, :]
    err = max(err, float((ph @ v[b] - out[rr][:, b * n : (b + 1) * n]).abs().max()))
flops = 8.0 * n * n * rows * batch
print(f"route {route}: n {n} batch {batch} rows {rows}: {ms:.2f} ms per call = {flops / ms / 1e9:.1f} TFLOP/s of 8 N^2 T, "
      f"{rows * batch * n * 16 / ms / 1e6:.0f} GB/s of output; max err vs direct sum {err:.2e}")
