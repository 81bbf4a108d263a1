"""Time sqb_eigh_batched on random Hermitian blocks (GPU).  python tools/eigh_probe.py [n] [batch] [reps]"""
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from slate_quantum_b200 import kernels  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
g = torch.Generator(device="cuda").manual_seed(0)
a = torch.randn((batch, n, n), dtype=torch.float64, device="cuda", generator=g) + 1j * torch.randn(
    (batch, n, n), dtype=torch.float64, device="cuda", generator=g)
h = (a + a.mH).contiguous()
del a
ws = torch.empty(kernels._lib.load().sqb_eigh_workspace_bytes(n, batch), dtype=torch.uint8, device="cuda")
best = 1e9
for rep in range(reps):
    x = h.clone()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    w, v, info = kernels.eigh_batched(x, workspace=ws)
    e1.record()
    torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1))
assert int(info.abs().sum()) == 0
hv = h[:4] @ v[:4].mT
res = (hv - v[:4].mT * w[:4, None, :]).abs().max().item() / h[:4].abs().max().item()
print(f"n={n} batch={batch} eigh best {best:.2f} ms  ({(68 / 3) * n**3 * batch / best / 1e9:.2f} TFLOP/s)  residual {res:.1e}")
