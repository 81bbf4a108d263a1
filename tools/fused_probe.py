import sys
from pathlib import Path
import numpy as np, torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from slate_quantum_b200 import kernels
shape, repeat, lead = (16, 16), (64, 64), 512
m = 256 * 4096
x = torch.randn((lead, m), dtype=torch.complex128, device="cuda")
sp, ln = [2 * np.pi / 16] * 2, [2 * np.pi * 64] * 2
for _ in range(2):
    kernels.cell_fft_moments(shape, repeat, x, sp, ln, stage=1)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record(); kernels.cell_fft_moments(shape, repeat, x, sp, ln, stage=1); b.record(); torch.cuda.synchronize()
print("fused last pass", a.elapsed_time(b), "ms for", x.numel() * 16 / 1e9, "GB ->", x.numel() * 16 / a.elapsed_time(b) / 1e6, "GB/s")
out = torch.empty_like(x)
for rep in range(4):
    a.record(); kernels.cell_fft(shape, repeat, x, 1, out=out); b.record(); torch.cuda.synchronize()
    print("plain cell_fft (2 passes)", a.elapsed_time(b), "ms")
for rep in range(3):
    a.record(); kernels.cell_fft_moments(shape, repeat, x, sp, ln, stage=1); b.record(); torch.cuda.synchronize()
    print("fused last pass", a.elapsed_time(b), "ms")
