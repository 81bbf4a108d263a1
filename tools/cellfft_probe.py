"""Dev probe: cell FFT (cell layout -> position) bandwidth for several supercell shapes (CUDA events + ncu-friendly)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, torch
from slate_quantum_b200 import kernels

def timeit(fn, reps=3):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return min(ts)

shape = (16, 16)
for repeat in [(64, 64), (128, 64), (256, 64), (512, 64)]:
    m = int(np.prod(shape)) * int(np.prod(repeat))
    lead = max(1, (8 << 30) // (m * 16))
    x = torch.randn(lead, m, dtype=torch.complex128, device="cuda")
    out = torch.empty_like(x)
    ms = timeit(lambda: kernels.cell_fft(shape, repeat, x, 1, out=out))
    gb = lead * m * 16 * 4 / 1e9
    print(f"repeat {repeat} lead {lead}: {ms:.2f} ms, {gb / ms:.2f} TB/s of 4 x array bytes")
