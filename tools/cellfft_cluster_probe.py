"""Dev probe: sqb_cell_fft on the cfg3 block grid (64 x 64 blocks of 16 x 16), a few hundred time rows."""
import json, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from slate_quantum_b200 import kernels

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 256
shape, repeat = (16, 16), (64, 64)
m = 256 * 4096
x = torch.randn(rows, m, dtype=torch.complex128, device="cuda")
out = torch.empty_like(x)
best = 1e9
for _ in range(4):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); kernels.cell_fft(shape, repeat, x, 1, out=out); b.record(); torch.cuda.synchronize()
    best = min(best, a.elapsed_time(b))
print(json.dumps({"rows": rows, "ms": best, "ms_per_4096_rows": best * 4096 / rows,
                  "TB/s_of_one_read_plus_one_write": 2 * rows * m * 16 / best / 1e9}))
