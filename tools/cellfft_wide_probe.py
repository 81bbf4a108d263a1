"""Dev probe: the position stage (sqb_cell_fft) of the 8-GPU weak-scaling k-grid (512 x 64 blocks of 16 x 16) on one
GPU, for a few time rows.  SQB_FFT_INPLACE=0 gives the two-buffer FFT pass for comparison."""
import json, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from slate_quantum_b200 import kernels

rows = 32
res = {}
for name, repeat in (("n8_512x64", (512, 64)), ("n8_256x128", (256, 128)), ("n4_256x64", (256, 64)), ("n4_128x128", (128, 128)),
                     ("n2_128x64", (128, 64)), ("n1_64x64", (64, 64))):
    shape = (16, 16)
    m = 256 * repeat[0] * repeat[1]
    t = rows if repeat[0] > 64 else rows * 8
    x = torch.randn(t, m, dtype=torch.complex128, device="cuda")
    out = torch.empty_like(x)
    best = 1e9
    for _ in range(4):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); kernels.cell_fft(shape, repeat, x, 1, out=out); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    gb = 4 * t * m * 16 / 1e9  # two passes, read + write each
    res[name] = {"ms": best, "TB/s": gb / best, "ms_per_full_rank_tile": best * (68.7e9 / (t * m * 16))}
    # parity of one row against torch's FFT of the same cell layout (ifft over the two block axes, scattered)
    ref = torch.fft.ifftn(x[0].reshape(repeat[0], repeat[1], 16, 16), dim=(0, 1), norm="ortho").permute(0, 2, 1, 3).reshape(-1) if False else None
    del x, out
print(json.dumps(res))
