"""Dev probe: throughput of sqb_zgemm_batched at the shapes the operator algebra produces (CUDA events)."""
import json, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from slate_quantum_b200 import kernels

def timeit(fn, reps=5, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return min(ts)

only = sys.argv[sys.argv.index("--only") + 1] if "--only" in sys.argv else None
res = {}
g = torch.Generator(device="cuda").manual_seed(0)
def rnd(*s):
    return torch.randn(*s, dtype=torch.complex128, device="cuda", generator=g)
for name, (batch, m, n, k, tb) in {
    "op_2048": (None, 2048, 2048, 2048, False),
    "op_4096": (None, 4096, 4096, 4096, False),
    "op_4096_Bt": (None, 4096, 4096, 4096, True),   # expectation_of_each: Psi @ O^T
    "list_64x512": (64, 512, 512, 512, False),
    "blocks_4096x256": (4096, 256, 256, 256, False),
    "states_4096xM4096": (None, 4096, 4096, 4096, True),
}.items():
    if only and name != only:
        continue
    lead = () if batch is None else (batch,)
    a = rnd(*lead, m, k)
    b = rnd(*lead, n, k).mT if tb else rnd(*lead, k, n)
    out = torch.empty(*lead, m, n, dtype=torch.complex128, device="cuda")
    ms = timeit(lambda: kernels.zgemm(a, b, out=out))
    flop = 8.0 * m * n * k * (batch or 1)
    res[name] = {"ms": ms, "tflops": flop / ms / 1e9}
    del a, b, out
print(json.dumps(res, indent=1))
