"""Dev probe: DMMA / DFMA micro-benchmark peaks in different GPU states (cold, after allocations, after load)."""
import ctypes, subprocess, sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from slate_quantum_b200 import _lib
lib = _lib.load()
sink = torch.zeros(8, dtype=torch.float64, device="cuda"); fl = ctypes.c_double(0)
def clk():
    return subprocess.run(["nvidia-smi","--query-gpu=clocks.sm,power.draw,clocks_event_reasons.active","--format=csv,noheader"],capture_output=True,text=True).stdout.strip()
def peak(fn, *lead, iters=20000, reps=4):
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream); best=0; vals=[]
    for _ in range(reps):
        a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        a.record(); fn(*lead, iters, ctypes.c_void_p(sink.data_ptr()), ctypes.byref(fl), st); b.record(); torch.cuda.synchronize()
        vals.append(round(fl.value/a.elapsed_time(b)/1e9,2))
    return vals
print("cold dmma v0,v1,v2", [peak(lib.sqb_peak_dmma, v) for v in range(3)], clk())
print("cold dfma", peak(lib.sqb_peak_dfma), clk())
x = torch.empty(60*(1<<30), dtype=torch.uint8, device="cuda")
print("after 60GB alloc dmma", [peak(lib.sqb_peak_dmma, v) for v in range(3)], clk())
a = torch.randn(8192,8192,device="cuda",dtype=torch.float64)
for _ in range(5): b = a@a
torch.cuda.synchronize()
print("after fp64 gemm load dmma", [peak(lib.sqb_peak_dmma, v) for v in range(3)], clk())
print("long dmma (200k iters)", peak(lib.sqb_peak_dmma, 0, iters=200000, reps=3), clk())
time.sleep(2)
print("after 2s idle dmma", [peak(lib.sqb_peak_dmma, v) for v in range(3)], clk())
y = torch.empty(1<<30, dtype=torch.uint8).pin_memory()
print("after 1GB pinned dmma", [peak(lib.sqb_peak_dmma, v) for v in range(3)], clk())
