"""Scratch timing of every kernel at a chosen config (CUDA events). Not the bench; a dev probe."""
import argparse, json, sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, torch
from slate_quantum_b200 import kernels, _lib

def timeit(fn, reps=3, warm=1):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    ts=[]
    for _ in range(reps):
        a=torch.cuda.Event(enable_timing=True); b=torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return min(ts)

ap=argparse.ArgumentParser()
ap.add_argument("--shape", default="16,16"); ap.add_argument("--repeat", default="64,64")
ap.add_argument("--nk", type=int, default=0); ap.add_argument("--T", type=int, default=4096)
ap.add_argument("--peaks", action="store_true")
a=ap.parse_args()
shape=tuple(int(x) for x in a.shape.split(",")); repeat=tuple(int(x) for x in a.repeat.split(","))
nd=len(shape); n=int(np.prod(shape)); nk=int(np.prod(repeat)); 
cnt = a.nk or nk
hb=1.054571817e-34
res={}
if a.peaks:
    lib=_lib.load(); import ctypes
    sink=torch.zeros(8,dtype=torch.float64,device="cuda"); fl=ctypes.c_double(0)
    for name,fn,lead in (("dmma0",lib.sqb_peak_dmma,(0,)),("dmma1",lib.sqb_peak_dmma,(1,)),("dmma2",lib.sqb_peak_dmma,(2,)),("dfma",lib.sqb_peak_dfma,())):
        it=20000
        def run(): fn(*lead, it, ctypes.c_void_p(sink.data_ptr()), ctypes.byref(fl), ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
        ms=timeit(run,reps=5,warm=2)
        res[f"peak_{name}_tflops"]=fl.value/ms/1e9
dk=2*np.pi*np.linalg.inv(2*np.pi*np.eye(nd)).T
# cos potential comps
comp=np.ones((3,)*nd,dtype=complex)*0.5
tab=np.zeros(shape,dtype=complex); idx=[np.array([0,1,-1])%s for s in shape]; tab[np.ix_(*idx)]=comp*np.sqrt(n)
tabd=kernels.to_device(tab.ravel(),torch.complex128)
H=torch.empty((cnt,n,n),dtype=torch.complex128,device="cuda")
ms=timeit(lambda: kernels.bloch_build(shape,repeat,dk,tabd,hb**2,hb,0,cnt,out=H))
res["build_ms"]=ms; res["build_GBs"]=cnt*n*n*16/ms/1e6
lib=_lib.load()
wsb=lib.sqb_eigh_workspace_bytes(n,cnt); ws=torch.empty(wsb,dtype=torch.uint8,device="cuda")
res["eigh_ws_GB"]=wsb/1e9
def eig():
    kernels.bloch_build(shape,repeat,dk,tabd,hb**2,hb,0,cnt,out=H)
    return kernels.eigh_batched(H,workspace=ws)
ms2=timeit(eig,reps=2)
res["build+eigh_ms"]=ms2; res["eigh_ms"]=ms2-ms; res["blocks_per_s"]=cnt/ms2*1e3
res["eigh_tflops_68n3/3"]=cnt*(68/3)*n**3/((ms2-ms)*1e9)
w,V,info=eig(); res["info_bad"]=int((info!=0).sum())
psi=torch.randn(cnt,n,dtype=torch.complex128,device="cuda")
res["project_ms"]=timeit(lambda: kernels.project(V,psi))
c=kernels.project(V,psi)
T=a.T
times=torch.linspace(0,8*np.pi*hb,T,dtype=torch.float64,device="cuda")
# limit output memory: tile T
free,_=torch.cuda.mem_get_info()
tt=T
while tt*cnt*n*16 > free*0.7: tt//=2
out=torch.empty((tt,cnt*n),dtype=torch.complex128,device="cuda")
def ev():
    for t0 in range(0,T,tt):
        kernels.evolve_bloch(V,w,c,times[t0:t0+tt].contiguous(),hb,out=out)
ms3=timeit(ev,reps=2)
res["evolve_ms"]=ms3; res["evolve_tflops"]=8.0*n*n*T*cnt/(ms3*1e9); res["evolve_t_tile"]=tt
res["evolutions_per_s"]=T/ms3*1e3
# fft + permute of a time-tile
big=tuple(s*r for s,r in zip(shape,repeat))
if cnt==nk:
    tf=min(tt,64)
    sub=out[:tf]
    res["permute_ms_per_time"]=timeit(lambda: kernels.bloch_permute(shape,repeat,sub,1))/tf
    tmp=kernels.bloch_permute(shape,repeat,sub,1)
    res["fft_ms_per_time"]=timeit(lambda: kernels.fft_c2c(tmp,big,+1,out=tmp))/tf
    res["fft_GBs(32B/elem/axis)"]=tf*np.prod(big)*32*nd/(res["fft_ms_per_time"]*tf)/1e6
print(json.dumps(res,indent=1))
