import sys, time, ctypes
sys.path.insert(0,'/root/repo')
import numpy as np, torch
from gsimcli_b200 import _capi
R=1000; dims=[200,200,100]; nn=4000000; ND=-999.9
QS=[round(0.05*k,2) for k in range(1,20)]
stack=_capi.Stack(R,nn); stack.synth_fill(1003,dims,0,ND)
flags=(_capi.MEAN|_capi.MEDIAN|_capi.SKEW|_capi.VAR|_capi.STD|_capi.COEFVAR|_capi.MODE)
npl=_capi.lib().gsb_stats_grid_planes(flags,len(QS))
pitch=stack.pitch
host=torch.empty((R,pitch),dtype=torch.float32,pin_memory=True)
_capi.check(_capi.lib().gsb_stack_download_f32(stack.handle,0,R,0,nn,ctypes.c_void_p(host.data_ptr()),pitch,None))
maps=torch.empty((npl,nn),dtype=torch.float32,pin_memory=True)
def T(f,n=3):
    f(); torch.cuda.synchronize(); t0=time.perf_counter()
    for _ in range(n): f()
    torch.cuda.synchronize(); return (time.perf_counter()-t0)/n*1e3
print('full upload %.1f ms'%T(lambda: stack.upload_ptr(host.data_ptr(),R,nn,pitch)))
def slabs(k):
    step=(-(-nn//k)+31)//32*32
    for a in range(0,nn,step):
        b=min(nn,a+step); stack.upload_ptr(host.data_ptr()+4*a,R,b-a,pitch,n0=a)
for k in (2,8,32): print('slab uploads x%d %.1f ms'%(k,T(lambda: slabs(k))))
for k in (2,4,8,16): print('pipelined nslabs=%d %.1f ms'%(k,T(lambda: (stack.stats_grid_from_host(host.data_ptr(),pitch,flags,0.95,QS,ND,maps,nslabs=k), torch.cuda.current_stream().synchronize()))))
