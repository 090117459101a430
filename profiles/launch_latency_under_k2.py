import os, sys, time
sys.path.insert(0, '/root/repo')
import numpy as np, torch
from gsimcli_b200 import _capi
R=1000; dims=[200,200,100]; nn=4000000; ND=-999.9
QS=[round(0.05*k,2) for k in range(1,20)]
stack=_capi.Stack(R,nn); stack.synth_fill(1003,dims,0,ND)
flags=(_capi.MEAN|_capi.MEDIAN|_capi.SKEW|_capi.VAR|_capi.STD|_capi.COEFVAR|_capi.MODE)
npl=_capi.lib().gsb_stats_grid_planes(flags,len(QS))
d_out=torch.empty((npl,nn),dtype=torch.float32,device='cuda')
for res in (4, 0, 16):
    stack.set_option('k2_sm_reserve', res)
    s_grid=torch.cuda.Stream(); s2=torch.cuda.Stream()
    x=torch.zeros(1024,device='cuda'); h=torch.zeros(1024).pin_memory()
    def k2(): stack.stats_grid_device(flags,0.95,QS,ND,0,nn,d_out.data_ptr(),nn,s_grid.cuda_stream)
    k2(); torch.cuda.synchronize()
    for label, run_k2 in (('idle',False),('under K2',True)):
        if run_k2: k2()
        lat=[]
        with torch.cuda.stream(s2):
            for i in range(8):
                t0=time.perf_counter(); x.add_(1.0); h.copy_(x, non_blocking=True); s2.synchronize(); lat.append((time.perf_counter()-t0)*1e3)
        torch.cuda.synchronize()
        print('reserve',res,label,'small kernel+D2H latency ms:',[round(v,3) for v in lat])
