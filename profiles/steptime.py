import os, sys, time
sys.path.insert(0, '/root/repo')
import numpy as np, pandas as pd, torch
from gsimcli_b200 import _capi, synth
from gsimcli_b200.tools import grid as gr, homog as hmg
R=1000; dims=[200,200,100]; FIRST=[0.0,0.0,1900]; CELLS=[1000.0,1000.0,1]; ND=-999.9
QS=[round(0.05*k,2) for k in range(1,20)]
grids=gr.GridFiles.from_synthetic(1003,R,dims,FIRST,CELLS,ND,device=0)
stack=grids._stack
nn=int(np.prod(dims))
flags=(_capi.MEAN|_capi.MEDIAN|_capi.SKEW|_capi.VAR|_capi.STD|_capi.COEFVAR|_capi.MODE)
npl=_capi.lib().gsb_stats_grid_planes(flags,len(QS))
d_out=torch.empty((npl,nn),dtype=torch.float32,device='cuda')
stations=synth.stations(1003,20,R,dims,FIRST,CELLS)
cols=['x','y','time','station','clim']
def obs_list():
    return [gr.PointSet('cand%d'%i,ND,5,list(cols),pd.DataFrame(s['obs'],columns=cols)) for i,s in enumerate(stations)]
stack.set_option('k2_sm_reserve', int(os.environ.get('RES','4')))
s_grid=torch.cuda.Stream(); s_stat=torch.cuda.Stream()
def k2():
    stack.stats_grid_device(flags,0.95,QS,ND,0,nn,d_out.data_ptr(),nn,s_grid.cuda_stream)
def det():
    return hmg.detect_many(grids,obs_list(),method='median',prob=0.95,stream=s_stat.cuda_stream)
for name,fn in [('k2 only',lambda:(k2(),torch.cuda.synchronize())),('detect only',lambda:(det(),torch.cuda.synchronize())),('both',lambda:(k2(),det(),torch.cuda.synchronize()))]:
    for _ in range(2): fn()
    t0=time.perf_counter()
    for _ in range(5): fn()
    print(name, (time.perf_counter()-t0)/5*1e3,'ms')
