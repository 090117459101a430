import os, sys, time
sys.path.insert(0, '/root/repo')
import numpy as np, torch, torch.distributed as td
from gsimcli_b200 import _capi, synth, dist
from gsimcli_b200.tools import grid as gr, homog as hmg
world = int(os.environ.get('WORLD_SIZE', 1)); rank = int(os.environ.get('RANK', 0)); local = int(os.environ.get('LOCAL_RANK', 0))
torch.cuda.set_device(local)
td.init_process_group('nccl', device_id=torch.device('cuda', local))
R = 1000; dims = [200, 200, 100 * world]; FIRST = [0.0, 0.0, 1900]; CELLS = [1000.0, 1000.0, 1]; ND = -999.9
QS = [round(0.05 * k, 2) for k in range(1, 20)]
n0, n1 = dist.slab_range(dims, rank, world); nn = n1 - n0
grids = gr.GridFiles.from_synthetic(1003, R, dims, FIRST, CELLS, ND, device=local, node_range=(n0, n1))
stack = grids._stack
flags = (_capi.MEAN | _capi.MEDIAN | _capi.SKEW | _capi.VAR | _capi.STD | _capi.COEFVAR | _capi.MODE)
npl = _capi.lib().gsb_stats_grid_planes(flags, len(QS))
d_out = torch.empty((npl, nn), dtype=torch.float32, device='cuda')
stations = synth.stations(1003, 20, R, dims, FIRST, CELLS)
cols = ['x', 'y', 'time', 'station', 'clim']
def obs_list():
    return [gr.PointSet('cand%d' % i, ND, 5, list(cols), np.array(s['obs'], dtype=np.float64)) for i, s in enumerate(stations)]
stack.set_option('k2_sm_reserve', 4)
s_grid = torch.cuda.Stream(); s_stat = torch.cuda.Stream()
T = {}
def wrap(obj, name, label):
    f = getattr(obj, name)
    def g(*a, **k):
        t0 = time.perf_counter(); r = f(*a, **k); T.setdefault(label, []).append((time.perf_counter() - t0) * 1e3); return r
    setattr(obj, name, g)
wrap(_capi.Stack, 'stats_columns', 'stats_columns(K2c+D2H)')
if os.environ.get('GLOO'):
    _g = td.new_group(backend='gloo')
    _orig = dist.allgather_tables
    dist.allgather_tables = lambda local, dims, group=None, device=None: _orig(local, dims, group=_g, device=device)
wrap(dist, 'allgather_tables', 'allgather_tables')
wrap(_capi, 'detect', 'detect(K3)')
wrap(td, 'all_gather', ' td.all_gather')
def k2(): stack.stats_grid_device(flags, 0.95, QS, ND, 0, nn, d_out.data_ptr(), nn, s_grid.cuda_stream)
def det(): return hmg.detect_many(grids, obs_list(), method='median', prob=0.95, stream=s_stat.cuda_stream)
for mode in ('detect only', 'both'):
    for it in range(6):
        if it == 2: T.clear()
        t0 = time.perf_counter()
        if mode == 'both': k2()
        det(); t1 = time.perf_counter(); torch.cuda.synchronize()
        T.setdefault('det total', []).append((t1 - t0) * 1e3); T.setdefault('step', []).append((time.perf_counter() - t0) * 1e3)
    if True:
        print('rank', rank, mode, {k: round(float(np.mean(v)), 2) for k, v in T.items()}, flush=True)
td.destroy_process_group()
