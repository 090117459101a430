"""Per-component step timing under torchrun (node-sharded stack)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as td
from gsimcli_b200 import _capi, synth, dist
from gsimcli_b200.tools import grid as gr, homog as hmg
world = int(os.environ.get('WORLD_SIZE', 1)); rank = int(os.environ.get('RANK', 0)); local = int(os.environ.get('LOCAL_RANK', 0))
torch.cuda.set_device(local)
if world > 1:
    td.init_process_group('nccl', device_id=torch.device('cuda', local))
R = 1000; dims = [200, 200, 100 * world]; FIRST = [0.0, 0.0, 1900]; CELLS = [1000.0, 1000.0, 1]; ND = -999.9
QS = [round(0.05 * k, 2) for k in range(1, 20)]
n0, n1 = dist.slab_range(dims, rank, world); nn = n1 - n0
grids = gr.GridFiles.from_synthetic(1003, R, dims, FIRST, CELLS, ND, device=local, node_range=(n0, n1) if world > 1 else None)
stack = grids._stack
flags = (_capi.MEAN | _capi.MEDIAN | _capi.SKEW | _capi.VAR | _capi.STD | _capi.COEFVAR | _capi.MODE)
npl = _capi.lib().gsb_stats_grid_planes(flags, len(QS))
d_out = torch.empty((npl, nn), dtype=torch.float32, device='cuda')
stations = synth.stations(1003, 20, R, dims, FIRST, CELLS)
cols = ['x', 'y', 'time', 'station', 'clim']
def obs_list():
    return [gr.PointSet('cand%d' % i, ND, 5, list(cols), np.array(s['obs'], dtype=np.float64)) for i, s in enumerate(stations)]
stack.set_option('k2_sm_reserve', int(os.environ.get('RES', '4')))
s_grid = torch.cuda.Stream(); s_stat = torch.cuda.Stream()
def k2(): stack.stats_grid_device(flags, 0.95, QS, ND, 0, nn, d_out.data_ptr(), nn, s_grid.cuda_stream)
def det(): return hmg.detect_many(grids, obs_list(), method='median', prob=0.95, stream=s_stat.cuda_stream)
for name, fn in [('k2 only', lambda: (k2(), torch.cuda.synchronize())), ('detect only', lambda: (det(), torch.cuda.synchronize())), ('both', lambda: (k2(), det(), torch.cuda.synchronize()))]:
    for _ in range(2): fn()
    if world > 1: td.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5): fn()
    dt = (time.perf_counter() - t0) / 5 * 1e3
    if rank == 0: print(name, '%.2f ms' % dt, flush=True)
if os.environ.get('PROF') and rank == 0:
    import cProfile, pstats, io
    pr = cProfile.Profile(); pr.enable()
    for _ in range(5): det()
    pr.disable(); s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats('cumulative').print_stats(25); print(s.getvalue()[:5000])
elif os.environ.get('PROF'):
    for _ in range(5): det()
if world > 1: td.destroy_process_group()
