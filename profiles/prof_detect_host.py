"""cProfile of the host glue of detect_many (20 stations, config-3 stack)."""
import cProfile, pstats, sys, os, io
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, pandas as pd, torch
from gsimcli_b200 import _capi, synth
from gsimcli_b200.tools import grid as gr, homog as hmg
R = 200; dims = [100, 100, 60]; FIRST = [0.0, 0.0, 1900]; CELLS = [1000.0, 1000.0, 1]; ND = -999.9
grids = gr.GridFiles.from_synthetic(1003, R, dims, FIRST, CELLS, ND, device=0)
stations = synth.stations(1003, 20, R, dims, FIRST, CELLS)
cols = ['x', 'y', 'time', 'station', 'clim']
def obs_list():
    return [gr.PointSet('cand%d' % i, ND, 5, list(cols), pd.DataFrame(s['obs'], columns=cols)) for i, s in enumerate(stations)]
def step():
    return hmg.detect_many(grids, obs_list(), method='median', prob=0.95)
for _ in range(3): step()
pr = cProfile.Profile(); pr.enable()
for _ in range(20): step()
pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats('cumulative').print_stats(45); print(s.getvalue()[:9000])
