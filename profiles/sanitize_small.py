# -*- coding: utf-8 -*-
"""Small run of every kernel of the path for compute-sanitizer:
    compute-sanitizer --tool memcheck  python profiles/sanitize_small.py
    compute-sanitizer --tool racecheck python profiles/sanitize_small.py
K1 (fixed-width fast path, general form, encoder), K2 (counting sort at R = 50,
200 and 1000, register tile, bitonic at R = 20, moments), K2c + K3 (host-array
form and the device-resident plan)."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import pandas as pd
from gsimcli_b200 import _capi, synth
from gsimcli_b200.tools import grid as gr, homog as hmg

ND = -999.9
QS = [round(0.05 * k, 2) for k in range(1, 20)]
ALL = (_capi.MEAN | _capi.MEDIAN | _capi.SKEW | _capi.VAR | _capi.STD | _capi.COEFVAR
       | _capi.PERC | _capi.MODE)
dims, first, cells = [20, 16, 10], [0.0, 0.0, 1990], [100.0, 100.0, 1]
N = int(np.prod(dims))
chk = 0.0
for R, algos in ((20, (0,)), (50, (0, 3)), (200, (0, 3)), (1000, (0, 3))):
    g = gr.GridFiles.from_synthetic(5, R, dims, first, cells, ND)
    for algo in algos:
        g._stack.set_option('k2_algo', algo)
        out = g._stack.stats_grid(ALL, 0.95, QS, ND)
        chk += float(np.nansum(out))
    out = g._stack.stats_grid(_capi.MEAN | _capi.SKEW | _capi.VAR, 0.95, None, ND)   # moments kernel
    chk += float(np.nansum(out))
    if R == 200:
        host = synth.stack(5, R, dims)
        k1 = _capi.Stack(3, N)
        k1.decode_text(0, synth.to_text(host[0], '%12.4f', '\r\n'), 0, 14)
        k1.decode_text(1, synth.to_text(host[1], '%.4f', '\n', ['a', '1', 'v']), 3, 0)
        k1.decode_text(2, synth.to_text(host[2], '%-10.6f', '\n'), 0, 0)
        chk += float(np.nansum(k1.download()))
        k1.close()
        chk += float(_capi.encode_values(host[0], 12, 4, crlf=True).sum())
        sts = synth.stations(5, 4, R, dims, first, cells, tol=1)
        cols = ['x', 'y', 'time', 'station', 'clim']
        obs = lambda: [gr.PointSet('c%d' % i, ND, 5, list(cols), pd.DataFrame(s['obs'], columns=cols))
                       for i, s in enumerate(sts)]
        res = hmg.detect_many(g, obs(), rad=1, method='skewness', skewness=0.3)
        chk += sum(r[1] for r in res)
        plan = hmg.DetectPlan(g, obs(), rad=1, method='median')
        plan.enqueue()
        chk += float(plan.fetch()[3].sum())
    g.dump()
print('sanitize_small done, checksum', chk)
