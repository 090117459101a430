# -*- coding: utf-8 -*-
"""Generate ``tests/golden/*.json`` by running the reference's OWN code.

Run in the build container only (needs ``/root/reference``):

    python -m oracle.make_golden

For each case a small realization stack is written as GSLIB ASCII files, the
reference's ``GridFiles.load/stats/stats_area`` (tools/grid.py) and ``detect``
(tools/homog.py) are executed through ``oracle/refshim.py`` on those files, and
inputs + outputs are stored as JSON (Python ``repr`` floats: exact round trip).
The committed fixtures pin ``oracle/np_oracle.py`` (tests/test_oracle_golden.py)
and, through it, the CUDA path.
"""
import json
import os
import shutil
import sys
import tempfile
import warnings

import numpy as np
import pandas as pd

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import refshim  # noqa: E402
from gsimcli_b200 import synth  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                   'tests', 'golden')
ND = -999.9
ALL = dict(lmean=True, lmed=True, lskew=True, lvar=True, lstd=True,
           lcoefvar=True, lperc=True)


def _tolist(a):
    return np.asarray(a, dtype=np.float64).tolist()


def _frame(df):
    return {'columns': [str(c) for c in df.columns],
            'values': _tolist(df.values)}


def run_case(name, st, dims, first, cells, fmt, eol, header, stations,
             detect_cases, p_list=(0.95, 0.9), tols=(0, 1, 2)):
    """Execute the reference on stack ``st`` and collect everything."""
    _, grid, homog = refshim.load_reference()
    tmp = tempfile.mkdtemp(prefix='gsb_golden_')
    case = {'name': name, 'dims': list(dims), 'first_coord': list(first),
            'cells_size': list(cells), 'nodata': ND, 'fmt': fmt,
            'eol': eol, 'header': header, 'stack': _tolist(st),
            'stats': [], 'stats_area': [], 'detect': []}
    try:
        hdr = ['golden', '1', 'var'] if header else None
        firstfile = synth.write_stack(tmp, 'g_sim', st, fmt, eol, hdr)
        R = st.shape[0]
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            for p in p_list:
                g = grid.GridFiles()
                g.load(firstfile, R, dims, first, cells, ND, header)
                res = g.stats(p=p, **ALL)
                g.dump()
                case['stats'].append({
                    'p': p, 'maps': {k: _tolist(v.val) for k, v in res.items()}})
            if not header:      # stats_area is broken for header>0
                for loc in stations:
                    for tol in tols:
                        g = grid.GridFiles()
                        g.load(firstfile, R, dims, first, cells, ND, header)
                        ps = g.stats_area(list(loc), tol=tol, p=0.95, **ALL)
                        g.dump()
                        case['stats_area'].append({
                            'loc': list(loc), 'tol': tol, 'p': 0.95,
                            'name': ps.name, 'table': _frame(ps.values)})
                for dc in detect_cases:
                    g = grid.GridFiles()
                    g.load(firstfile, R, dims, first, cells, ND, header)
                    df = pd.DataFrame(np.array(dc['obs'], dtype=float),
                                      columns=['x', 'y', 'time', 'station',
                                               'clim'])
                    ps = grid.PointSet('cand', ND, 5, list(df.columns), df)
                    kw = {k: v for k, v in dc.items() if k != 'obs'}
                    hom, dn, md = homog.detect(g, ps, **kw)
                    g.dump()
                    case['detect'].append({
                        'obs': dc['obs'], 'kwargs': kw, 'name': hom.name,
                        'homogenised': _frame(hom.values),
                        'detected_number': int(dn), 'missing_data': int(md)})
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    os.makedirs(OUT, exist_ok=True)
    with open(os.path.join(OUT, name + '.json'), 'w') as fh:
        json.dump(case, fh)
    print('wrote', name, os.path.getsize(os.path.join(OUT, name + '.json')))


def station_obs(st, dims, first, cells, ix, iy, sid, rng, drop=(), nd_rows=(),
                shift=None):
    """Observed series at node (ix, iy) (1-based): an off-stack perturbation of
    realization 0 with a break, nodata rows and dropped (missing) years."""
    dx, dy, dz = dims
    nodes = (ix - 1) + dx * (iy - 1) + dx * dy * np.arange(dz)
    clim = st[0, nodes].astype(np.float64)
    clim[clim == ND] = 700.0
    if shift:
        a, b, d = shift
        clim[a:b] += d
    for k in nd_rows:
        clim[k] = ND
    wx = first[0] + (ix - 1) * cells[0] + 0.2 * cells[0]
    wy = first[1] + (iy - 1) * cells[1] - 0.3 * cells[1]
    t = first[2] + cells[2] * np.arange(dz)
    rows = [[wx, wy, float(t[k]), float(sid), float(clim[k])]
            for k in range(dz) if k not in drop]
    return (wx, wy), rows


def special_columns(st):
    """Overwrite a few columns with the edge cases the reference's numerics
    define: constant column, all nodata, n == 1, n == 2, ties."""
    st = st.copy()
    R = st.shape[0]
    st[:, 3] = 812.5
    st[:, 4] = ND
    st[:, 5] = ND
    st[R // 2, 5] = 640.25
    st[:, 6] = ND
    st[0, 6] = 7.5
    st[R - 1, 6] = 8.5
    st[:, 8] = np.where(np.arange(R) % 2 == 0, 500.0, 600.0)
    return st


def main():
    if not refshim.available():
        raise SystemExit('reference checkout not available; cannot regenerate')
    # --- case 1: SURVEY Appendix C known-answer stack -----------------------
    cols = {0: [3, 1, ND, 2, 10, 4], 1: [5] * 6,
            2: [7.5, ND, ND, ND, ND, 8.5], 3: [ND] * 6}
    st = np.array([[cols[n][r] for n in range(4)] for r in range(6)])
    dets = [dict(obs=[[0., 0., 2000., 1., c0], [0., 0., 2001., 1., ND]],
                 method='median', prob=0.95) for c0 in (1.1, 1.0999, 9.5)]
    run_case('kat_appendix_c', st, [2, 1, 2], [0, 0, 2000], [10, 10, 1],
             '%12.4f', '\r\n', 0, [(0, 0), (10, 0)], dets, tols=(0,))

    # --- cases 2/3: dyadic synthetic stacks, even and odd R -----------------
    rng = np.random.Generator(np.random.PCG64(42))
    for name, R in (('dyadic_even', 10), ('dyadic_odd', 9)):
        dims, first, cells = [9, 8, 6], [1000.0, 5000.0, 1990], [100.0, 50.0, 1]
        st = special_columns(synth.stack(11 + R, R, dims).astype(np.float64))
        st[st == np.float64(np.float32(ND))] = ND
        locs, dets = [], []
        opt = dict(lmean=True, lmed=True, lskew=True, lvar=True, lstd=True,
                   lcoefvar=True, lperc=True)
        specs = [
            (4, 4, 1, (), (), (2, 4, 300.0)),
            (5, 3, 2, (), (1, 4), (0, 2, -300.0)),
            (3, 5, 3, (2, 5), (0,), (3, 6, 250.0)),
        ]
        for ix, iy, sid, drop, nd_rows, shift in specs:
            loc, rows = station_obs(st, dims, first, cells, ix, iy, sid, rng,
                                    drop, nd_rows, shift)
            locs.append(loc)
            dets.append(dict(obs=rows, method='mean', prob=0.95))
            dets.append(dict(obs=rows, method='median', prob=0.9, rad=1))
            dets.append(dict(obs=rows, method='skewness', skewness=0.3,
                             prob=0.95))
            dets.append(dict(obs=rows, method='percentile', prob=0.95,
                             percentile=0.95))
            dets.append(dict(obs=rows, method='percentile', prob=0.95,
                             percentile=0.8, rad=2))
            dets.append(dict(obs=rows, method='median', prob=0.95,
                             optional_stats=opt))
            dets.append(dict(obs=rows, method='percentile', prob=0.9,
                             percentile=0.9, optional_stats=dict(lperc=True)))
        run_case(name, st, dims, first, cells, '%12.4f', '\r\n', 0, locs, dets)

    # --- case 4: arbitrary 4-decimal values, LF, variable width -------------
    rng = np.random.Generator(np.random.PCG64(7))
    dims, first, cells = [6, 5, 4], [0.0, 0.0, 1.0], [1.0, 1.0, 1.0]
    N = int(np.prod(dims))
    st = np.round(rng.gamma(2.0, 150.0, size=(8, N)) + 300.0, 4)
    st[rng.uniform(size=st.shape) < 0.04] = ND
    loc, rows = station_obs(st, dims, first, cells, 3, 3, 1, rng, (1,), (2,),
                            (0, 2, 400.0))
    dets = [dict(obs=rows, method='mean', prob=0.95),
            dict(obs=rows, method='skewness', skewness=0.5, prob=0.9)]
    run_case('decimals_lf', st, dims, first, cells, '%.4f', '\n', 0, [loc],
             dets, tols=(0, 1))

    # --- case 5: 3 header lines; stats() only ----------
    st5 = np.round(rng.normal(800.0, 100.0, size=(7, 24)), 3)
    run_case('header3_stats', st5, [4, 3, 2], [0, 0, 0], [1, 1, 1], '%-10.6f',
             '\n', 3, [], [], p_list=(0.95,))


if __name__ == '__main__':
    main()
