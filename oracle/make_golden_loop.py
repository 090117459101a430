# -*- coding: utf-8 -*-
"""Generate ``tests/golden/candidate_loop.json`` by running the reference's OWN
candidate loop (launchers/method_classic.py:151-248 with ``skip_dss=True``) on
pre-existing realization files.  Build container only:

    python -m oracle.make_golden_loop

The launcher module itself imports the DSS / GUI machinery, so the loop body is
driven here exactly as ``gsimcli()`` drives it -- ``take_candidate`` ->
``GridFiles.load`` -> ``detect`` -> ``update_station`` per candidate, then
``save_output(fformat='gsimcli', save_stations=True)`` -- with the reference's
tools loaded through ``oracle/refshim.py``.  Inputs (network point-set, one
small realization stack per candidate) and outputs (detections, fills, the
final CSVs) are stored as JSON.
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
                   'tests', 'golden', 'candidate_loop.json')
ND = -999.9


def network(dims, first, cells, rng):
    """Three stations on grid nodes, complete yearly series (the reference
    writes a homogenised station back by row label, ``DataFrame.update``: a
    series that ``fill_station`` had to extend comes back with fresh labels and
    lands on other stations' rows, so the loop is only meaningful for series
    without absent years); stations 20 and 30 have a nodata value each."""
    rows = []
    for sid, (ix, iy) in ((30, (2, 3)), (10, (5, 2)), (20, (7, 6))):
        x = first[0] + (ix - 1) * cells[0]
        y = first[1] + (iy - 1) * cells[1]
        for t in range(dims[2]):
            year = first[2] + t
            v = np.round(rng.normal(800.0, 160.0) * 16) / 16
            if (sid == 30 and t == 3) or (sid == 20 and t == 1):
                v = ND
            rows.append([x, y, float(year), float(sid), v])
    return pd.DataFrame(rows, columns=['x', 'y', 'time', 'station', 'clim'])


def main():
    _, grid, homog = refshim.load_reference()
    orig = homog.list_stations
    homog.list_stations = lambda *a, **k: list(orig(*a, **k))     # py2 map() is a list
    dims, first, cells = [9, 8, 6], [1000.0, 5000.0, 1990], [100.0, 50.0, 1]
    nsim = 16
    rng = np.random.default_rng(11)
    df = network(dims, first, cells, rng)
    order = [20, 30, 10]
    stacks = [synth.stack(500 + i, nsim, dims).astype(np.float64) for i in range(len(order))]
    tmp = tempfile.mkdtemp(prefix='gsb_loop_')
    outfolder = os.path.join(tmp, 'net')
    os.makedirs(outfolder)
    basename = os.path.basename(outfolder)
    case = {'dims': dims, 'first_coord': first, 'cells_size': cells, 'nodata': ND,
            'nsim': nsim, 'order': order, 'method': 'median', 'prob': 0.95,
            'stations': {'columns': list(df.columns), 'values': df.values.tolist()},
            'stacks': [s.tolist() for s in stacks]}
    try:
        stations_pset = grid.PointSet('net', ND, 5, list(df.columns), df.copy())
        dn, fn = [], []
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            for i, sid in enumerate(order):
                candidate, references = homog.take_candidate(stations_pset, sid)
                outfile = os.path.join(outfolder, basename + '_dss_map_st' + str(i) + '_sim.out')
                synth.write_stack(outfolder, basename + '_dss_map_st' + str(i) + '_sim',
                                  stacks[i], '%12.4f', '\r\n', None)
                sim_maps = grid.GridFiles()
                sim_maps.load(outfile, nsim, dims, first, cells, ND, headerin=0)
                hom, d, f = homog.detect(grids=sim_maps, obs_file=candidate, method='median',
                                         prob=0.95, flag=True, save=False, outfile=None,
                                         header=True, skewness=None, rad=0, percentile=None,
                                         optional_stats=None)
                dn.append(int(d))
                fn.append(int(f))
                stations_pset = homog.update_station(stations_pset, hom)
                sim_maps.dump()
            csvfile = os.path.join(outfolder, basename + '_homogenised_data.csv')
            homog.save_output(pset_file=stations_pset, outfile=csvfile, fformat='gsimcli',
                              header=True, save_stations=True)
        case['detected'] = dn
        case['filled'] = fn
        case['csv'] = open(csvfile).read()
        case['stations_csv'] = open(os.path.join(outfolder, basename
                                                 + '_homogenised_data_stations.csv')).read()
        case['final'] = {'columns': list(stations_pset.values.columns),
                         'values': np.asarray(stations_pset.values.values, float).tolist()}
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    with open(OUT, 'w') as fh:
        json.dump(case, fh)
    print('wrote', OUT, 'detected', dn, 'filled', fn)


if __name__ == '__main__':
    main()
