# -*- coding: utf-8 -*-
"""Worker of tests/test_multi_gpu.py (launched by torch.distributed.run, one
rank per GPU): node-sharded stack -> full-grid statistics of the rank's slab +
the device-resident station path (K2c -> NCCL all-gather -> K3).  Rank 0
collects everything, checks it against the oracle and writes a JSON summary."""
import json
import os
import sys
import zlib

import numpy as np
import pandas as pd

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main(out_path):
    import torch
    import torch.distributed as td
    from gsimcli_b200 import _capi, synth, dist
    from gsimcli_b200.tools import grid as gr, homog as hmg
    from oracle import np_oracle as O
    world = int(os.environ.get('WORLD_SIZE', 1))
    rank = int(os.environ.get('RANK', 0))
    local = int(os.environ.get('LOCAL_RANK', 0))
    torch.cuda.set_device(local)
    if world > 1:
        td.init_process_group('nccl', device_id=torch.device('cuda', local))
    nd = -999.9
    dims, first, cells = [24, 20, 13], [0.0, 0.0, 1990], [100.0, 100.0, 1]   # 13 layers: uneven slabs
    R, seed = 200, 4242
    n0, n1 = dist.slab_range(dims, rank, world)
    grids = gr.GridFiles.from_synthetic(seed, R, dims, first, cells, nd, device=local,
                                        node_range=(n0, n1) if world > 1 else None)
    qs = [0.05, 0.5, 0.95]
    flags = _capi.MEAN | _capi.MEDIAN | _capi.SKEW | _capi.PERC | _capi.MODE
    maps = grids._stack.stats_grid(flags, 0.95, qs, nd)            # [planes][nn]
    sts = synth.stations(seed, 7, R, dims, first, cells, tol=1)
    cols = ['x', 'y', 'time', 'station', 'clim']
    obs = [gr.PointSet('c%d' % i, nd, 5, list(cols), pd.DataFrame(s['obs'], columns=cols))
           for i, s in enumerate(sts)]
    plan = hmg.DetectPlan(grids, obs, rad=1, method='skewness', skewness=0.3, prob=0.95)
    side = torch.cuda.Stream()
    plan.enqueue(side)
    filled, homog, flag, det = [np.array(a, copy=True) for a in plan.fetch()]
    table = plan.table()
    parts = [None] * world
    if world > 1:
        td.all_gather_object(parts, (n0, n1, maps, flag, homog, det, table))
    else:
        parts = [(n0, n1, maps, flag, homog, det, table)]
    if rank == 0:
        N = int(np.prod(dims))
        full = np.empty((maps.shape[0], N), dtype=np.float32)
        for a, b, m, *_ in parts:
            full[:, a:b] = m
        same_tables = all(np.array_equal(p[3], flag) and np.array_equal(p[4], homog)
                          and np.array_equal(p[5], det) and np.array_equal(p[6], table, equal_nan=True)
                          for p in parts)
        host = synth.stack(seed, R, dims)
        ref = O.grid_stats(host, nd, O.FLAG_MEAN | O.FLAG_MED | O.FLAG_SKEW | O.FLAG_PERC
                           | O.FLAG_MODE, 0.95, qs)
        f32 = lambda a: np.asarray(a, np.float64).astype(np.float32)
        eq = lambda a, b: bool(((f32(a) == f32(b)) | (np.isnan(f32(a)) & np.isnan(f32(b)))).all())
        grid_ok = (eq(full[1], ref['medianmap']) and eq(full[3], ref['percmap'][:, 0])
                   and eq(full[4], ref['percmap'][:, 1])
                   and all(eq(full[5 + i], ref['percentiles'][:, i]) for i in range(3))
                   and eq(full[8], ref['modemap'])
                   and np.allclose(full[0], ref['meanmap'], rtol=1e-5)
                   and np.allclose(full[2], ref['skewmap'], rtol=1e-5, atol=1e-5))
        st_ok = True
        for s, stn in enumerate(sts):
            df = pd.DataFrame(stn['obs'], columns=cols)
            ohom, odn, omd = O.detect(host, dims, first, cells, nd, df, nd, rad=1,
                                      method='skewness', skewness=0.3, prob=0.95)
            st_ok = st_ok and odn == int(det[s]) and np.array_equal(flag[s], ohom['Flag'].values) \
                and np.allclose(homog[s], ohom['clim'].values, rtol=1e-12)
        with open(out_path, 'w') as fh:
            json.dump({'world': world, 'grid_ok': bool(grid_ok), 'stations_ok': bool(st_ok),
                       'tables_identical_on_all_ranks': bool(same_tables),
                       'maps_crc': zlib.crc32(full.tobytes()),
                       'flags_crc': zlib.crc32(flag.tobytes()),
                       'homog_crc': zlib.crc32(homog.tobytes()),
                       'table_crc': zlib.crc32(np.nan_to_num(table, nan=-1.0).tobytes()),
                       'detected': int(det.sum())}, fh)
    grids.dump()
    if world > 1:
        td.destroy_process_group()


if __name__ == '__main__':
    main(sys.argv[1])
