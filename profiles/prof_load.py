# -*- coding: utf-8 -*-
"""GridFiles.load from realization FILES (the public path of config 4): R files of
200x200x100 nodes, '%12.4f\\r\\n' records, written to a scratch directory from
the synthetic stack (device encoder), then loaded through the pinned read ring
(file read -> pinned buffer -> H2D -> K1) and reduced.  python profiles/prof_load.py [R]"""
import os
import shutil
import sys
import tempfile
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from gsimcli_b200 import _capi
from gsimcli_b200.tools import grid as gr
from gsimcli_b200.tools.utils import filename_indexing

R = int(sys.argv[1]) if len(sys.argv) > 1 else 40
dims, first, cells, ND = [200, 200, 100], [0.0, 0.0, 1900], [1000.0, 1000.0, 1], -999.9
N = int(np.prod(dims))
src = gr.GridFiles.from_synthetic(1003, R, dims, first, cells, ND)
ref = src._stack.stats_grid(_capi.MEDIAN, 0.95, None, ND)
tmp = tempfile.mkdtemp(prefix='gsb_load_')
try:
    dtext = torch.empty(N * 14, dtype=torch.uint8, device='cuda')
    firstfile = os.path.join(tmp, 'net_dss_map_st0_sim.out')
    t0 = time.perf_counter()
    for r in range(R):
        _capi.check(_capi.lib().gsb_encode_text_device(src._stack.handle, r, 0, N, 12, 4, 1,
                                                       dtext.data_ptr(), None))
        path = firstfile if r == 0 else filename_indexing(firstfile, r + 1)
        with open(path, 'wb') as fh:
            fh.write(dtext.cpu().numpy().tobytes())
    src.dump()
    print('wrote %d files, %.2f GB, in %.1f s' % (R, R * N * 14 / 1e9, time.perf_counter() - t0), flush=True)
    for rep in range(2):
        g = gr.GridFiles()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        g.load(firstfile, R, dims, first, cells, ND, headerin=0)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        got = g._stack.stats_grid(_capi.MEDIAN, 0.95, None, ND)
        ok = np.array_equal(got.view(np.int32), ref.view(np.int32))
        print('GridFiles.load of %d files (page cache %s): %.3f s = %.2f GB/s of text, %.3e node-realizations/s; median map identical: %s'
              % (R, 'warm' if rep else 'as written', dt, R * N * 14 / dt / 1e9, R * N / dt, ok), flush=True)
        g.dump()
finally:
    shutil.rmtree(tmp, ignore_errors=True)
