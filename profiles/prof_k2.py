# -*- coding: utf-8 -*-
"""Small driver for ncu: fills a synthetic stack and runs the K2 grid kernel a
few times.  Usage: python profiles/prof_k2.py [R] [dx,dy,dz] [mode|nomode|moments] [reps]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from gsimcli_b200 import _capi

R = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
dims = [int(x) for x in (sys.argv[2] if len(sys.argv) > 2 else '200,200,10').split(',')]
kind = sys.argv[3] if len(sys.argv) > 3 else 'mode'
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
N = int(os.environ.get("GSB_PROF_N", np.prod(dims)))
NODE0 = int(os.environ.get('GSB_PROF_NODE0', 0))
QS = [round(0.05 * k, 2) for k in range(1, 20)]
stack = _capi.Stack(R, N)
stack.synth_fill(1003, dims, NODE0, -999.9)
# test-only kernel selection: register tile (32 / 16 warps), counting sort, bitonic
ALGO = os.environ.get('GSB_PROF_ALGO', 'csort')
k2_algo, k2_nw = {'rsel': (3, 0), 'rsel16': (3, 16), 'csort': (1, 0), 'bitonic': (2, 0)}[ALGO]
stack.set_option('k2_algo', k2_algo)
stack.set_option('k2_nw', k2_nw)
if os.environ.get('GSB_K2_STATS'):
    stack.set_option('k2_stats', 1)
flags = _capi.MEAN | _capi.SKEW | _capi.VAR | _capi.STD | _capi.COEFVAR
qs = None
if kind == 'refapi':          # the reference's stats(): moments + median + percentile pair
    flags |= _capi.MEDIAN | _capi.PERC
elif kind != 'moments':
    flags |= _capi.MEDIAN
    qs = QS
if kind == 'mode':
    flags |= _capi.MODE
npl = _capi.lib().gsb_stats_grid_planes(flags, len(qs) if qs else 0)
out = torch.empty((npl, N), dtype=torch.float32, device='cuda')
st = torch.cuda.current_stream().cuda_stream
e0 = torch.cuda.Event(enable_timing=True)
e1 = torch.cuda.Event(enable_timing=True)
for i in range(reps):
    e0.record()
    stack.stats_grid_device(flags, 0.95, qs, -999.9, 0, N, out.data_ptr(), N, st)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    gbs = (4.0 * R * N + 4.0 * npl * N) / ms / 1e6
    print('%s %s R=%d N=%d planes=%d: %.3f ms  %.1f GB/s  %.3e node-real/s  flagged tiles %d' % (
        ALGO, kind, R, N, npl, ms, gbs, R * N / ms * 1e3, stack.k2_flagged()))
if os.environ.get('GSB_K2_STATS'):
    import ctypes
    buf = (ctypes.c_ulonglong * 8)()
    _capi.lib().gsb_debug_cs_counters(buf, 1)
    c = list(buf)
    print('csort counters: columns %d, with big bins %d (%.3f%%), generic %d (%.3f%%), mean phases %.2f, big bins %d' % (
        c[0], c[1], 100.0 * c[1] / max(c[0], 1), c[2], 100.0 * c[2] / max(c[0], 1), c[3] / max(c[0], 1), c[4]))
    print('tile wait: %.1f%% of warp time (%.0f cycles per column)' % (100.0 * c[5] / max(c[6], 1), c[5] / max(c[0], 1)))
print('checksum', float(torch.nan_to_num(out).double().sum()))
