# -*- coding: utf-8 -*-
"""BASELINE config 4 in small: GSLIB ASCII decode + stats end to end from pinned
host text buffers.

    python profiles/prof_cfg4.py [files] [nodes]

`files` realization files of `nodes` lines ('%12.4f\\r\\n', 14 B per line) sit in
pinned host memory; the timed region uploads and decodes every file
(gsb_decode_text, fixed-width mode) and then runs the full-grid statistics
(moments + median + percentiles 5..95 + mode).  Reported: text GB/s against the
PCIe ceiling, decode-only GB/s on resident text, node-realizations/s."""
import ctypes as C
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from gsimcli_b200 import _capi

FILES = int(sys.argv[1]) if len(sys.argv) > 1 else 100
N = int(sys.argv[2]) if len(sys.argv) > 2 else 4000000
lib = _capi.lib()
QS = [round(0.05 * k, 2) for k in range(1, 20)]
src = _capi.Stack(FILES, N)
src.synth_fill(1004, [200, 200, max(1, N // 40000)], 0, -999.9)
st = torch.cuda.current_stream().cuda_stream
host = torch.empty((FILES, N * 14), dtype=torch.uint8).pin_memory()
dtext = torch.empty((N * 14,), dtype=torch.uint8, device='cuda')
for r in range(FILES):                                   # setup: render + download the text
    _capi.check(lib.gsb_encode_text_device(src.handle, r, 0, N, 12, 4, 1,
                                           C.c_void_p(dtext.data_ptr()), C.c_void_p(st)))
    host[r].copy_(dtext)
torch.cuda.synchronize()
ref = src.download(0, FILES, 0, min(N, 100000))
dst = _capi.Stack(FILES, N)
flags = (_capi.MEAN | _capi.MEDIAN | _capi.SKEW | _capi.VAR | _capi.STD | _capi.COEFVAR | _capi.MODE)
npl = lib.gsb_stats_grid_planes(flags, len(QS))
out = torch.empty((npl, N), dtype=torch.float32, device='cuda')
bad = C.c_int64(-1)
for rep in range(2):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for r in range(FILES):
        _capi.check(lib.gsb_decode_text(dst.handle, r, C.c_void_p(host[r].data_ptr()), N * 14, 0, 14,
                                        0, 0, N, C.byref(bad), C.c_void_p(st)))
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    dst.stats_grid_device(flags, 0.95, QS, -999.9, 0, N, out.data_ptr(), N, st)
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    gb = FILES * N * 14 / 1e9
    print('cfg4 x%d files of %d lines: upload+decode %.1f ms (%.1f GB/s of text), stats %.1f ms, '
          'end to end %.3e node-realizations/s' % (FILES, N, (t1 - t0) * 1e3, gb / (t1 - t0),
                                                   (t2 - t1) * 1e3, FILES * N / (t2 - t0)))
assert np.array_equal(dst.download(0, FILES, 0, ref.shape[1]), ref), 'decoded stack differs'
print('decoded stack identical to the source stack')
