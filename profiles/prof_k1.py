# -*- coding: utf-8 -*-
"""K1 (GSLIB ASCII decode) timing on device-resident text.

    python profiles/prof_k1.py [nodes] [rows] [reps]

Row r of a synthetic stack is rendered to '%12.4f\\r\\n' text on the device
(gsb_encode_text_device, 14 bytes per node as interface/gui.py:1414 assumes),
then decoded back with gsb_decode_text_device in fixed-width mode (line_width
14) and in the general mode (line_width 0, newline index built on the device).
Algorithmic bytes per node-realization: 14 read + 4 written."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from gsimcli_b200 import _capi

N = int(sys.argv[1]) if len(sys.argv) > 1 else 4000000
ROWS = int(sys.argv[2]) if len(sys.argv) > 2 else 8
REPS = int(sys.argv[3]) if len(sys.argv) > 3 else 3
lib = _capi.lib()
src = _capi.Stack(ROWS, N)
src.synth_fill(1004, [200, 200, N // 40000], 0, -999.9)
dst = _capi.Stack(ROWS, N)
text = torch.empty((ROWS, N * 14), dtype=torch.uint8, device='cuda')
st = torch.cuda.current_stream().cuda_stream
for r in range(ROWS):
    _capi.check(lib.gsb_encode_text_device(src.handle, r, 0, N, 12, 4, 1,
                                           C.c_void_p(text[r].data_ptr()), C.c_void_p(st)))
bad = torch.full((1,), -1, dtype=torch.int64, device='cuda')
e0 = torch.cuda.Event(enable_timing=True)
e1 = torch.cuda.Event(enable_timing=True)
for width, name in ((14, 'fixed-width'), (0, 'general')):
    for rep in range(REPS):
        e0.record()
        for r in range(ROWS):
            _capi.check(lib.gsb_decode_text_device(
                dst.handle, r, C.c_void_p(text[r].data_ptr()), N * 14, 0, width, 0, 0, N,
                C.c_void_p(bad.data_ptr()), C.c_void_p(st)))
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        print('K1 %-11s %d rows x %d nodes: %.3f ms  %.1f GB/s (18 B per node-realization)  %.3e node-real/s'
              % (name, ROWS, N, ms, 18.0 * ROWS * N / ms / 1e6, ROWS * N / ms * 1e3))
    assert int(bad.item()) == -1
    a = src.download()
    b = dst.download()
    assert np.array_equal(a, b), 'decode(encode(x)) != x'
print('round trip exact')
