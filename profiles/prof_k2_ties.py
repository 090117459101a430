"""K2 on heavily tied data (quantised values): times the counting-sort kernel's
over-full-bin path against the bitonic kernel.  python profiles/prof_k2_ties.py [levels]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from gsimcli_b200 import _capi
levels = int(sys.argv[1]) if len(sys.argv) > 1 else 40
R, N = 1000, 200000
QS = [round(0.05 * k, 2) for k in range(1, 20)]
g = torch.Generator(device='cuda'); g.manual_seed(5)
vals = (torch.randint(0, levels, (R, N), device='cuda', generator=g).float() * 0.25 + 500.0)
stack = _capi.Stack(R, N)
host = vals.cpu().numpy()
stack.upload(host)
flags = _capi.MEAN | _capi.MEDIAN | _capi.SKEW | _capi.VAR | _capi.STD | _capi.COEFVAR | _capi.MODE
npl = _capi.lib().gsb_stats_grid_planes(flags, len(QS))
out = torch.empty((npl, N), dtype=torch.float32, device='cuda')
st = torch.cuda.current_stream().cuda_stream
res = {}
for algo, code in (('csort', 1), ('bitonic', 2), ('rsel', 3)):
    stack.set_option('k2_algo', code)       # test-only kernel selection
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    for i in range(3):
        e0.record(); stack.stats_grid_device(flags, 0.95, QS, -999.9, 0, N, out.data_ptr(), N, st); e1.record()
        torch.cuda.synchronize()
    print('%d levels, %s: %.3f ms' % (levels, algo, e0.elapsed_time(e1)))
    res[algo] = out[1:].clone()     # all planes but the mean (moments differ in the last bits)
q = slice(5, None)
assert torch.equal(res['csort'][q].view(torch.int32), res['bitonic'][q].view(torch.int32)), 'order statistics differ'
assert torch.equal(res['rsel'][q].view(torch.int32), res['bitonic'][q].view(torch.int32)), 'order statistics differ (rsel)'
print('order statistics and mode identical')
