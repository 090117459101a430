"""Randomised cross-check of the K2 sort kernels (register tile, 32 and 16 warps,
vs counting sort vs register bitonic; forced through the test-only stack options
k2_algo / k2_nw): order statistics and mode must agree bit for bit on every column of
many random stacks (mixtures of smooth, tied, clamped, clustered and outlier
columns, random R and missing-value rates).  python profiles/stress_k2.py [rounds]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from gsimcli_b200 import _capi

rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 40
ND = -999.9
rng = np.random.default_rng(2026)
flags = _capi.MEAN | _capi.MEDIAN | _capi.SKEW | _capi.VAR | _capi.STD | _capi.COEFVAR | _capi.PERC | _capi.MODE
bad = 0
for it in range(rounds):
    R = int(rng.integers(33, 1025))
    N = int(rng.integers(3000, 9000))
    kind = rng.integers(0, 8, size=N)
    base = rng.normal(0.0, 1.0, size=(R, N))
    st = np.empty((R, N))
    for k in range(8):
        m = kind == k
        if not m.any():
            continue
        x = base[:, m]
        if k == 0:   x = np.round((800 + 150 * x) * 16) / 16                          # smooth, 1/16 lattice
        elif k == 1: x = np.round(x * rng.integers(1, 30)) / 4.0 + 500                 # heavy ties
        elif k == 2: x = np.clip(np.round((800 + 300 * x) * 16) / 16, 600, 1000)      # clamped both ends
        elif k == 3: x = np.where(x > 0.3, 0.0, np.round(np.abs(x) * 800) / 16)       # zero-inflated
        elif k == 4: x = x + np.where(base[::-1, m] > 0, -1.0e6, 1.0e6)                # two far clusters
        elif k == 5: x = 1.0e7 + np.round(x * 64) / 16                                 # big offset, tiny range
        elif k == 6:                                                                   # outliers
            x = 800 + 150 * x
            x[rng.integers(0, R, size=3), :] *= 1.0e12
        else:        x = np.exp(3 * x)                                                 # heavy right tail
        st[:, m] = x
    miss = rng.uniform(size=st.shape) < rng.choice([0.0, 0.001, 0.05, 0.6])
    st[miss] = ND
    st32 = np.ascontiguousarray(st, dtype=np.float32)
    qs = sorted(set(np.round(rng.uniform(0, 1, size=int(rng.integers(1, 12))), 3).tolist()))
    stack = _capi.Stack(R, N)
    stack.upload(st32)
    outs = {}
    pp = float(rng.choice([0.9, 0.95, 0.5]))
    for algo, (k2_algo, k2_nw) in (('rsel', (3, 0)), ('rsel16', (3, 16)), ('csort', (1, 0)), ('bitonic', (2, 0))):
        stack.set_option('k2_algo', k2_algo)
        stack.set_option('k2_nw', k2_nw)
        outs[algo] = stack.stats_grid(flags, pp, qs, ND)
    stack.close()
    b = outs['bitonic']
    # planes: mean, median, skew, var, std, coefvar, lperc, rperc, q..., mode
    exact = [1, 6, 7] + list(range(8, b.shape[0]))
    tame = np.isin(kind, [0, 1, 2, 3])       # float32 moments are only comparable without 1e12 outliers / 1e7 offsets
    for algo in ('rsel', 'rsel16', 'csort'):
        a = outs[algo]
        same = (a[exact].view(np.int32) == b[exact].view(np.int32)) | (np.isnan(a[exact]) & np.isnan(b[exact]))
        nbad = int((~same).sum())
        mom_ok = np.allclose(a[[0, 3, 4]][:, tame], b[[0, 3, 4]][:, tame], rtol=2e-4, atol=1e-6, equal_nan=True)
        print('round %2d R=%4d N=%5d nq=%2d %-6s vs bitonic: exact-mismatches=%d moments-close=%s' % (it, R, N, len(qs), algo, nbad, mom_ok), flush=True)
        if nbad:
            idx = np.argwhere(~same)[:5]
            for pl, col in idx:
                print('   plane', exact[pl], 'col', col, 'kind', kind[col], a[exact[pl], col], b[exact[pl], col])
            bad += nbad
print('TOTAL exact mismatches:', bad)
sys.exit(1 if bad else 0)
