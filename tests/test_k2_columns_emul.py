# -*- coding: utf-8 -*-
"""CPU run of K2c (station columns) and K3 (detect / correct):
``gsimcli_b200/csrc/k2_columns_kernel.cuh`` -- the kernels nvcc compiles, SASS
byte-identical after the move into the header -- built by g++ and run with one
OS thread per CUDA thread (tests/emul/k2_columns_threads.cpp).  Fuzz against the
oracle's scalar column statistics (float64 moments to 1e-12, order statistics and
mode exactly) and against a restatement of homog.py:191-239; ThreadSanitizer over
the block-wide sort and the fixed-order reductions."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from emul import build_emul, fuzz_csort
from oracle import np_oracle as O

ND = -999.9
HAVE_CUDA_H = os.path.exists(os.path.join(os.environ.get('CUDA_HOME', '/usr/local/cuda'), 'include', 'cuda.h'))
pytestmark = pytest.mark.skipif(not HAVE_CUDA_H, reason='CUDA headers not installed')
KIND = np.array([0, 1, 2, 3, 4, 5, 1, 1, 1, 6], np.uint8)      # mean med skew var std cv lperc rperc q mode


def test_thread_sanitizer():
    exe = build_emul.build_columns('thread')
    r = subprocess.run([exe], capture_output=True, text=True, timeout=900,
                       env=dict(os.environ, TSAN_OPTIONS='halt_on_error=0 exitcode=66'))
    assert r.returncode == 0 and 'clean' in r.stdout, (r.stdout[-1500:], r.stderr[-3000:])
    assert 'ThreadSanitizer' not in r.stderr


@pytest.mark.parametrize('seed', [1, 2, 3])
def test_station_columns_fuzz_against_the_oracle(seed):
    lib = C.CDLL(build_emul.build_columns('lib'))
    rng = np.random.default_rng(seed)
    for _ in range(12):
        R = int(rng.choice([2, 3, 7, 20, 50, 100, 200]))
        K = int(rng.choice([1, 5, 13]))
        nodes = 40
        data = np.stack([fuzz_csort.column(rng, R, int(rng.integers(0, 8))) for _ in range(nodes)],
                        axis=1).astype(np.float32)
        buf = np.full((R, 64), 7.0e30, np.float32)
        buf[:, :nodes] = data
        ncols = 9
        ids = rng.integers(0, nodes, size=(ncols, K)).astype(np.int64)
        ids[rng.uniform(size=ids.shape) < 0.15] = -1
        ids[0, :] = -1                                         # a column with no node at all
        p = float(rng.choice([0.95, 0.9, 0.5]))
        qx = float(rng.uniform(0, 1))
        q = np.array([0, .5, 0, 0, 0, 0, (1 - p) / 2, 1 - (1 - p) / 2, qx, 0], np.float64)
        table = np.full((ncols, len(KIND)), -7.0)
        for force_global in (0, 1):
            rc = lib.k2c_emulate(C.c_void_p(buf.ctypes.data), C.c_longlong(64), C.c_int(R), C.c_int(K),
                                 C.c_void_p(ids.ctypes.data), C.c_longlong(ncols), C.c_float(ND),
                                 C.c_int(len(KIND)), C.c_void_p(KIND.ctypes.data), C.c_void_p(q.ctypes.data),
                                 C.c_void_p(table.ctypes.data), C.c_int(force_global))
            assert rc == 0
            for c in range(ncols):
                # reference order: arr[i + j*K] (grid.py:823)
                arr = np.full(R * K, ND, np.float64)
                for i in range(K):
                    if ids[c, i] >= 0:
                        arr[i::K] = data[:, ids[c, i]].astype(np.float64)
                arr[np.isclose(arr, np.float64(np.float32(ND)))] = ND
                ref = O.column_stats(arr, ND, 127 | O.FLAG_MODE, p, [qx])
                got = table[c]
                exp = [ref.get('mean'), ref.get('median'), ref.get('skewness'), ref.get('variance'), ref.get('std'),
                       ref.get('coefvar'), ref.get('lperc'), ref.get('rperc'),
                       ref['percentiles'][0] if 'percentiles' in ref else np.nan, ref.get('mode')]
                for k, (g, e) in enumerate(zip(got, exp)):
                    e = np.nan if e is None else float(e)
                    if np.isnan(e) or np.isnan(g):
                        assert np.isnan(e) and np.isnan(g), (seed, R, K, c, k, g, e)
                    elif KIND[k] in (1, 6):
                        assert g == e, (seed, R, K, c, k, g, e)
                    else:
                        assert abs(g - e) <= 1e-12 * abs(e) + 1e-300 or (KIND[k] == 2 and abs(g - e) < 1e-9), \
                            (seed, R, K, c, k, g, e)


def test_detect_kernel_against_homog_restatement():
    """homog.py:191-239 on float32-representable observations: fill NaN with the
    local mean, flag outside the inclusive interval, substitute per method."""
    lib = C.CDLL(build_emul.build_columns('lib'))
    rng = np.random.default_rng(4)
    for method in (0, 1, 2, 3):          # GSB_METHOD_MEAN, MEDIAN, SKEWNESS, PERCENTILE
        S, dz = 5, 37
        f32 = lambda a: a.astype(np.float32).astype(np.float64)
        clim = f32(rng.normal(10, 4, (S, dz)))
        clim[rng.uniform(size=clim.shape) < 0.1] = np.nan
        mean = f32(rng.normal(10, 1, (S, dz)))
        lo = f32(mean - rng.uniform(0, 5, (S, dz)))
        hi = f32(mean + rng.uniform(0, 5, (S, dz)))
        clim[0, :5] = lo[0, :5]                              # ties with the bounds: inside
        clim[1, :5] = hi[1, :5]
        fa, fb, fc = f32(rng.normal(9, 1, (S, dz))), f32(rng.normal(11, 1, (S, dz))), rng.normal(0, 1, (S, dz))
        thr = 0.3
        filled, homog, flag = np.empty((S, dz)), np.empty((S, dz)), np.empty((S, dz))
        det = np.zeros(S, np.int32)
        args = [np.ascontiguousarray(a) for a in (clim, mean, lo, hi, fa, fb, fc)]
        lib.k3_emulate(*[C.c_void_p(a.ctypes.data) for a in args], C.c_int(S), C.c_int(dz), C.c_int(method),
                       C.c_double(thr), C.c_double(ND), C.c_void_p(filled.ctypes.data),
                       C.c_void_p(homog.ctypes.data), C.c_void_p(flag.ctypes.data), C.c_void_p(det.ctypes.data))
        c = np.where(np.isnan(clim), mean, clim)
        hw = ~((c >= lo) & (c <= hi))
        if method == 2:
            fix = np.where(fc > thr, fa, fb)
        elif method == 3:
            fix = np.where(c > hi, fa, fb)
        else:
            fix = fa
        assert np.array_equal(filled, c)
        assert np.array_equal(homog, np.where(hw, fix, c))
        assert np.array_equal(flag, np.where(hw, c, ND))
        assert np.array_equal(det, hw.sum(axis=1))
