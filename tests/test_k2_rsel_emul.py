# -*- coding: utf-8 -*-
"""CPU check of the K2 register-tile kernel's algorithm.

``gsimcli_b200/csrc/k2_rsel_phases.cuh`` holds the per-thread phases the CUDA
kernel inlines; ``tests/emul`` compiles the same header with g++ and runs the
phases thread by thread with the kernel's barriers as loop boundaries.  These
tests compare that emulation with the oracle: bit-exact order statistics and
mode, 1e-5 moments -- the same gates as the GPU parity suite.  (The CUDA kernel
itself is checked on the GPU by tests/test_gpu_parity.py.)
"""
import ctypes as C

import numpy as np
import pytest

from gsimcli_b200 import synth
from oracle import np_oracle as O
from emul import build_emul

ND = -999.9
QS = [round(0.05 * k, 2) for k in range(1, 20)]


@pytest.fixture(scope='module')
def emul_lib():
    lib = C.CDLL(build_emul.build())
    lib.k2_rsel_emulate.restype = C.c_int
    lib.k2_rsel_emulate.argtypes = [C.c_void_p, C.c_longlong, C.c_longlong, C.c_int, C.c_float,
                                    C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_longlong,
                                    C.c_int, C.c_void_p]
    return lib


def run_emul(lib, st, nw, p=0.95, qs=QS):
    st = np.ascontiguousarray(st, np.float32)
    R, N = st.shape
    kind = np.array([0, 1, 2, 3, 4, 5, 1, 1] + [1] * len(qs) + [6], np.uint8)
    q = np.array([0, .5, 0, 0, 0, 0, (1 - p) / 2, 1 - (1 - p) / 2] + list(qs) + [0], np.float64)
    out = np.full((len(kind), N), -12345.0, np.float32)
    fl = np.zeros(N // 32 + 2, np.int32)
    nf = lib.k2_rsel_emulate(st.ctypes.data, N, N, R, ND, len(kind), kind.ctypes.data,
                             q.ctypes.data, out.ctypes.data, N, nw, fl.ctypes.data)
    assert nf >= 0
    return out, fl[:nf]


def check(lib, st, nw, expect_flagged=None):
    st = np.ascontiguousarray(st, np.float32)
    out, fl = run_emul(lib, st, nw)
    cols = np.ones(st.shape[1], bool)
    for t in fl:                      # flagged tiles belong to the counting-sort kernel
        cols[t * 32:(t + 1) * 32] = False
    if expect_flagged is not None:
        assert sorted(fl.tolist()) == sorted(expect_flagged)
    flags = 255
    ref = O.grid_stats(st, ND, flags, 0.95, QS)

    def bits(a, b):
        a = np.asarray(a, np.float64).astype(np.float32)[cols]
        b = np.asarray(b, np.float64).astype(np.float32)[cols]
        assert ((a == b) | (np.isnan(a) & np.isnan(b))).all()

    def close(a, b, rtol=1e-5, atol=0.0):
        a, b = np.asarray(a, float)[cols], np.asarray(b, float)[cols]
        assert np.array_equal(np.isnan(a), np.isnan(b))
        np.testing.assert_allclose(a, b, rtol=rtol, atol=atol, equal_nan=True)
    sigma = np.nanmax(ref['stdmap']) if np.isfinite(ref['stdmap']).any() else 1.0
    close(out[0], ref['meanmap'])
    bits(out[1], ref['medianmap'])
    close(out[2], ref['skewmap'], atol=1e-5)
    close(out[3], ref['varmap'], atol=1e-8 * sigma ** 2)
    close(out[4], ref['stdmap'], atol=1e-6 * sigma)
    close(out[5], ref['coefvarmap'], atol=1e-6)
    bits(out[6], ref['percmap'][:, 0])
    bits(out[7], ref['percmap'][:, 1])
    for i in range(len(QS)):
        bits(out[8 + i], ref['percentiles'][:, i])
    bits(out[8 + len(QS)], ref['modemap'])
    return fl


@pytest.mark.parametrize('nw', [32, 16])
@pytest.mark.parametrize('R', [33, 64, 65, 100, 200, 256, 500, 513, 1000, 1024])
def test_emulated_kernel_matches_oracle(emul_lib, nw, R):
    st = synth.stack(1000 + R, R, [13, 7, 5])
    st[:, 3] = 812.5                                  # constant column
    st[:, 4] = np.float32(ND)                         # empty column
    st[:, 5] = np.float32(ND)
    st[R // 2, 5] = 640.25                            # n == 1
    st[:, 6] = np.float32(ND)
    st[0, 6], st[R - 1, 6] = 7.5, 8.5                 # n == 2
    st[:, 8] = np.where(np.arange(R) % 2 == 0, 500.0, 600.0)    # two atoms
    check(emul_lib, st, nw, expect_flagged=[])


@pytest.mark.parametrize('nw', [32, 16])
def test_emulated_kernel_distributions(emul_lib, nw):
    rng = np.random.default_rng(17)
    for R in (100, 1000):
        N = 96
        st = np.round(rng.normal(800.0, 150.0, size=(R, N)) * 16) / 16
        st[:, 8:16] = np.round(rng.gamma(2.0, 3.0, size=(R, 8)), 4)
        st[:, 16:24] = np.clip(rng.normal(0, 1, size=(R, 8)), -0.5, 1.0)       # clamp atoms
        st[:, 24:32] = 1.0e7 + np.round(rng.uniform(0, 8, size=(R, 8)) * 16) / 16
        st[:, 32:40] = np.round(rng.exponential(50.0, size=(R, 8)) * 16) / 16 - 200.0
        st[:, 40:48] = np.where(rng.uniform(size=(R, 8)) < 0.7, 0.0,
                                np.round(rng.gamma(2.0, 20.0, size=(R, 8)) * 16) / 16)
        st[:40, 56:64] = ND                            # first realizations all missing
        st[:, 64:72] = -np.abs(st[:, 64:72])
        st[rng.uniform(size=st.shape) < 0.01] = ND
        check(emul_lib, st, nw)


def test_emulated_kernel_flags_what_it_cannot_finish(emul_lib):
    """An outlier that collapses the bin map, heavy ties in one bin and a range
    that overflows float32 raise the tile's flag (the CUDA path redoes those
    tiles with the counting-sort kernel); every other tile stays exact."""
    rng = np.random.default_rng(3)
    R = 1000
    st = rng.normal(0, 1, size=(R, 128)).astype(np.float32)
    st[3, 0:4] = 1.0e6                                  # tile 0: outlier
    st[:, 40] = np.float32(3.4e38)
    st[::2, 40] = np.float32(-3.4e38)                   # tile 1: range overflow
    st[:, 70] = np.where(rng.uniform(size=R) < 0.5, 2.0, 2.0 + 1e-6 * rng.integers(0, 3, size=R))
    st[0, 70], st[1, 70] = -50.0, 50.0                  # tile 2: heavy near-ties in one bin
    fl = check(emul_lib, st, 32)
    assert sorted(fl.tolist()) == [0, 1, 2]


@pytest.mark.parametrize('sanitizer', [None, 'thread'])
def test_kernel_function_with_one_thread_per_cuda_thread(sanitizer):
    """The kernel function itself (csrc/k2_rsel_kernel.cuh: tile loop, barriers,
    register prefetch of the next tile, the sort buffer that aliases the partial
    sums, flagged-tile list) compiled for the host and run with one OS thread per
    CUDA thread (tests/emul/cuda_threads_shim.h): planes bit-identical to the
    serial emulation, same flagged tiles, no data race under ThreadSanitizer."""
    import os
    import subprocess
    if not os.path.exists(os.path.join(os.environ.get('CUDA_HOME', '/usr/local/cuda'), 'include', 'cuda.h')):
        pytest.skip('CUDA headers not installed')
    exe = build_emul.build_rsel_threads(sanitizer)
    env = dict(os.environ, TSAN_OPTIONS='halt_on_error=0 exitcode=66')
    r = subprocess.run([exe], capture_output=True, text=True, timeout=1800, env=env)
    assert r.returncode == 0 and 'clean' in r.stdout, (r.stdout[-2000:], r.stderr[-3000:])
    assert 'ThreadSanitizer' not in r.stderr
