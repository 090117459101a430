# -*- coding: utf-8 -*-
"""CPU check of the K2 staged-select kernel's algorithm (experimental kernel,
stack option ``k2_algo = 4``).

``gsimcli_b200/csrc/k2_ssel_phases.cuh`` holds the per-thread phases the CUDA
kernel inlines; ``tests/emul`` compiles the same header with g++ and runs the
phases thread by thread with the kernel's barriers as loop boundaries and the
cp.async tile copy as a plain copy.  These tests compare that emulation with
the oracle for the call shape the kernel serves -- the reference's
``stats(lmean, lmed, lskew, lperc)`` (grid.py:601-602) plus the other moment
planes: bit-exact order statistics, 1e-5 moments -- and with the emulation of
the register-tile kernel (the default for that shape), with which it must agree
bit for bit on every plane.
"""
import ctypes as C

import numpy as np
import pytest

from gsimcli_b200 import synth
from oracle import np_oracle as O
from emul import build_emul

ND = -999.9
ARGT = [C.c_void_p, C.c_longlong, C.c_longlong, C.c_int, C.c_float, C.c_int, C.c_void_p,
        C.c_void_p, C.c_void_p, C.c_longlong, C.c_int, C.c_void_p]
# mean, median, skew, var, std, coefvar, percentile pair
KIND = np.array([0, 1, 2, 3, 4, 5, 1, 1], np.uint8)


@pytest.fixture(scope='module')
def libs():
    ss = C.CDLL(build_emul.build_ssel())
    ss.k2_ssel_emulate.restype = C.c_int
    ss.k2_ssel_emulate.argtypes = ARGT
    rs = C.CDLL(build_emul.build())
    rs.k2_rsel_emulate.restype = C.c_int
    rs.k2_rsel_emulate.argtypes = ARGT
    return ss, rs


def padded(st):
    """[R, pitch] with the pitch a multiple of 32 floats, as gsb_stack holds it;
    the padding columns carry garbage that must never reach an output."""
    st = np.ascontiguousarray(st, np.float32)
    R, N = st.shape
    pitch = (N + 31) // 32 * 32
    buf = np.full((R, pitch), 1.0e30, np.float32)
    buf[:, :N] = st
    return buf, N


def run(fn, st, nw, p=0.95, kind=KIND):
    buf, N = padded(st)
    R, pitch = buf.shape
    q = np.array([0, .5, 0, 0, 0, 0, (1 - p) / 2, 1 - (1 - p) / 2], np.float64)[:len(kind)]
    out = np.full((len(kind), N), -12345.0, np.float32)
    fl = np.zeros(N // 32 + 2, np.int32)
    nf = fn(buf.ctypes.data, pitch, N, R, ND, len(kind), kind.ctypes.data, q.ctypes.data,
            out.ctypes.data, N, nw, fl.ctypes.data)
    assert nf >= 0
    return out, fl[:nf]


def check(libs, st, nw, p=0.95, expect_flagged=None):
    ss, rs = libs
    st = np.ascontiguousarray(st, np.float32)
    out, fl = run(ss.k2_ssel_emulate, st, nw, p)
    cols = np.ones(st.shape[1], bool)
    for t in fl:                      # flagged tiles belong to the counting-sort kernel
        cols[t * 32:(t + 1) * 32] = False
    if expect_flagged is not None:
        assert sorted(fl.tolist()) == sorted(expect_flagged)
    ref = O.grid_stats(st, ND, 255 & ~O.FLAG_MODE, p, None)

    def bits(a, b):
        a = np.asarray(a, np.float64).astype(np.float32)[cols]
        b = np.asarray(b, np.float64).astype(np.float32)[cols]
        assert ((a == b) | (np.isnan(a) & np.isnan(b))).all()

    def close(a, b, rtol=1e-5, atol=0.0):
        a, b = np.asarray(a, float)[cols], np.asarray(b, float)[cols]
        assert np.array_equal(np.isnan(a), np.isnan(b))
        np.testing.assert_allclose(a, b, rtol=rtol, atol=atol, equal_nan=True)
    sigma = np.nanmax(ref['stdmap']) if np.isfinite(ref['stdmap']).any() else 1.0
    # (float32 data: a mean within 1e-6 sigma of zero has no relative digits to compare)
    close(out[0], ref['meanmap'], atol=1e-6 * sigma)
    bits(out[1], ref['medianmap'])
    close(out[2], ref['skewmap'], atol=1e-5)
    close(out[3], ref['varmap'], atol=1e-8 * sigma ** 2)
    close(out[4], ref['stdmap'], atol=1e-6 * sigma)
    close(out[5], ref['coefvarmap'], atol=1e-6)
    bits(out[6], ref['percmap'][:, 0])
    bits(out[7], ref['percmap'][:, 1])
    # the register-tile kernel (same pivot, same partial sums, same finals):
    # every plane of every tile neither kernel flagged agrees bit for bit
    out2, fl2 = run(rs.k2_rsel_emulate, st, nw, p)
    both = cols.copy()
    for t in fl2:
        both[t * 32:(t + 1) * 32] = False
    a, b = out[:, both].view(np.uint32), out2[:, both].view(np.uint32)
    nan = np.isnan(out[:, both]) & np.isnan(out2[:, both])
    assert ((a == b) | nan).all()
    assert (out != np.float32(-12345.0)).all()
    return fl


@pytest.mark.parametrize('nw', [16, 32])
@pytest.mark.parametrize('R', [33, 64, 100, 200, 256, 257, 500, 513, 1000, 1024])
def test_emulated_kernel_matches_oracle(libs, nw, R):
    st = synth.stack(2000 + R, R, [13, 7, 5])
    st[:, 3] = 812.5                                  # constant column
    st[:, 4] = np.float32(ND)                         # empty column
    st[:, 5] = np.float32(ND)
    st[R // 2, 5] = 640.25                            # n == 1
    st[:, 6] = np.float32(ND)
    st[0, 6], st[R - 1, 6] = 7.5, 8.5                 # n == 2
    st[:, 8] = np.where(np.arange(R) % 2 == 0, 500.0, 600.0)    # two atoms
    st[:, 9] = np.float32(ND)
    st[1, 9], st[5, 9], st[R - 2, 9] = 3.0, 1.0, 2.0  # n == 3
    check(libs, st, nw, expect_flagged=[])


@pytest.mark.parametrize('nw', [16, 32])
@pytest.mark.parametrize('p', [0.95, 0.9, 0.5, 0.999, 1.0, 0.0])
def test_emulated_kernel_probabilities(libs, nw, p):
    """detection probabilities incl. the degenerate ones: the percentile pair
    reaches the column minimum / maximum (p = 1) or collapses on the median."""
    st = synth.stack(77, 1000, [16, 8, 1])
    check(libs, st, nw, p=p, expect_flagged=[])


@pytest.mark.parametrize('nw', [16, 32])
def test_emulated_kernel_distributions(libs, nw):
    rng = np.random.default_rng(17)
    for R in (200, 1000):
        N = 96
        st = np.round(rng.normal(800.0, 150.0, size=(R, N)) * 16) / 16
        st[:, 8:16] = np.round(rng.gamma(2.0, 3.0, size=(R, 8)), 4)
        st[:, 16:24] = np.clip(rng.normal(0, 1, size=(R, 8)), -0.5, 1.0)       # clamp atoms
        st[:, 24:32] = 1.0e7 + np.round(rng.uniform(0, 8, size=(R, 8)) * 16) / 16
        st[:, 32:40] = np.round(rng.exponential(50.0, size=(R, 8)) * 16) / 16 - 200.0
        st[:, 40:48] = np.where(rng.uniform(size=(R, 8)) < 0.7, 0.0,
                                np.round(rng.gamma(2.0, 20.0, size=(R, 8)) * 16) / 16)
        st[:, 48:52] = np.clip(rng.normal(0, 1, size=(R, 4)), -3.0, -1.7)      # 95 % of the column is the maximum
        st[:, 52:56] = np.clip(rng.normal(0, 1, size=(R, 4)), 1.7, 3.0)        # ... the minimum
        st[:40, 56:64] = ND                            # first realizations all missing
        st[:, 64:72] = -np.abs(st[:, 64:72])
        st[rng.uniform(size=st.shape) < 0.01] = ND
        check(libs, st, nw)


def test_emulated_kernel_flags_what_it_cannot_finish(libs):
    """An outlier that collapses the bin map, heavy ties in a wanted bin and a
    range that overflows float32 raise the tile's flag (the CUDA path redoes
    those tiles with the counting-sort kernel); every other tile stays exact."""
    rng = np.random.default_rng(3)
    R = 1000
    st = rng.normal(10, 1, size=(R, 160)).astype(np.float32)
    st[3, 0:4] = 1.0e6                                  # tile 0: outlier
    st[:, 40] = np.float32(3.4e38)
    st[::2, 40] = np.float32(-3.4e38)                   # tile 1: range overflow
    st[:, 70] = np.where(rng.uniform(size=R) < 0.5, 12.0, 12.0 + 1e-6 * rng.integers(0, 3, size=R))
    st[0, 70], st[1, 70] = -40.0, 60.0                  # tile 2: heavy near-ties in one bin
    fl = check(libs, st, 16)
    assert sorted(fl.tolist()) == [0, 1, 2]


def test_emulated_kernel_list_lengths(libs):
    """Wanted bins of 9..32 values take the 16- and 32-value networks; 33 and
    more overflow the list (flag).  The column is built so that the median's bin
    holds exactly m distinct values."""
    R = 1000
    cols = []
    for m in (8, 9, 16, 17, 31, 32, 33, 40):
        x = np.empty(R, np.float32)
        half = (R - m) // 2
        x[:half] = np.linspace(0.0, 100.0, half)
        x[half:half + m] = 256.0 + np.arange(m) * 1e-4          # one bin (width 1): the median's
        x[half + m:] = np.linspace(412.0, 512.0, R - half - m)
        cols.append(np.random.default_rng(m).permutation(x))
    st = np.zeros((R, 32 * len(cols)), np.float32)
    st[:] = np.linspace(0, 512, R, dtype=np.float32)[:, None]
    for k, x in enumerate(cols):
        st[:, 32 * k + 5] = x
    fl = check(libs, st, 16)
    assert sorted(fl.tolist()) == [6, 7]


def test_emulation_runs_clean_under_sanitizers():
    """AddressSanitizer + UBSan over the emulation (the shared-memory arrays have
    their exact sizes, so an out-of-range index of a phase is an error)."""
    import os
    import subprocess
    import sys
    import glob
    asan = sorted(glob.glob('/usr/lib/gcc/x86_64-linux-gnu/*/libasan.so')
                  + glob.glob('/usr/lib/x86_64-linux-gnu/libasan.so.*'))
    if not asan:
        pytest.skip('libasan not installed')
    lib = build_emul.build_ssel(sanitize=True)
    code = r'''
import ctypes as C, numpy as np, sys
sys.path.insert(0, %r)
from gsimcli_b200 import synth
lib = C.CDLL(%r)
lib.k2_ssel_emulate.restype = C.c_int
for nw in (16, 32):
    for R in (40, 200, 1000, 1024):
        st = synth.stack(5, R, [16, 5, 1])
        st[:, 7] = -999.9
        buf = np.full((R, 96), 1e30, np.float32); buf[:, :80] = st
        kind = np.array([0, 1, 2, 1, 1], np.uint8)
        q = np.array([0, .5, 0, .025, .975])
        out = np.zeros((5, 80), np.float32); fl = np.zeros(8, np.int32)
        nf = lib.k2_ssel_emulate(C.c_void_p(buf.ctypes.data), C.c_longlong(96), C.c_longlong(80), C.c_int(R),
                                 C.c_float(-999.9), C.c_int(5), C.c_void_p(kind.ctypes.data),
                                 C.c_void_p(q.ctypes.data), C.c_void_p(out.ctypes.data), C.c_longlong(80),
                                 C.c_int(nw), C.c_void_p(fl.ctypes.data))
        assert nf == 0, nf
print('clean')
''' % (os.path.dirname(os.path.dirname(os.path.abspath(__file__))), lib)
    env = dict(os.environ, LD_PRELOAD=asan[-1], ASAN_OPTIONS='detect_leaks=0')
    r = subprocess.run([sys.executable, '-c', code], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and 'clean' in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]


@pytest.mark.parametrize('sanitizer', [None, 'thread', 'address'])
def test_kernel_function_with_one_thread_per_cuda_thread(sanitizer):
    """The kernel function itself (csrc/k2_ssel_kernel.cuh: tile loop, barriers,
    cp.async staging of the next tile, shared-memory layout, sub-range and output
    addressing) compiled for the host and run with one OS thread per CUDA thread
    and a real barrier, the asynchronous copy performed as early and as late as
    the hardware may: planes bit-identical to the serial emulation, and no data
    race (ThreadSanitizer) / no out-of-bounds access of the exact-size shared
    block (AddressSanitizer).  Removing one barrier makes this test fail (tried:
    3 races reported, 2394 mismatching values)."""
    import subprocess
    exe = build_emul.build_ssel_threads(sanitizer)
    args = [exe] + (['quick'] if sanitizer == 'thread' else [])
    env = dict(__import__('os').environ, TSAN_OPTIONS='halt_on_error=0 exitcode=66',
               ASAN_OPTIONS='detect_leaks=0')
    r = subprocess.run(args, capture_output=True, text=True, timeout=900, env=env)
    assert r.returncode == 0 and 'clean' in r.stdout, (r.stdout[-1500:], r.stderr[-3000:])
    assert 'ThreadSanitizer' not in r.stderr and 'AddressSanitizer' not in r.stderr
