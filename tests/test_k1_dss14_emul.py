# -*- coding: utf-8 -*-
"""CPU check of K1's '%12.4f\\r\\n' record parser: csrc/k1_dss14.cuh (the code
the CUDA decoder inlines) compiled by g++ (tests/emul) against Python's
float() -> float32, on every kind of spelling the layout allows, on exact
float32 ties, and on malformed records (which must be REJECTED, i.e. left to
the general parsers, never mis-read)."""
import ctypes as C

import numpy as np
import pytest

from emul import build_emul


@pytest.fixture(scope='module')
def k1():
    lib = C.CDLL(build_emul.build_k1())
    lib.k1_dss14_emulate.restype = None
    lib.k1_dss14_emulate.argtypes = [C.c_void_p, C.c_longlong, C.c_void_p, C.c_void_p]

    def run(lines):
        text = np.frombuffer(''.join(s + '\r\n' for s in lines).encode('latin-1'), dtype=np.uint8)
        assert text.size == 14 * len(lines)
        out = np.zeros(len(lines), np.float32)
        ok = np.zeros(len(lines), np.uint8)
        lib.k1_dss14_emulate(text.ctypes.data, len(lines), out.ctypes.data, ok.ctypes.data)
        return out, ok.astype(bool)
    return run


def test_values_equal_python_float(k1):
    rng = np.random.default_rng(123)
    vals = np.concatenate([
        np.round(rng.uniform(-999999.0, 9999999.0, size=400000), 4),
        np.round(rng.gamma(2.0, 150.0, size=200000) + 300.0, 4),
        np.round(rng.normal(0, 1, size=100000), 4),
        8388608.0 + np.arange(0, 20000) * 0.5,           # exact ties of the float32 grid
        4194304.0 + np.arange(0, 20000) * 0.25,
        2097152.0 + np.arange(0, 20000) * 0.125,
        -(131072.0 + np.arange(0, 4000) * 0.015625),
        [0.0, -0.0, 0.0001, -0.0001, 9999999.9999, -999999.9999, 1.0, 0.5, 0.1, -999.9,
         16777215.0 / 2, 0.3333, 1.2345]])
    lines = ['%12.4f' % v for v in vals]
    assert all(len(s) == 12 for s in lines)
    out, ok = k1(lines)
    assert ok.all()
    exp = np.array([np.float32(float(s)) for s in lines], dtype=np.float32)
    assert np.array_equal(out.view(np.int32), exp.view(np.int32))


def test_every_integer_times_1e_minus_4_near_float32_boundaries(k1):
    """The multiply-by-reciprocal is guarded by the distance of the float64
    result to a float32 rounding boundary; sweep dense ranges of m = value *
    10^4, where consecutive values straddle such boundaries."""
    for start in (0, 16777216 * 10000 // 3, 83886080000 - 50000, 99999990000 - 100000):
        m = np.arange(start, start + 100000, dtype=np.int64)
        lines = ['%7d.%04d' % (v // 10000, v % 10000) for v in m]
        lines = [s for s in lines if len(s) == 12]
        out, ok = k1(lines)
        assert ok.all()
        exp = np.array([np.float32(float(s)) for s in lines], dtype=np.float32)
        assert np.array_equal(out.view(np.int32), exp.view(np.int32))


BAD = ['12a.5000', '1 2.5000', '--12.5000', '-+12.5000', '+12.5000', '1.2.5000'[:8] + '0',
       '.5000', '-.5000', '12.50e0', '', '-  12.5000', '12,5000', '12.50x0', '\t12.5000',
       '0x12.5000', '1e.5000', '12.5', '12.50000', '1234567.12']


@pytest.mark.parametrize('bad', [b.rjust(12)[:12] for b in BAD] + ['12.5000     '])
def test_malformed_records_are_rejected(k1, bad):
    assert len(bad) == 12
    out, ok = k1([bad, '     12.5000'])
    assert not ok[0]
    assert ok[1] and out[1] == np.float32(12.5)


def test_wrong_line_ending_is_rejected(k1):
    import ctypes
    lib = ctypes.CDLL(build_emul.build_k1())
    for tail in (b'\n\r', b' \n', b'\r\r', b'\n\n'):
        text = np.frombuffer(b'     12.5000' + tail, dtype=np.uint8).copy()
        out = np.zeros(1, np.float32)
        ok = np.zeros(1, np.uint8)
        lib.k1_dss14_emulate(text.ctypes.data, 1, out.ctypes.data, ok.ctypes.data)
        assert ok[0] == 0


def test_eight_records_per_thread_form_agrees_with_the_per_record_form():
    """k1_decode_dss14x8 (experimental, stack option k1_algo = 1) takes record k of
    a 112-byte group from a compile-time word / half-word position: same values,
    same accept / reject decisions as the per-record form, at every position of
    the group, for good and malformed records."""
    lib = C.CDLL(build_emul.build_k1())
    for f in (lib.k1_dss14_emulate, lib.k1_dss14_emulate_x8):
        f.restype = None
        f.argtypes = [C.c_void_p, C.c_longlong, C.c_void_p, C.c_void_p]
    rng = np.random.default_rng(9)
    vals = np.round(rng.uniform(-999999.0, 9999999.0, size=80000), 4)
    lines = ['%12.4f' % v for v in vals]
    bad = [b.rjust(12)[:12] for b in BAD]
    for i, pos in enumerate(rng.integers(0, len(lines), size=4000)):
        lines[pos] = bad[i % len(bad)]
    text = np.frombuffer(''.join(s + '\r\n' for s in lines).encode('latin-1'), dtype=np.uint8).copy()
    # wrong line endings at every position of a group
    for k in range(8):
        text[14 * (800 + 9 * k) + 12: 14 * (800 + 9 * k) + 14] = (10, 13)
    n = len(lines)
    o1, k1_ = np.zeros(n, np.float32), np.zeros(n, np.uint8)
    o2, k2_ = np.zeros(n, np.float32), np.zeros(n, np.uint8)
    lib.k1_dss14_emulate(text.ctypes.data, n, o1.ctypes.data, k1_.ctypes.data)
    lib.k1_dss14_emulate_x8(text.ctypes.data, n, o2.ctypes.data, k2_.ctypes.data)
    assert np.array_equal(k1_, k2_)
    assert 3000 < int((k1_ == 0).sum()) < 4100
    good = k1_ == 1
    assert np.array_equal(o1[good].view(np.int32), o2[good].view(np.int32))
    exp = np.array([np.float32(float(s)) for s, g in zip(lines, good) if g], dtype=np.float32)
    assert np.array_equal(o2[good].view(np.int32), exp.view(np.int32))
