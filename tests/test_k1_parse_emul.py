# -*- coding: utf-8 -*-
"""CPU fuzz of K1's general number parsers (csrc/k1_parse.cuh, the code the CUDA
decoder compiles -- moving it into the header left the SASS byte-identical)
against Python's float(): see tests/emul/fuzz_k1.py."""
import numpy as np

from emul import fuzz_k1


def test_general_parsers_against_python_float():
    lib = fuzz_k1.load()
    rng = np.random.default_rng(2026)
    for which in (0, 1):
        accepted, problems = fuzz_k1.run(lib, rng, which, 40000)
        assert accepted > 10000
        assert not problems, problems[:10]


def test_known_spellings():
    lib = fuzz_k1.load()
    import ctypes as C
    lines = ['1', ' 1.5 ', '-0.0', '1e3', '1E-3', '.5', '5.', '+7', '  12.5000\r', '0001.2500', '1e', 'e1', '.', '',
             '--1', '1.2.3', '12,5', 'inf', '-Infinity', 'nan', '123456789012345678', '0.1234567890123456789',
             '9007199254740993', '1.0000000596046448', '16777217', '1e22', '1e23', '123456789012345e-37']
    blob = '\n'.join(lines).encode('ascii') + b'\n'
    off = np.zeros(2 * len(lines), np.int64)
    pos = 0
    for i, s in enumerate(lines):
        off[2 * i], off[2 * i + 1] = pos, pos + len(s)
        pos += len(s) + 1
    text = np.frombuffer(blob + b' ' * 64, dtype=np.uint8).copy()
    out = np.zeros(len(lines), np.float32)
    ok = np.zeros(len(lines), np.uint8)
    lib.k1_parse_emulate(C.c_void_p(text.ctypes.data), C.c_longlong(len(blob)), C.c_void_p(off.ctypes.data),
                         C.c_longlong(len(lines)), C.c_int(0), C.c_void_p(out.ctypes.data), C.c_void_p(ok.ctypes.data))
    for i, s in enumerate(lines):
        try:
            exp = np.float32(float(s))
        except ValueError:
            assert not ok[i], s
            continue
        if fuzz_k1.documented_refusal(s):
            continue
        assert ok[i], s
        assert (np.isnan(exp) and np.isnan(out[i])) or exp.view(np.int32) == out[i:i + 1].view(np.int32)[0], s
