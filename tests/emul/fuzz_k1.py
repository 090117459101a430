# -*- coding: utf-8 -*-
"""TEST INFRASTRUCTURE: fuzzing of K1's general number parsers (csrc/k1_parse.cuh
compiled for the host, tests/emul/k1_parse_emul.cpp) against Python's float().

    python tests/emul/fuzz_k1.py [seed] [seconds]

Random spellings: signs, 0-22 integer and 0-21 fraction digits, leading zeros,
trailing zeros, exponents from e0 to e300 with and without sign, blanks / tabs /
carriage returns around, inf / nan words and their truncations, random
corruption.  Gates: an accepted line must be a line Python accepts AND its float32
must equal ``np.float32(float(line))`` bit for bit; a refused line that Python
accepts must belong to the documented classes (more than 19 significant digits, a
decimal exponent outside the exact-arithmetic window: the C ABI reports those as
GSB_EPARSE, never as a guess).
Recorded run (round 2): seed 1, 120 s, 1 160 000 spellings through both entry
points (parse_float alone / fast path + parse_float as the kernel combines them):
no wrong value, no invalid line accepted."""
import ctypes as C
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

WS = [' ', '\t', '\r', ' ', ' ']
DIGITS = list('0123456789')


def spelling(rng):
    if rng.integers(0, 100) < 2:
        return str(rng.choice(['inf', '-inf', '+Infinity', 'nan', 'NaN', '-nan', 'INF', 'infinit', 'na', 'in']))
    sign = rng.choice(['', '', '-', '+'])
    ni = int(rng.choice([0, 1, 1, 2, 3, 5, 7, 9, 12, 15, 16, 17, 18, 19, 20, 22]))
    nf = int(rng.choice([0, 0, 1, 2, 4, 4, 6, 8, 10, 15, 16, 17, 19, 21]))
    lead0 = '0' * int(rng.choice([0, 0, 0, 1, 3]))
    ip = ''.join(rng.choice(DIGITS, ni)) if ni else ''
    fp = ''.join(rng.choice(DIGITS, nf)) if nf else ''
    if rng.integers(0, 6) == 0 and nf:
        fp = fp[:-min(3, nf)] + '0' * min(3, nf)
    dot = '.' if (nf or rng.integers(0, 4) == 0) else ''
    ex = ''
    if rng.integers(0, 3) == 0:
        ex = rng.choice(['e', 'E']) + rng.choice(['', '+', '-']) + \
            str(int(rng.choice([0, 1, 2, 5, 10, 15, 20, 22, 23, 27, 30, 38, 45, 300])))
    s = sign + lead0 + ip + dot + fp + ex
    if rng.integers(0, 25) == 0:
        k = rng.integers(0, len(s) + 1)
        s = s[:k] + rng.choice(list('x-+. e_,')) + s[k:]
    return ''.join(rng.choice(WS, rng.integers(0, 4))) + s + ''.join(rng.choice(WS, rng.integers(0, 3)))


def documented_refusal(s):
    """True when a line Python accepts may be refused: > 19 significant digits or a
    decimal exponent outside the window of the exact paths."""
    t = s.strip().lstrip('+-').lower()
    mant, _, ex = t.partition('e')
    ip, _, fp = mant.partition('.')
    digits = (ip + fp).lstrip('0')
    sig = len(digits.rstrip('0')) if digits else 0
    if sig > 19 or len(digits) > 19:
        return True
    e10 = (int(ex) if ex else 0) - len(fp)
    return not (-40 <= e10 <= 27) or e10 < -21 or len(digits) + e10 > 38


def run(lib, rng, which, n):
    lines = [spelling(rng) for _ in range(n)]
    blob = '\n'.join(lines).encode('latin-1') + b'\n'
    off = np.zeros(2 * n, np.int64)
    pos = 0
    for i, s in enumerate(lines):
        off[2 * i] = pos
        off[2 * i + 1] = pos + len(s)
        pos += len(s) + 1
    text = np.frombuffer(blob + b' ' * 64, dtype=np.uint8).copy()
    out = np.zeros(n, np.float32)
    ok = np.zeros(n, np.uint8)
    lib.k1_parse_emulate(C.c_void_p(text.ctypes.data), C.c_longlong(len(blob)), C.c_void_p(off.ctypes.data),
                         C.c_longlong(n), C.c_int(which), C.c_void_p(out.ctypes.data), C.c_void_p(ok.ctypes.data))
    problems = []
    accepted = 0
    for i, s in enumerate(lines):
        try:
            if '_' in s:                      # Python >= 3.6 digit grouping: not the reference's Python 2
                raise ValueError
            with np.errstate(over='ignore'):
                exp = np.float32(float(s))
            valid = True
        except (ValueError, OverflowError):
            valid = False
        if ok[i]:
            accepted += 1
            if not valid:
                problems.append(('accepted an invalid line', s, float(out[i])))
            elif not ((np.isnan(exp) and np.isnan(out[i])) or exp.view(np.int32) == out[i:i + 1].view(np.int32)[0]):
                problems.append(('wrong value', s, float(out[i]), float(exp)))
        elif valid and not documented_refusal(s):
            problems.append(('refused a supported line', s))
    return accepted, problems


def load():
    from emul import build_emul
    lib = C.CDLL(build_emul.build_k1_parse())
    lib.k1_parse_emulate.restype = None
    return lib


if __name__ == '__main__':
    seed = int(sys.argv[1]) if len(sys.argv) > 1 else 0
    secs = float(sys.argv[2]) if len(sys.argv) > 2 else 60.0
    lib_, rng_ = load(), np.random.default_rng(seed)
    t_end, total, bad = time.time() + secs, 0, []
    while time.time() < t_end:
        for which_ in (0, 1):
            _, pr = run(lib_, rng_, which_, 20000)
            total += 20000
            bad += pr
    print('spellings', total, 'problems', len(bad))
    for b in bad[:20]:
        print(b)
