# -*- coding: utf-8 -*-
"""TEST INFRASTRUCTURE: differential fuzzing of the K2 counting-sort kernel run on
the CPU (tests/emul/k2_csort_threads.cpp) against the oracle.

    python tests/emul/fuzz_csort.py [seed] [seconds]

Random stacks: R from 33 to 1024, 32-64 columns drawn from eight families
(lattice normals, gammas, clamped normals = clamp atoms, few-valued = heavy ties,
zero-inflated, values centred on zero over 13 decades, constants, heavy tails)
with six missing-value patterns (leading, trailing, random, every second, all).
Gates: medians, percentile pair, 19 percentiles and the mode bit for bit; variance,
std and skewness at the GPU suite's tolerances; the mean to 1e-5 relative OR 2e-7
sigma (a float32 quotient cannot do better on a mean that is ~0 against its sigma:
the one class of deviations the fuzzing found, 1.5e-7 sigma with the first pivot
rule, below 5e-8 sigma since).
Recorded runs (round 2): seed 7, 20 minutes, 4 251 stacks (~2*10^5 columns): no
order-statistic, mode, variance or skewness deviation; seed 99, 30 minutes, 6 876
stacks: one skewness 1.8e-5 off (a column whose first 127 realizations are
missing: the pivot group held a single, tail value) -> the pivot group must hold 8
values; seed 123 after that, 20 minutes, 4 264 stacks: clean."""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

ND = -999.9
QS = [round(0.05 * k, 2) for k in range(1, 20)]


def column(rng, R, kind):
    if kind == 0:
        x = np.round(rng.normal(rng.uniform(-1000, 1000), rng.uniform(0.1, 300), R) * 16) / 16
    elif kind == 1:
        x = np.round(rng.gamma(rng.uniform(0.5, 5), rng.uniform(1, 100), R), 4)
    elif kind == 2:
        x = np.clip(rng.normal(0, 1, R), rng.uniform(-3, 0), rng.uniform(0, 3)) * rng.uniform(1, 500) \
            + rng.uniform(-500, 500)
    elif kind == 3:
        x = rng.integers(0, rng.integers(2, 40), R) * rng.choice([0.5, 0.125, 1.0, 3.0]) + rng.uniform(-50, 50)
    elif kind == 4:
        x = np.where(rng.uniform(size=R) < rng.uniform(0.1, 0.9), rng.uniform(-5, 5),
                     np.round(rng.exponential(rng.uniform(1, 80), R) * 16) / 16)
    elif kind == 5:
        x = rng.uniform(-1, 1, R) * 10.0 ** rng.integers(-6, 7) + rng.uniform(-1, 1) * 10.0 ** rng.integers(-3, 8)
    elif kind == 6:
        x = np.full(R, rng.uniform(-100, 100))
    else:
        x = np.round(rng.standard_t(2.5, R) * rng.uniform(1, 50) * 16) / 16 + rng.uniform(0, 2000)
    x = x.astype(np.float32).astype(np.float64)
    pat = rng.integers(0, 6)
    if pat == 1:
        x[:rng.integers(1, R)] = ND
    elif pat == 2:
        x[rng.integers(1, R):] = ND
    elif pat == 3:
        x[rng.uniform(size=R) < rng.uniform(0, 0.9)] = ND
    elif pat == 4:
        x[::2] = ND
    elif pat == 5 and rng.uniform() < 0.3:
        x[:] = ND
    return x


def random_stack(rng):
    R = int(rng.choice([33, 40, 64, 65, 100, 127, 128, 129, 200, 256, 257, 333, 500, 512, 513, 777,
                        1000, 1023, 1024]))
    N = int(rng.choice([32, 45, 64]))
    st = np.stack([column(rng, R, int(rng.integers(0, 8))) for _ in range(N)], axis=1).astype(np.float32)
    return st, float(rng.choice([0.95, 0.9, 0.5, 0.99]))


def check(st, p):
    import test_k2_csort_emul as T
    from oracle import np_oracle as O
    out = T.csort_on_cpu(st, p=p)
    ref = O.grid_stats(st, ND, 127 | O.FLAG_MODE, p, QS)

    def f32(a):
        return np.asarray(a, np.float64).astype(np.float32)

    def bits(a, b, what):
        a, b = f32(a), f32(b)
        ok = (a == b) | (np.isnan(a) & np.isnan(b))
        assert ok.all(), (what, np.argwhere(~ok)[:3].tolist(), a[~ok][:3], b[~ok][:3])
    bits(out[1], ref['medianmap'], 'median')
    bits(out[6], ref['percmap'][:, 0], 'lower percentile')
    bits(out[7], ref['percmap'][:, 1], 'upper percentile')
    for i in range(len(QS)):
        bits(out[8 + i], ref['percentiles'][:, i], 'percentile %g' % QS[i])
    bits(out[8 + len(QS)], ref['modemap'], 'mode')
    sd = np.nan_to_num(np.asarray(ref['stdmap'], float), nan=0.0, posinf=0.0)

    def close(a, b, rtol, atol, what):
        a, b = np.asarray(a, float), np.asarray(b, float)
        assert np.array_equal(np.isnan(a), np.isnan(b)), (what, 'NaN pattern')
        with np.errstate(all='ignore'):
            bad = (np.abs(a - b) > rtol * np.abs(b) + atol) & ~np.isnan(b)
        assert not bad.any(), (what, np.argwhere(bad)[:3].tolist(), a[bad][:3], b[bad][:3])
    close(out[0], ref['meanmap'], 1e-5, 2e-7 * sd, 'mean')
    close(out[3], ref['varmap'], 1e-5, 1e-8 * sd ** 2, 'variance')
    close(out[4], ref['stdmap'], 1e-5, 1e-6 * sd, 'std')
    close(out[2], ref['skewmap'], 1e-5, 1e-5, 'skewness')


def run(seed, seconds=None, stacks=None):
    rng = np.random.default_rng(seed)
    t_end = time.time() + seconds if seconds else None
    runs = fails = 0
    while (t_end is None or time.time() < t_end) and (stacks is None or runs < stacks):
        st, p = random_stack(rng)
        runs += 1
        try:
            check(st, p)
        except AssertionError as exc:
            fails += 1
            print('FAIL stack', runs, st.shape, 'p', p, str(exc)[:400], flush=True)
    return runs, fails


if __name__ == '__main__':
    seed = int(sys.argv[1]) if len(sys.argv) > 1 else 0
    secs = float(sys.argv[2]) if len(sys.argv) > 2 else 300.0
    r, f = run(seed, seconds=secs)
    print('stacks', r, 'failures', f)
