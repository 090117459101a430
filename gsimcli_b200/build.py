# -*- coding: utf-8 -*-
"""In-tree build of libgsimcli_b200.so for sm_100a (nvcc, no GPU needed).

    python -m gsimcli_b200.build [--force] [--verbose]

One translation unit per kernel family, compiled in parallel, linked into
``gsimcli_b200/libgsimcli_b200.so`` next to this file so that the built library
travels with the source tree.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
OBJ = os.path.join(HERE, 'csrc', 'build')
LIB = os.path.join(HERE, 'libgsimcli_b200.so')
SOURCES = ['gsb_capi.cu', 'k1_decode.cu', 'k2_grid_sort.cu', 'k2_grid_csort.cu', 'k2_grid_rsel.cu', 'k2_grid_ssel.cu', 'k2_moments.cu',
           'k2_columns.cu', 'gsb_synth.cu']
NVCC = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-std=c++17',
         '-lineinfo', '-Xcompiler', '-fPIC', '--expt-relaxed-constexpr']


def _deps():
    hdrs = [os.path.join(CSRC, f) for f in os.listdir(CSRC)
            if f.endswith(('.cuh', '.h'))]
    hdrs.append(os.path.join(HERE, '..', 'include', 'gsimcli_b200.h'))
    return hdrs


def _stale(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    hdrs = _deps()
    jobs = []
    for src in SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(OBJ, src.replace('.cu', '.o'))
        if force or _stale(o, [s] + hdrs):
            cmd = [NVCC] + FLAGS + (['-Xptxas', '-v'] if verbose else []) + \
                ['-c', s, '-o', o]
            jobs.append(cmd)

    def run(cmd):
        r = subprocess.run(cmd, capture_output=True, text=True)
        return cmd, r

    with ThreadPoolExecutor(max_workers=min(8, max(1, len(jobs)))) as ex:
        for cmd, r in ex.map(run, jobs):
            if verbose or r.returncode:
                sys.stderr.write(' '.join(cmd) + '\n' + r.stdout + r.stderr)
            if r.returncode:
                raise RuntimeError('nvcc failed for ' + cmd[-3])
    objs = [os.path.join(OBJ, s.replace('.cu', '.o')) for s in SOURCES]
    if force or jobs or _stale(LIB, objs):
        cmd = [NVCC, '-shared', '-o', LIB] + objs + ['-lcudart']
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError('link failed')
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='--verbose' in sys.argv))
