# -*- coding: utf-8 -*-
"""Builds tests/emul/libk2rsel_emul.so: the host emulation of the K2 register-
tile kernel (the phases header of the CUDA kernel compiled by g++).  Test
infrastructure only."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, 'k2_rsel_emul.cpp')
HDR = os.path.join(HERE, '..', '..', 'gsimcli_b200', 'csrc', 'k2_rsel_phases.cuh')
LIB = os.path.join(HERE, 'libk2rsel_emul.so')


def build(force=False):
    if (not force and os.path.exists(LIB)
            and os.path.getmtime(LIB) >= max(os.path.getmtime(SRC), os.path.getmtime(HDR))):
        return LIB
    subprocess.run(['g++', '-O2', '-std=c++17', '-ffp-contract=off', '-Wno-unknown-pragmas',
                    '-shared', '-fPIC', '-o', LIB, SRC], check=True)
    return LIB


K1_SRC = os.path.join(HERE, 'k1_dss14_emul.cpp')
K1_HDR = os.path.join(HERE, '..', '..', 'gsimcli_b200', 'csrc', 'k1_dss14.cuh')
K1_LIB = os.path.join(HERE, 'libk1dss14_emul.so')


def build_k1(force=False):
    """The K1 '%12.4f CRLF' record parser (csrc/k1_dss14.cuh) for the host."""
    if (not force and os.path.exists(K1_LIB)
            and os.path.getmtime(K1_LIB) >= max(os.path.getmtime(K1_SRC), os.path.getmtime(K1_HDR))):
        return K1_LIB
    subprocess.run(['g++', '-O2', '-std=c++17', '-ffp-contract=off', '-shared', '-fPIC',
                    '-o', K1_LIB, K1_SRC], check=True)
    return K1_LIB


SSEL_SRC = os.path.join(HERE, 'k2_ssel_emul.cpp')
SSEL_HDR = os.path.join(HERE, '..', '..', 'gsimcli_b200', 'csrc', 'k2_ssel_phases.cuh')
SSEL_LIB = os.path.join(HERE, 'libk2ssel_emul.so')


def build_ssel(force=False, sanitize=False):
    """The staged-select K2 kernel's phases (csrc/k2_ssel_phases.cuh, which
    includes k2_rsel_phases.cuh) for the host.  ``sanitize`` builds an
    AddressSanitizer + UBSan variant next to it (run in a subprocess with
    libasan preloaded)."""
    lib = SSEL_LIB.replace('.so', '_asan.so') if sanitize else SSEL_LIB
    newest = max(os.path.getmtime(SSEL_SRC), os.path.getmtime(SSEL_HDR), os.path.getmtime(HDR))
    if not force and os.path.exists(lib) and os.path.getmtime(lib) >= newest:
        return lib
    flags = ['-O1', '-g', '-fsanitize=address,undefined', '-fno-omit-frame-pointer'] if sanitize else ['-O2']
    subprocess.run(['g++'] + flags + ['-std=c++17', '-ffp-contract=off', '-Wno-unknown-pragmas',
                    '-shared', '-fPIC', '-o', lib, SSEL_SRC], check=True)
    return lib


SSEL_MAIN = os.path.join(HERE, 'k2_ssel_tsan_main.cpp')
SSEL_THR = os.path.join(HERE, 'k2_ssel_threads.cpp')
SSEL_KRN = os.path.join(HERE, '..', '..', 'gsimcli_b200', 'csrc', 'k2_ssel_kernel.cuh')


def build_ssel_threads(sanitizer=None, force=False):
    """Stand-alone driver of the THREADED emulation of the staged-select kernel
    (k2_ssel_kernel.cuh run with one OS thread per CUDA thread):
    ``sanitizer`` = None, 'thread' or 'address'.  Returns the executable."""
    exe = os.path.join(HERE, 'k2_ssel_threads_' + (sanitizer or 'plain'))
    srcs = [SSEL_MAIN, SSEL_THR, SSEL_SRC, SSEL_HDR, SSEL_KRN, HDR]
    if not force and os.path.exists(exe) and os.path.getmtime(exe) >= max(os.path.getmtime(f) for f in srcs):
        return exe
    flags = ['-fsanitize=' + sanitizer + (',undefined' if sanitizer == 'address' else '')] if sanitizer else []
    subprocess.run(['g++', '-O1', '-g', '-std=c++17', '-pthread', '-ffp-contract=off',
                    '-Wno-unknown-pragmas'] + flags + ['-o', exe, SSEL_MAIN], check=True)
    return exe


CSORT_SRC = os.path.join(HERE, 'k2_csort_threads.cpp')
CSORT_DEPS = [CSORT_SRC, os.path.join(HERE, 'cuda_threads_shim.h'), SRC, HDR,
              os.path.join(HERE, '..', '..', 'gsimcli_b200', 'csrc', 'k2_csort_kernel.cuh'),
              os.path.join(HERE, '..', '..', 'gsimcli_b200', 'csrc', 'k2_args.cuh'),
              os.path.join(HERE, '..', '..', 'gsimcli_b200', 'csrc', 'gsb_common.cuh')]


def build_csort_threads(sanitizer=None, force=False):
    """Stand-alone driver of the THREADED emulation of the counting-sort kernel
    (csrc/k2_csort_kernel.cuh on the CUDA model of cuda_threads_shim.h):
    ``sanitizer`` = None or 'thread'.  Needs the CUDA headers (types only)."""
    exe = os.path.join(HERE, 'k2_csort_threads_' + (sanitizer or 'plain'))
    if not force and os.path.exists(exe) and os.path.getmtime(exe) >= max(os.path.getmtime(f) for f in CSORT_DEPS):
        return exe
    cuda_inc = os.path.join(os.environ.get('CUDA_HOME', '/usr/local/cuda'), 'include')
    flags = ['-fsanitize=' + sanitizer, '-Wno-tsan'] if sanitizer else []
    subprocess.run(['g++', '-O1', '-g', '-std=c++17', '-pthread', '-ffp-contract=off', '-Wno-unknown-pragmas',
                    '-Wno-attributes', '-I' + cuda_inc] + flags + ['-o', exe, CSORT_SRC], check=True)
    return exe


def build_csort_lib(force=False):
    """The same emulation as a shared library (k2_csort_emulate_threads)."""
    lib = os.path.join(HERE, 'libk2csort_threads.so')
    if not force and os.path.exists(lib) and os.path.getmtime(lib) >= max(os.path.getmtime(f) for f in CSORT_DEPS):
        return lib
    cuda_inc = os.path.join(os.environ.get('CUDA_HOME', '/usr/local/cuda'), 'include')
    subprocess.run(['g++', '-O2', '-std=c++17', '-pthread', '-ffp-contract=off', '-Wno-unknown-pragmas',
                    '-Wno-attributes', '-DCSORT_LIB', '-I' + cuda_inc, '-shared', '-fPIC', '-o', lib, CSORT_SRC],
                   check=True)
    return lib


K1P_SRC = os.path.join(HERE, 'k1_parse_emul.cpp')
K1P_HDR = os.path.join(HERE, '..', '..', 'gsimcli_b200', 'csrc', 'k1_parse.cuh')
K1P_LIB = os.path.join(HERE, 'libk1parse_emul.so')


def build_k1_parse(force=False):
    """K1's general parsers (csrc/k1_parse.cuh) for the host."""
    if (not force and os.path.exists(K1P_LIB)
            and os.path.getmtime(K1P_LIB) >= max(os.path.getmtime(K1P_SRC), os.path.getmtime(K1P_HDR))):
        return K1P_LIB
    subprocess.run(['g++', '-O2', '-std=c++17', '-ffp-contract=off', '-shared', '-fPIC', '-o', K1P_LIB, K1P_SRC],
                   check=True)
    return K1P_LIB


RSELT_SRC = os.path.join(HERE, 'k2_rsel_threads.cpp')
RSELT_DEPS = [RSELT_SRC, os.path.join(HERE, 'cuda_threads_shim.h'), SRC, HDR,
              os.path.join(HERE, '..', '..', 'gsimcli_b200', 'csrc', 'k2_rsel_kernel.cuh'),
              os.path.join(HERE, '..', '..', 'gsimcli_b200', 'csrc', 'gsb_common.cuh')]


def build_rsel_threads(sanitizer=None, force=False):
    """Stand-alone driver of the THREADED emulation of the register-tile kernel's
    tile loop (csrc/k2_rsel_kernel.cuh): ``sanitizer`` = None or 'thread'."""
    exe = os.path.join(HERE, 'k2_rsel_threads_' + (sanitizer or 'plain'))
    if not force and os.path.exists(exe) and os.path.getmtime(exe) >= max(os.path.getmtime(f) for f in RSELT_DEPS):
        return exe
    cuda_inc = os.path.join(os.environ.get('CUDA_HOME', '/usr/local/cuda'), 'include')
    flags = ['-fsanitize=' + sanitizer, '-Wno-tsan'] if sanitizer else []
    subprocess.run(['g++', '-O1', '-g', '-std=c++17', '-pthread', '-ffp-contract=off', '-Wno-unknown-pragmas',
                    '-Wno-attributes', '-I' + cuda_inc] + flags + ['-o', exe, RSELT_SRC], check=True)
    return exe


KC_SRC = os.path.join(HERE, 'k2_columns_threads.cpp')
KC_DEPS = [KC_SRC, os.path.join(HERE, 'cuda_threads_shim.h'),
           os.path.join(HERE, '..', '..', 'gsimcli_b200', 'csrc', 'k2_columns_kernel.cuh'),
           os.path.join(HERE, '..', '..', 'gsimcli_b200', 'csrc', 'gsb_common.cuh')]


def build_columns(kind='lib', force=False):
    """K2c / K3 (csrc/k2_columns_kernel.cuh) on the threaded CUDA model: ``kind`` =
    'lib' (shared library for the Python fuzz) or 'thread' (stand-alone
    ThreadSanitizer binary)."""
    out = os.path.join(HERE, 'libk2columns_threads.so' if kind == 'lib' else 'k2_columns_threads_thread')
    if not force and os.path.exists(out) and os.path.getmtime(out) >= max(os.path.getmtime(f) for f in KC_DEPS):
        return out
    cuda_inc = os.path.join(os.environ.get('CUDA_HOME', '/usr/local/cuda'), 'include')
    base = ['g++', '-std=c++17', '-pthread', '-ffp-contract=off', '-Wno-unknown-pragmas', '-Wno-attributes',
            '-I' + cuda_inc]
    if kind == 'lib':
        cmd = base + ['-O2', '-shared', '-fPIC', '-o', out, KC_SRC]
    else:
        cmd = base + ['-O1', '-g', '-DKC_MAIN', '-fsanitize=thread', '-Wno-tsan', '-o', out, KC_SRC]
    subprocess.run(cmd, check=True)
    return out
