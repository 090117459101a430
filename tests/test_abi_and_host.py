# -*- coding: utf-8 -*-
"""CPU-side tests: the C-ABI library loads and exports what the header
declares, the host logic mirrors the reference, and the product never routes
through the oracle or a CPU fallback."""
import ast
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pandas as pd
import pytest

from conftest import ROOT, load_golden
from gsimcli_b200 import _capi, synth, dist
from gsimcli_b200.tools import grid as gr, homog as hmg
from oracle import np_oracle as O


def header_symbols():
    text = open(os.path.join(ROOT, 'include', 'gsimcli_b200.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(gsb_[a-z0-9_]+)\s*\(', text)))


def test_library_exports_every_header_symbol():
    lib = _capi.lib()
    syms = header_symbols()
    assert len(syms) >= 19
    for name in syms:
        assert hasattr(lib, name), name
    # and the ctypes table binds exactly that set
    assert sorted(_capi.SIGNATURES) == syms
    assert lib.gsb_version() >= 100


def test_error_buffer_and_plane_count():
    lib = _capi.lib()
    assert lib.gsb_stats_grid_planes(_capi.MEAN | _capi.PERC, 3) == 6
    assert lib.gsb_stats_grid_planes(0xFF, 19) == 9 + 19
    assert lib.gsb_stats_grid_planes(0xFF, 60) == -1
    buf = ctypes.create_string_buffer(64)
    assert lib.gsb_last_error(buf, 64) == 0


def test_no_cpu_fallback_without_device():
    if _capi.device_count() > 0:
        pytest.skip('a CUDA device is present')
    with pytest.raises(RuntimeError):
        _capi.Stack(4, 16)
    with pytest.raises(RuntimeError):
        gr.GridFiles.from_array(np.zeros((2, 8), np.float32), [2, 2, 2],
                                [0, 0, 0], [1, 1, 1], -999.9)
    with pytest.raises(RuntimeError):
        gr.GridFiles().stats(lmean=True)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, 'gsimcli_b200')
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if not f.endswith('.py'):
                continue
            tree = ast.parse(open(os.path.join(dirpath, f)).read())
            for node in ast.walk(tree):
                names = []
                if isinstance(node, ast.Import):
                    names = [a.name for a in node.names]
                elif isinstance(node, ast.ImportFrom):
                    names = [node.module or '']
                for n in names:
                    assert not n.split('.')[0] == 'oracle', (f, n)
                    assert 'refshim' not in n, (f, n)


def test_indexing_helpers_match_oracle():
    rng = np.random.default_rng(1)
    for _ in range(200):
        cells = [float(rng.integers(1, 2000)), float(rng.integers(1, 2000)), 1]
        first = [float(rng.integers(-5000, 5000)), float(rng.integers(0, 9e6)),
                 1990]
        loc = [first[0] + rng.uniform(0, 80) * cells[0],
               first[1] + rng.uniform(0, 120) * cells[1]]
        assert list(gr.coord_to_grid(loc, cells, first)) == \
            list(O.coord_to_grid(loc, cells, first))
    # exact .5 cases: half-to-even, like np.around (grid.py:953)
    assert list(gr.coord_to_grid([0.5, 1.5, 2.5], [1, 1, 1], [0, 0, 0])) == \
        [2, 2, 4]
    for r in range(6):
        for xc, yc in ((10, 10), (2, 1), (1, 1), (0, 3)):
            assert np.array_equal(gr.circle(xc, yc, r), O.circle(xc, yc, r))
            assert np.array_equal(gr.circle(xc, yc, r, [12, 11, 3]),
                                  O.circle(xc, yc, r, [12, 11, 3]))
    assert gr.line_zmirror([2, 3], [4, 5, 3]) == O.line_zmirror([2, 3],
                                                                [4, 5, 3])
    assert gr.grid_to_line([46, 98, 1], [81, 122, 10]) == 46 + 81 * 97


def test_pointset_and_gridarr_io(tmp_path):
    df = pd.DataFrame({'x': [1.0, 2.0], 'y': [3.0, 4.0], 'clim': [5.5, -999.9]})
    ps = gr.PointSet('demo', -999.9, 3, list(df.columns), df)
    path = str(tmp_path / 'p.prn')
    ps.save(path)
    back = gr.PointSet(psetpath=path)
    assert back.name == 'demo' and back.nvars == 3
    assert back.varnames == ['x', 'y', 'clim']
    assert np.allclose(back.values.values, df.values)
    # header=False: names var1..varN, name from the file name (grid.py:159-163)
    ps.save(path, header=False)
    nh = gr.PointSet()
    nh.load(path, header=False)
    assert nh.varnames == ['var1', 'var2', 'var3'] and nh.name == 'p'
    # add_var suffix on clash (grid.py:126-127)
    ps.add_var([1, 2], 'clim')
    assert ps.varnames[-1] == 'clim_new' and ps.nvars == 4
    # PointSet(values=ndarray) works (reference HEAD raises, grid.py:99)
    assert gr.PointSet(values=np.ones((2, 2))).values.shape == (2, 2)
    ga = gr.GridArr(name='m', dx=2, dy=1, dz=2, val=np.array([1.5, 2, 3, 4.25]))
    gpath = str(tmp_path / 'm.out')
    ga.save(gpath, 'mean')
    lines = open(gpath).read().splitlines()
    assert lines[:3] == ['m', '1', 'mean'] and lines[3] == '1.500000  '
    gb = gr.GridArr()
    gb.load(gpath, [2, 1, 2], [0, 0, 0], [1, 1, 1])
    assert np.array_equal(gb.val, ga.val)
    well = ga.drill([1.0, 0.0])       # node (2, 1): values 2 and 4.25
    assert list(well.values['var']) == [2.0, 4.25]


def test_fill_station_matches_oracle():
    g = load_golden('dyadic_even')
    rows = None
    for rec in g['detect']:
        if len(rec['obs']) < g['dims'][2]:
            rows = rec['obs']
            break
    assert rows is not None
    cols = ['x', 'y', 'time', 'station', 'clim']
    df = pd.DataFrame(np.array(rows), columns=cols)
    fill = np.arange(g['dims'][2]) * 10.0 + 1
    zi = g['first_coord'][2]
    ps = gr.PointSet('c', -999.9, 5, list(cols), df.copy())
    got, n = hmg.fill_station(ps, fill, zi, zi + g['dims'][2], 1)
    exp, m = O.fill_station(df, fill, zi, zi + g['dims'][2], 1)
    assert n == m == g['dims'][2] - len(rows)
    assert np.array_equal(got.values.values, exp.values)


@pytest.mark.parametrize('years', [
    [1990, 1991, 1993, 1996],            # regular: increasing, on the axis
    [1991, 1990, 1993],                  # out of order: the merge walk stalls
    [1990, 1992, 1992, 1995],            # duplicate year
    [1989, 1991, 1994],                  # a year before the axis
    [1990, 1993, 2005],                  # a year beyond the axis
    [1992.5, 1994],                      # off the axis
])
def test_fill_station_irregular_series_follow_the_reference_walk(years):
    """The vectorised fill_station must give what the reference's merge walk
    (homog.py:279-337, restated in the oracle) gives, also for series the walk
    handles oddly (it never skips an observed row it cannot match)."""
    cols = ['x', 'y', 'time', 'station', 'clim']
    rows = np.array([[3.0, 4.0, y, 7.0, 100.0 + i] for i, y in enumerate(years)])
    df = pd.DataFrame(rows, columns=cols)
    fill = np.arange(8) * 10.0 + 1
    ps = gr.PointSet('c', -999.9, 5, list(cols), df.copy())
    got, n = hmg.fill_station(ps, fill, 1990, 1998, 1)
    exp, m = O.fill_station(df, fill, 1990, 1998, 1)
    assert n == m
    assert np.array_equal(got.values.values, exp.values)


def test_take_candidate_update_station_roundtrip():
    cols = ['x', 'y', 'time', 'station', 'clim']
    rows = [[0, 0, t, s, 10.0 * s + t] for s in (1, 2, 3) for t in range(4)]
    net = gr.PointSet('net', -999.9, 5, list(cols),
                      pd.DataFrame(np.array(rows, float), columns=cols))
    assert hmg.list_stations(net) == [1, 2, 3]
    cand, refs = hmg.take_candidate(net, 2)
    assert cand.name == 'Candidate_2' and refs.name == 'References_2'
    assert len(cand.values) == 4 and len(refs.values) == 8
    cand.values = cand.values.copy()
    cand.values['clim'] = -1.0
    cand.add_var(np.full(4, 7.0), 'Flag')
    upd = hmg.update_station(net, cand)
    assert (upd.values.loc[upd.values['station'] == 2, 'clim'] == -1.0).all()
    assert (upd.values.loc[upd.values['station'] == 1, 'clim'] != -1.0).all()
    assert 'Flag' in upd.values.columns
    with pytest.raises(ValueError):
        hmg._select_stats('nonsense', None, None)
    with pytest.raises(ValueError):
        hmg._select_stats('skewness', None, None)   # threshold missing


def test_slab_ranges_partition_the_grid():
    dims = [20, 10, 7]
    for world in (1, 2, 3, 4, 8):
        edges = [dist.slab_range(dims, r, world) for r in range(world)]
        assert edges[0][0] == 0 and edges[-1][1] == 1400
        for a, b in zip(edges[:-1], edges[1:]):
            assert a[1] == b[0]
        assert all((b - a) % 200 == 0 for a, b in edges)


def test_synth_is_dyadic_and_text_roundtrips():
    st = synth.stack(3, 5, [11, 7, 3])
    valid = st[st != np.float32(-999.9)]
    assert np.array_equal(valid * 16, np.round(valid * 16))
    text = synth.to_text(st[0])
    assert len(text) == 14 * st.shape[1]
    back = np.array([float(x) for x in text.split()], dtype=np.float32)
    assert np.array_equal(back, st[0])


WORKER = r'''
import os, sys
sys.path.insert(0, {root!r})
import numpy as np
import torch.distributed as td
from gsimcli_b200 import dist
rank = int(sys.argv[1]); world = 2
os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT={port!r})
td.init_process_group('gloo', rank=rank, world_size=world)
dims = [4, 3, 5]
full = np.arange(3 * 5 * 2, dtype=float).reshape(3, 5, 2)
z0, z1 = dist.layer_range(dims, rank, world)
local = np.full_like(full, np.nan)
local[:, z0:z1] = full[:, z0:z1]
got = dist.allgather_tables(local, dims)
assert np.array_equal(got, full), (rank, got)
# the real layer ranges are exchanged: an uneven split the default would not guess
zr = (0, 1) if rank == 0 else (1, 5)
local = np.full_like(full, np.nan)
local[:, zr[0]:zr[1]] = full[:, zr[0]:zr[1]]
got = dist.allgather_tables(local, dims, zrange=zr, sharded=True)
assert np.array_equal(got, full), (rank, got)
# slabs that do not tile the time axis are an error, not NaN rows
try:
    dist.allgather_tables(local, dims, zrange=(0, 2) if rank == 0 else (1, 5))
    raise SystemExit('no error for overlapping slabs')
except ValueError:
    pass
td.destroy_process_group()
print('ok', rank)
'''


def test_allgather_tables_needs_a_process_group_when_sharded():
    from gsimcli_b200 import dist
    local = np.zeros((1, 4, 2))
    assert dist.allgather_tables(local, [2, 2, 4]) is local          # single process
    with pytest.raises(RuntimeError):
        dist.allgather_tables(local, [2, 2, 4], sharded=True)


def test_allgather_tables_world2_gloo(tmp_path):
    port = str(29500 + os.getpid() % 2000)
    script = tmp_path / 'w.py'
    script.write_text(WORKER.format(root=ROOT, port=port))
    procs = [subprocess.Popen([sys.executable, str(script), str(r)],
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
             for r in range(2)]
    outs = [p.communicate(timeout=120)[0].decode() for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o
        assert 'ok' in o


def test_library_reads_no_environment_switches():
    """kernel selection must not hang off getenv in product code;
    the test-only knobs are stack options (gsb_stack_set_option)."""
    csrc = os.path.join(ROOT, 'gsimcli_b200', 'csrc')
    for name in os.listdir(csrc):
        if name.endswith(('.cu', '.cuh')):
            with open(os.path.join(csrc, name)) as fh:
                assert 'getenv' not in fh.read(), name
