# -*- coding: utf-8 -*-
"""Live checks of the oracle against the reference's own source, executed
through oracle/refshim.py.  They run only where /root/reference exists (the
build container); on the GPU box they are skipped and the committed golden
fixtures stand in."""
import os
import warnings

import numpy as np
import pandas as pd
import pytest

from gsimcli_b200 import synth
from oracle import np_oracle as O, refshim

pytestmark = [pytest.mark.refshim,
              pytest.mark.skipif(not refshim.available(),
                                 reason='reference checkout not present')]
ND = -999.9
ALL = dict(lmean=True, lmed=True, lskew=True, lvar=True, lstd=True,
           lcoefvar=True, lperc=True)
ALLF = (O.FLAG_MEAN | O.FLAG_MED | O.FLAG_SKEW | O.FLAG_VAR | O.FLAG_STD
        | O.FLAG_COEFVAR | O.FLAG_PERC)


def _load(tmp_path, st, dims, first, cells, header=0):
    _, grid, _ = refshim.load_reference()
    hdr = ['h', '1', 'v'] if header else None
    firstfile = synth.write_stack(str(tmp_path), 'r_sim', st, header=hdr)
    g = grid.GridFiles()
    g.load(firstfile, st.shape[0], dims, first, cells, ND, header)
    return g


def test_fresh_random_stack_stats_and_area(tmp_path):
    dims, first, cells = [8, 7, 4], [100.0, 200.0, 1980], [10.0, 20.0, 1]
    st = synth.stack(int.from_bytes(os.urandom(2), 'little'), 11,
                     dims).astype(np.float64)
    st[st == np.float64(np.float32(ND))] = ND
    g = _load(tmp_path, st, dims, first, cells)
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        ref = g.stats(p=0.9, **ALL)
        mine = O.grid_stats(st, ND, ALLF, 0.9)
        for key in ref:
            np.testing.assert_allclose(mine[key], ref[key].val, rtol=5e-13,
                                       atol=1e-10, equal_nan=True)
        assert np.array_equal(mine['medianmap'], ref['medianmap'].val,
                              equal_nan=True)
        loc = [100.0 + 3.2 * 10, 200.0 + 2.7 * 20]
        for tol in (0, 1, 2):
            g.reset_read()
            ps = g.stats_area(loc, tol=tol, p=0.95, **ALL)
            name, tab = O.stats_area(st, dims, first, cells, ND, loc, tol,
                                     ALLF, 0.95)
            assert name == ps.name
            np.testing.assert_allclose(tab.values, ps.values.values,
                                       rtol=5e-13, atol=1e-10, equal_nan=True)


def test_quirk_circle_at_grid_edge_is_confined_to_edges(tmp_path):
    """the reference keeps node 0 / nodes beyond dx, which alias
    other lines.  Literal mode (clip=False) reproduces the requested lines;
    the intended behaviour (clip=True) differs only for edge-touching discs."""
    dims = [4, 3, 2]
    node, lit = O.area_lines([0.0, 0.0], 1, dims, [0, 0, 0], [1, 1, 1],
                             clip=False)
    assert list(node) == [1, 1]
    assert sorted(lit[:, 0]) == [-3, 0, 1, 2, 5]
    _, clipped = O.area_lines([0.0, 0.0], 1, dims, [0, 0, 0], [1, 1, 1])
    assert sorted(clipped[:, 0]) == [1, 2, 5]
    # interior station: identical either way
    _, a = O.area_lines([1.0, 1.0], 1, dims, [0, 0, 0], [1, 1, 1], clip=False)
    _, b = O.area_lines([1.0, 1.0], 1, dims, [0, 0, 0], [1, 1, 1], clip=True)
    assert np.array_equal(a, b)


def test_quirk_header_rows_left_zero_by_reference(tmp_path):
    """with header=3 the reference leaves the last 3 nodes 0."""
    dims, first, cells = [3, 2, 2], [0, 0, 0], [1, 1, 1]
    st = synth.stack(5, 6, dims).astype(np.float64)
    g = _load(tmp_path, st, dims, first, cells, header=3)
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        ref = g.stats(lmean=True)['meanmap'].val
    mine = O.grid_stats(st, np.float64(np.float32(ND)), O.FLAG_MEAN)['meanmap']
    assert np.all(ref[-3:] == 0)
    np.testing.assert_allclose(mine[:-3], ref[:-3], rtol=1e-13)


def test_detect_roundtrip_take_candidate_update_station(tmp_path):
    """take_candidate -> detect -> update_station through the reference and
    through the oracle agree on counts, flags and corrected values."""
    _, grid, homog = refshim.load_reference()
    dims, first, cells = [8, 7, 6], [0.0, 0.0, 2000], [10.0, 10.0, 1]
    R = 10
    st = synth.stack(77, R, dims).astype(np.float64)
    st[st == np.float64(np.float32(ND))] = ND
    g = _load(tmp_path, st, dims, first, cells)
    stations = synth.stations(77, 3, R, dims, first, cells)
    rows = np.vstack([s['obs'] for s in stations])
    cols = ['x', 'y', 'time', 'station', 'clim']
    net = grid.PointSet('net', ND, 5, list(cols),
                        pd.DataFrame(rows, columns=cols))
    cand, refs = homog.take_candidate(net, 2)
    obs_df = cand.values.copy()
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        hom, dn, md = homog.detect(g, cand, method='median', prob=0.9)
    ohom, odn, omd = O.detect(st, dims, first, cells, ND, obs_df, ND,
                              method='median', prob=0.9)
    assert (dn, md) == (odn, omd)
    assert np.array_equal(hom.values['Flag'].values, ohom['Flag'].values)
    np.testing.assert_allclose(hom.values['clim'].values, ohom['clim'].values,
                               rtol=5e-13)


def _network(seed=4):
    rng = np.random.default_rng(seed)
    rows = []
    for st in (3, 1, 7, 5):
        for y in range(2000, 2006):
            v = rng.normal(10 + st, 1 + 0.3 * st)
            if st == 7 and y == 2002:
                v = ND
            rows.append([float(st * 10), float(st * 20), float(y), float(st), v])
    return pd.DataFrame(rows, columns=['x', 'y', 'time', 'station', 'clim'])


def _py2_list_stations(homog):
    """py2's map() returns a list; station_order / save_output of the reference
    rely on that (``.sort``, ``len``).  Seventh py2->py3 adaptation, applied to
    the loaded module only."""
    orig = homog.list_stations
    if getattr(orig, '_listed', False):
        return
    def listed(*a, **k):
        return list(orig(*a, **k))
    listed._listed = True
    homog.list_stations = listed


def test_station_order_matches_reference():
    from gsimcli_b200.tools import grid as mygr, homog as myhmg
    _, grid, homog = refshim.load_reference()
    _py2_list_stations(homog)
    df = _network()
    ref_ps = grid.PointSet('net', ND, 5, list(df.columns), df.copy())
    my_ps = mygr.PointSet('net', ND, 5, list(df.columns), df.copy())
    for method in ('sorted', 'variance', 'network deviation'):
        for asc in (True, False):
            for md_last in (True, False):
                a = homog.station_order(method, ref_ps, nd=ND, ascending=asc, md_last=md_last)
                b = myhmg.station_order(method, my_ps, nd=ND, ascending=asc, md_last=md_last)
                assert list(a) == list(b), (method, asc, md_last)
    assert myhmg.station_order('user', my_ps, userset=[5, 1]) == \
        homog.station_order('user', ref_ps, userset=[5, 1])
    assert sorted(myhmg.station_order('random', my_ps)) == [1, 3, 5, 7]
    with pytest.raises(TypeError):
        myhmg.station_order('nonsense', my_ps)


def test_save_output_gsimcli_format_matches_reference(tmp_path):
    from gsimcli_b200.tools import grid as mygr, homog as myhmg
    _, grid, homog = refshim.load_reference()
    _py2_list_stations(homog)
    df = _network()
    df['Flag'] = ND
    df.loc[df.index[3], 'Flag'] = df.loc[df.index[3], 'clim']
    df['median'] = np.nan                      # optional statistic of one station only
    df.loc[df['station'] == 1.0, 'median'] = 11.25
    cols = list(df.columns)
    ref_ps = grid.PointSet('net', ND, len(cols), list(cols), df.copy())
    my_ps = mygr.PointSet('net', ND, len(cols), list(cols), df.copy())
    a, b = str(tmp_path / 'ref.csv'), str(tmp_path / 'mine.csv')
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        homog.save_output(ref_ps, a, fformat='gsimcli', save_stations=True)
        myhmg.save_output(my_ps, b, fformat='gsimcli', save_stations=True)
    ra, rb = pd.read_csv(a, index_col=0), pd.read_csv(b, index_col=0)
    assert list(ra.columns) == list(rb.columns)
    assert list(ra.index) == list(rb.index)
    np.testing.assert_allclose(ra.values.astype(float), rb.values.astype(float),
                               rtol=0, atol=0, equal_nan=True)
    sa = pd.read_csv(str(tmp_path / 'ref_stations.csv'), index_col=0)
    sb = pd.read_csv(str(tmp_path / 'mine_stations.csv'), index_col=0)
    assert sa.equals(sb)


@pytest.mark.parametrize('seed', [11, 12, 13, 14, 15, 16])
def test_detect_every_method_on_fresh_stacks(tmp_path, seed):
    """All four correction methods, discs of radius 0..2, several detection
    probabilities, R even and odd, on stacks and observed series drawn from the
    seed: the reference's detect() (tools/homog.py:27-276, through GridFiles.
    stats_area) and the oracle agree on the counts, the flags and every value
    of the homogenised frame."""
    _, grid, homog = refshim.load_reference()
    rng = np.random.default_rng(seed)
    dims, first, cells = [9, 8, 5 + seed % 3], [500.0, 100.0, 1970], [25.0, 50.0, 1]
    R = 8 + seed % 5
    st = synth.stack(seed, R, dims).astype(np.float64)
    st[st == np.float64(np.float32(ND))] = ND
    g = _load(tmp_path, st, dims, first, cells)
    cols = ['x', 'y', 'time', 'station', 'clim']
    cases = [dict(method='mean'), dict(method='median'),
             dict(method='skewness', skewness=0.3), dict(method='skewness', skewness=5.0),
             dict(method='percentile', percentile=0.9), dict(method='percentile', percentile=0.25)]
    checked = 0
    for s in synth.stations(seed, 3, R, dims, first, cells, tol=2):
        obs = pd.DataFrame(np.array(s['obs'], dtype=np.float64), columns=cols)
        for kw in cases:
            for rad in (0, 1, 2):
                prob = float(rng.choice([0.8, 0.9, 0.95, 0.99]))
                g.reset_read()
                cand = grid.PointSet('cand', ND, 5, list(cols), obs.copy())
                with warnings.catch_warnings():
                    warnings.simplefilter('ignore')
                    hom, dn, md = homog.detect(g, cand, rad=rad, prob=prob, **kw)
                ohom, odn, omd = O.detect(st, dims, first, cells, ND, obs.copy(), ND,
                                          rad=rad, prob=prob, **kw)
                assert (dn, md) == (odn, omd), (kw, rad, prob)
                assert list(hom.values.columns) == list(ohom.columns)
                np.testing.assert_allclose(hom.values.values.astype(float),
                                           ohom.values.astype(float), rtol=5e-13, atol=1e-10,
                                           equal_nan=True)
                checked += 1
    assert checked == 54


@pytest.mark.parametrize('seed', [21, 22, 23, 24])
def test_stats_every_map_on_fresh_stacks_with_edge_columns(tmp_path, seed):
    """GridFiles.stats of the reference (tools/grid.py:601-735: bottleneck nan*
    functions, pandas skew / quantile) against the oracle on a fresh stack with
    the columns whose numerics are corner cases: constant, empty, n = 1, n = 2,
    two atoms, heavy ties, a zero-inflated one; even and odd R, three p."""
    rng = np.random.default_rng(seed)
    dims, first, cells = [7, 6, 3], [0.0, 0.0, 1], [1.0, 1.0, 1]
    R = 9 + seed % 4
    st = synth.stack(seed, R, dims).astype(np.float64)
    st[st == np.float64(np.float32(ND))] = ND
    st[:, 3] = 812.5
    st[:, 4] = ND
    st[:, 5] = ND
    st[R // 2, 5] = 640.25
    st[:, 6] = ND
    st[0, 6], st[R - 1, 6] = 7.5, 8.5
    st[:, 8] = np.where(np.arange(R) % 2 == 0, 500.0, 600.0)
    st[:, 9] = rng.integers(0, 3, size=R) * 0.5
    st[:, 10] = np.where(rng.uniform(size=R) < 0.7, 0.0, np.round(rng.gamma(2.0, 20.0, size=R), 4))
    g = _load(tmp_path, st, dims, first, cells)
    for p in (0.95, 0.8, 0.5):
        g.reset_read()
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            ref = g.stats(p=p, **ALL)
            mine = O.grid_stats(st, ND, ALLF, p)
        for key in ref:
            np.testing.assert_allclose(mine[key], ref[key].val, rtol=5e-13, atol=1e-10,
                                       equal_nan=True, err_msg=key)
        assert np.array_equal(mine['medianmap'], ref['medianmap'].val, equal_nan=True)
        assert np.array_equal(mine['percmap'], ref['percmap'].val, equal_nan=True)


def test_stats_fuzz_against_the_reference(tmp_path):
    """Random float64 stacks of the fuzz families (tests/emul/fuzz_csort.py: lattice
    normals, gammas, clamped normals, few-valued, zero-inflated, values centred on
    zero over 13 decades, CONSTANT columns, heavy tails; six missing-value
    patterns), written as '%12.4f' text and reduced by the reference's own
    GridFiles.stats: all seven maps of the oracle agree.  (The first run of this
    fuzz showed the oracle's vectorised skewness leaving the reference on constant
    columns whose value does not sum exactly in float64 -- pandas zeroes the
    second moment below a tolerance and the oracle's mean, summed in another
    order, missed it; the oracle now sums as pandas does.  10^4 stacks clean.)"""
    import shutil
    from emul import fuzz_csort
    _, grid, _ = refshim.load_reference()
    rng = np.random.default_rng(5)
    for k in range(150):
        R = int(rng.choice([3, 5, 8, 9, 16, 33, 40]))
        st = np.stack([fuzz_csort.column(rng, R, int(rng.integers(0, 8))) for _ in range(24)], axis=1)
        st = np.round(st, 4)
        st[np.abs(st) > 9.9e6] = 1234.5
        d = tmp_path / ('s%d' % k)
        d.mkdir()
        g = _load(d, st, [4, 3, 2], [0, 0, 0], [1, 1, 1])
        p = float(rng.choice([0.95, 0.9, 0.5]))
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            ref = g.stats(p=p, **ALL)
            mine = O.grid_stats(st, ND, ALLF, p)
        g.dump()
        shutil.rmtree(str(d), ignore_errors=True)
        for key in ref:
            np.testing.assert_allclose(mine[key], ref[key].val, rtol=5e-12, atol=1e-9, equal_nan=True,
                                       err_msg='{0} (stack {1}, R = {2})'.format(key, k, R))
