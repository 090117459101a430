# -*- coding: utf-8 -*-
"""Pin oracle/np_oracle.py against outputs of the reference's own code
(tests/golden/*.json, produced by oracle/make_golden.py through the shim)."""
import numpy as np
import pandas as pd
import pytest

from oracle import np_oracle as O

ALLF = (O.FLAG_MEAN | O.FLAG_MED | O.FLAG_SKEW | O.FLAG_VAR | O.FLAG_STD
        | O.FLAG_COEFVAR | O.FLAG_PERC)
# bottleneck -> numpy summation order: a few f64 ulp (SURVEY 8c limits (2))
RTOL = 5e-13


def close(a, b, rtol=RTOL, atol=1e-10):
    a = np.asarray(a, float)
    b = np.asarray(b, float)
    assert a.shape == b.shape
    assert np.array_equal(np.isnan(a), np.isnan(b))
    np.testing.assert_allclose(a, b, rtol=rtol, atol=atol, equal_nan=True)


def test_grid_stats_match_reference(golden):
    st = np.array(golden['stack'])
    hdr = golden['header']
    for rec in golden['stats']:
        res = O.grid_stats(st, golden['nodata'], ALLF, rec['p'])
        n = st.shape[1] - hdr      # reference leaves last `header` nodes 0
        for key, ref in rec['maps'].items():
            ref = np.array(ref)
            close(res[key][:n], ref[:n])
            if hdr:
                assert np.all(ref[n:] == 0)
        # order statistics are selected elements: bit-exact
        assert np.array_equal(res['medianmap'][:n],
                              np.array(rec['maps']['medianmap'])[:n],
                              equal_nan=True)


def test_scalar_and_vector_forms_agree(golden):
    st = np.array(golden['stack'])
    res = O.grid_stats(st, golden['nodata'], ALLF | O.FLAG_MODE, 0.95,
                       percentiles=[0.05, 0.5, 0.95])
    for node in range(st.shape[1]):
        c = O.column_stats(st[:, node], golden['nodata'], ALLF | O.FLAG_MODE,
                           0.95, percentiles=[0.05, 0.5, 0.95])
        for key, mkey in (('mean', 'meanmap'), ('median', 'medianmap'),
                          ('skewness', 'skewmap'), ('variance', 'varmap'),
                          ('std', 'stdmap'), ('coefvar', 'coefvarmap'),
                          ('mode', 'modemap')):
            close(c[key], res[mkey][node], rtol=1e-12)
        close([c['lperc'], c['rperc']], res['percmap'][node], rtol=0, atol=0)
        close(c['percentiles'], res['percentiles'][node], rtol=0, atol=0)
        p = O.column_stats_pandas(st[:, node], golden['nodata'], ALLF, 0.95)
        for key in p:
            close(p[key], c[key], rtol=1e-12)
        assert (np.isnan(p['median']) and np.isnan(c['median'])) or \
            p['median'] == c['median']
        close([p['lperc'], p['rperc']], [c['lperc'], c['rperc']], rtol=0,
              atol=0)


def test_stats_area_match_reference(golden):
    st = np.array(golden['stack'])
    for rec in golden['stats_area']:
        name, tab = O.stats_area(st, golden['dims'], golden['first_coord'],
                                 golden['cells_size'], golden['nodata'],
                                 rec['loc'], rec['tol'], ALLF, rec['p'])
        assert name == rec['name']
        assert list(tab.columns) == rec['table']['columns']
        ref = np.array(rec['table']['values'])
        close(tab.values, ref)
        for col in ('x', 'y', 'z', 'median'):
            j = list(tab.columns).index(col)
            assert np.array_equal(tab.values[:, j], ref[:, j], equal_nan=True)


def test_detect_match_reference(golden):
    st = np.array(golden['stack'])
    for rec in golden['detect']:
        obs = pd.DataFrame(np.array(rec['obs'], float),
                           columns=['x', 'y', 'time', 'station', 'clim'])
        hom, dn, md = O.detect(st, golden['dims'], golden['first_coord'],
                               golden['cells_size'], golden['nodata'], obs,
                               golden['nodata'], **rec['kwargs'])
        assert dn == rec['detected_number']
        assert md == rec['missing_data']
        ref = pd.DataFrame(np.array(rec['homogenised']['values']),
                           columns=rec['homogenised']['columns'])
        assert sorted(hom.columns) == sorted(ref.columns)
        for col in ref.columns:
            close(hom[col].values, ref[col].values)
        # flags are bit-exact: same rows flagged, same stored observed value
        assert np.array_equal(hom['Flag'].values, ref['Flag'].values)


def test_known_answers_appendix_c():
    """SURVEY Appendix C, stated without the golden file."""
    nd = -999.9
    c1 = O.column_stats([3, 1, nd, 2, 10, 4], nd, ALLF, 0.95)
    assert c1['mean'] == 4.0 and c1['median'] == 3.0 and c1['variance'] == 12.5
    assert c1['std'] == 3.5355339059327378
    assert c1['coefvar'] == 88.38834764831844
    assert abs(c1['skewness'] - 1.697056274847714) < 1e-15
    assert c1['lperc'] == 1.1 and c1['rperc'] == 9.399999999999999
    c2 = O.column_stats([5] * 6, nd, ALLF, 0.95)
    assert c2['skewness'] == 0.0 and c2['variance'] == 0.0
    c3 = O.column_stats([7.5, nd, nd, nd, nd, 8.5], nd, ALLF, 0.95)
    assert np.isnan(c3['skewness']) and c3['lperc'] == 7.525
    c4 = O.column_stats([nd] * 6, nd, ALLF | O.FLAG_MODE, 0.95)
    assert all(np.isnan(c4[k]) for k in ('mean', 'median', 'lperc', 'mode'))
    assert O.column_stats([1, 2, 2, 3, 3], nd, O.FLAG_MODE)['mode'] == 2
    assert O.column_stats([4, 1, 9], nd, O.FLAG_MODE)['mode'] == 1


def test_indexing_helpers():
    # station 9 of the reference fixture -> node (46, 98, 1)
    g = O.coord_to_grid([1815109.147, 7190958.185], [1000, 1000, 1],
                        [1770000, 7094000, 1990])
    assert list(g) == [46, 98, 1990 - 1990 + 1]
    # half-to-even rounding (grid.py:953)
    assert list(O.coord_to_grid([0.5, 1.5], [1, 1, 1], [0, 0, 0])[:2]) == [2, 2]
    assert [len(O.circle(10, 10, r)) for r in range(5)] == [1, 5, 13, 29, 49]
    assert O.line_zmirror([2, 3], [4, 5, 3]) == [10, 30, 50]
    # intended edge behaviour: clipped to the grid
    assert len(O.circle(1, 1, 1, dims=[4, 3, 2])) == 3
