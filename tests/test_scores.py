# -*- coding: utf-8 -*-
"""tools/scores.py against hand-computed values and against the arithmetic of
the reference's crmse (scores.py:26-70) executed from its source text."""
import os
import re

import numpy as np
import pandas as pd
import pytest

from gsimcli_b200.tools import scores

REF = '/root/reference/GSIMCLI/tools/scores.py'


def test_crmse_known_answers():
    h = pd.Series([1.0, 2.0, 3.0, 4.0])
    o = pd.Series([1.0, 2.0, 3.0, 8.0])
    # centred: h - 2.5, o - 3.5 -> differences 1, 1, 1, -3 -> sqrt(12 / 4)
    assert scores.crmse(h, o) == pytest.approx(np.sqrt(3.0))
    assert scores.crmse(h, o, centered=False) == pytest.approx(2.0)
    assert scores.crmse(h, h + 10.0) == pytest.approx(0.0)          # a constant shift is no error
    hn = pd.Series([1.0, np.nan, 3.0, 4.0])
    assert scores.crmse(hn, o, centered=False) == pytest.approx(np.sqrt(16.0 / 3))
    assert scores.crmse(h, o, crop=1) == pytest.approx(0.0)


def test_network_and_submission_aggregation():
    years = np.arange(1990, 2000)
    rng = np.random.default_rng(1)
    orig = pd.DataFrame(rng.normal(10, 1, size=(10, 3)), index=years, columns=[1, 2, 3])
    homog = orig + rng.normal(0, 0.1, size=orig.shape)
    inho = orig.copy()
    inho.loc[1995:, 2] += 1.5                                        # a break
    nets_h = {'n1': (homog, orig), 'n2': (homog * 1.0, orig)}
    nets_i = {'n1': (inho, orig), 'n2': (inho, orig)}
    net, st = scores.crmse_submission(nets_h)
    assert net == pytest.approx(scores.crmse(homog.mean(axis=1), orig.mean(axis=1)))
    assert st == pytest.approx(np.mean([scores.crmse(homog[c], orig[c]) for c in orig.columns]))
    h, i, imp = scores.improvement(nets_h, nets_i)
    assert imp[0] == pytest.approx(h[0] / i[0]) and imp[1] < 1.0     # homogenised is better
    assert scores.crmse_submission(nets_h, over_station=False) == [pytest.approx(net)]


def test_scores_from_save_output_csv(tmp_path):
    from gsimcli_b200.tools import grid as gr, homog as hmg
    years = np.arange(2000, 2006)
    rows = []
    truth = {}
    for st in (10, 20):
        truth[st] = 5.0 + 0.1 * st + np.sin(years)
        for y, v in zip(years, truth[st]):
            rows.append([1.0 * st, 2.0 * st, float(y), float(st), v + (0.5 if st == 20 else 0.0),
                         -999.9])
    df = pd.DataFrame(rows, columns=['x', 'y', 'time', 'station', 'clim', 'Flag'])
    ps = gr.PointSet('net', -999.9, 6, list(df.columns), df)
    out = str(tmp_path / 'net_homogenised_data.csv')
    hmg.save_output(ps, out, fformat='gsimcli')
    table = scores.load_gsimcli_results(out)
    assert list(table.columns) == [10, 20] and list(table.index) == list(years.astype(float))
    orig = pd.DataFrame(truth, index=years.astype(float))
    res = scores.gsimcli_improvement({'net': out}, {'net': orig})
    assert res[0][1] == pytest.approx(0.0, abs=1e-12)               # shifts vanish when centred


@pytest.mark.skipif(not os.path.isfile(REF), reason='reference checkout not present')
def test_crmse_equals_reference_source():
    """The reference's crmse body, taken from its source and run on NumPy-backed
    bottleneck stand-ins (its module imports GUI code and cannot be imported)."""
    src = open(REF).read()
    body = src[src.index('def crmse('):src.index('def crmse_station(')]
    body = re.sub(r'"""(.|\n)*?"""', '', body, count=1)              # drop the docstring (raw escapes)
    ns = {'np': np, 'bn': type('bn', (), {'nanmean': staticmethod(np.nanmean)})}
    exec(body, ns)
    rng = np.random.default_rng(5)
    for _ in range(5):
        h = pd.Series(rng.normal(size=40))
        o = pd.Series(rng.normal(size=40))
        h[3] = np.nan
        for centered in (True, False):
            for crop in (None, 2):
                a = ns['crmse'](h.copy(), o.copy(), centered, crop)
                b = scores.crmse(h, o, centered, crop)
                assert a == pytest.approx(b, rel=1e-13)
