# -*- coding: utf-8 -*-
"""Scores of a homogenisation run (the reference's ``tools/scores.py``: the
COST-HOME benchmark measures of Venema et al. 2012): centred root-mean-square
error of homogenised against original (true) series, per station, per network
and per submission, and the improvement over the inhomogeneous data.

The reference computes these on its COST-HOME directory objects
(``parsers/costhome.py``: Station / Network / Submission), which are outside
this package's scope; the arithmetic is restated here on plain tables --
a network is a ``DataFrame`` indexed by year with one column per station -- and
``load_gsimcli_results`` reads the ``*_homogenised_data.csv`` that
``homog.save_output(fformat='gsimcli')`` writes, so the output of the candidate
loop can be scored directly.  Host-side analytics: nothing here touches the GPU.
"""
import numpy as np
import pandas as pd


def crmse(homog, orig, centered=True, crop=None):
    """Centred RMSE between a homogenised and an original data set
    (scores.py:26-70): both are centred on their own mean (missing values
    ignored) unless ``centered`` is False; ``crop`` drops that many first and
    last years.  Series, DataFrames or arrays."""
    h = np.array(getattr(homog, 'values', homog), dtype=np.float64)
    o = np.array(getattr(orig, 'values', orig), dtype=np.float64)
    if crop is not None:
        h, o = h[crop:-crop], o[crop:-crop]
    if centered:
        h = h - np.nanmean(h)
        o = o - np.nanmean(o)
    return float(np.sqrt(np.nanmean(np.power(h - o, 2))))


def crmse_station(homog, orig, crop=None):
    """CRMSE of one station's (yearly) series (scores.py:73-110)."""
    return crmse(homog, orig, True, crop)


def crmse_network(homog, orig, skip_missing=False):
    """Network CRMSE (scores.py:113-147): the CRMSE of the network-average
    series.  ``homog`` / ``orig``: DataFrames, years x stations.
    ``skip_missing`` drops the years in which any station has no value."""
    homog, orig = pd.DataFrame(homog), pd.DataFrame(orig)
    if skip_missing:
        keep = ~(homog.isnull().any(axis=1) | orig.isnull().any(axis=1))
        homog, orig = homog[keep], orig[keep]
    return crmse(homog.mean(axis=1), orig.mean(axis=1))


def crmse_submission(networks, over_station=True, over_network=True,
                     skip_missing=False):
    """Mean network CRMSE and / or mean station CRMSE of a set of networks
    (scores.py:150-216).  ``networks``: ``{network id: (homog, orig)}`` with
    DataFrames years x stations.  Returns the list the reference returns:
    ``[mean network CRMSE, mean station CRMSE]`` (those requested)."""
    net, st_sum, st_n = [], 0.0, 0
    for key in networks:
        homog, orig = networks[key]
        if over_network:
            net.append(crmse_network(homog, orig, skip_missing))
        if over_station:
            for col in homog.columns:
                st_sum += crmse_station(homog[col], orig[col])
                st_n += 1
    out = []
    if over_network:
        out.append(float(np.mean(net)))
    if over_station:
        out.append(st_sum / st_n)
    return out


def improvement(homog_networks, inho_networks, **kwargs):
    """Improvement over the inhomogeneous data (scores.py:219-267): quotient of
    the CRMSE of the homogenised networks and of the same networks before
    homogenisation (both against the original data).  Returns ``(homogenised
    CRMSEs, inhomogeneous CRMSEs, improvements)``."""
    h = crmse_submission(homog_networks, **kwargs)
    i = crmse_submission(inho_networks, **kwargs)
    return h, i, list(np.array(h) / np.array(i))


def load_gsimcli_results(path, no_data=-999.9, variable='clim'):
    """Years x stations table of the homogenised values in a
    ``*_homogenised_data.csv`` (``save_output(fformat='gsimcli')``: columns
    ``<station>_clim, <station>_Flag, ...``); ``no_data`` becomes NaN."""
    df = pd.read_csv(path, index_col=0)
    cols = [c for c in df.columns if c.endswith('_' + variable)]
    out = df[cols].copy()
    out.columns = [int(float(c[:-len(variable) - 1])) for c in cols]
    return out.replace(no_data, np.nan).astype(float)


def gsimcli_improvement(gsimcli_results, orig, inho=None, no_data=-999.9, **kwargs):
    """Scores of a GSIMCLI run (scores.py:270-340).  ``gsimcli_results``:
    ``{network id: path of its *_homogenised_data.csv}``; ``orig`` / ``inho``:
    ``{network id: DataFrame years x stations}`` of the true and of the
    inhomogeneous data.  With ``inho`` the result is ``improvement(...)``,
    otherwise ``[crmse_submission(...)]``."""
    homog = {}
    for key, path in gsimcli_results.items():
        table = load_gsimcli_results(path, no_data)
        ref = orig[key]
        homog[key] = (table.reindex(index=ref.index, columns=ref.columns), ref)
    if inho:
        inh = {key: (inho[key].reindex(index=orig[key].index, columns=orig[key].columns),
                     orig[key]) for key in gsimcli_results}
        return improvement(homog, inh, **kwargs)
    return [crmse_submission(homog, **kwargs)]
