# -*- coding: utf-8 -*-
"""Detection and correction of inhomogeneities against the local PDFs.

Drop-in for ``detect`` and ``fill_station`` of the reference's
``GSIMCLI/tools/homog.py`` (:27-276, :279-337) plus the small station
bookkeeping around them (``list_stations``, ``take_candidate``,
``update_station``, :340-545).  ``detect`` keeps the reference's signature and
return value; the local PDFs come from ``GridFiles.stats_area`` (K2c on the
GPU) and the fill / flag / substitute step runs as the fused K3 kernel
(``gsb_detect``).  ``detect_many`` is the batched extension: all stations of a
network against one stack with one K2c and one K3 launch.
"""
import numpy as np
import pandas as pd

from .. import _capi
from . import grid as gr

_STATS_NAMES = {'lmean': 'mean', 'lmed': 'median', 'lskew': 'skewness',
                'lvar': 'variance', 'lstd': 'std', 'lcoefvar': 'coefvar',
                'lperc': 'lperc'}


def _select_stats(method, skewness, optional_stats):
    """Statistics required by the correction method (homog.py:136-167)."""
    if method == 'mean':
        sel = dict(lmean=True, lmed=False, lskew=False)
    elif method == 'median':
        sel = dict(lmean=False, lmed=True, lskew=False)
    elif method == 'skewness' and skewness:
        sel = dict(lmean=True, lmed=True, lskew=True)
    elif method == 'percentile':
        sel = dict(lmean=False, lmed=False, lskew=False)
    else:
        raise ValueError('Method {0} invalid or incomplete.'.format(method))
    if optional_stats:
        sel.update(optional_stats)
    sel['lmean'] = True     # always needed to fill missing data (homog.py:165)
    sel['lperc'] = True     # always needed for the detection (homog.py:167)
    return sel


def _prepare_obs(obs, grids, header):
    """Steps (2)-(5) of detect on one candidate: first valid coordinates,
    Flag removal, nodata count and nodata -> NaN (homog.py:169-188).  Works on
    the point-set's rows as an array: no DataFrame is touched."""
    if not isinstance(obs, gr.PointSet):
        path = obs
        obs = gr.PointSet()
        obs.load(path, header=header)
    arr = obs.array()
    names = obs.columns()
    index = obs.row_index()
    valid = ~np.isnan(arr).all(axis=1)
    first = int(np.argmax(valid)) if valid.any() else 0
    obs_xy = [arr[first, names.index('x')], arr[first, names.index('y')]]
    if 'Flag' in names:
        k = names.index('Flag')
        arr = np.delete(arr, k, axis=1)
        names.pop(k)
    if 'Flag' in obs.varnames:
        obs.varnames.remove('Flag')
        obs.nvars -= 1
    nodatas = int((arr[:, names.index('clim')] == obs.nodata).sum())
    arr[arr == obs.nodata] = np.nan
    obs.varnames = names
    obs.set_array(arr, index)
    return obs, obs_xy, nodatas


def _table(pset):
    """{column name: 1-D array} view of a stats_area point-set."""
    arr = pset.array()
    return {nm: arr[:, i] for i, nm in enumerate(pset.columns())}


def _finish(obs, local, fixcols, method, skewness, hom_in, flag,
            optional_stats, grids, obs_xy, rad):
    """Steps (7)-(12): K3 and assembly of the homogenised PointSet
    (homog.py:204-269).  ``local`` maps column names to arrays."""
    filled, homog, flagcol, det = hom_in
    arr = obs.array()
    names = obs.columns()
    arr[:, names.index('clim')] = filled
    obs.set_array(arr, obs.row_index())
    # placeholders of optional statistics are columns holding NaN
    # (dropna(axis=1), homog.py:212)
    keep = ~np.isnan(arr).any(axis=0)
    varnames = [nm for i, nm in enumerate(names) if keep[i]]
    cols = [arr[:, i] for i in range(len(names)) if keep[i]]
    homog = np.asarray(homog, dtype=np.float64)
    if 'clim' in varnames:
        cols[varnames.index('clim')] = homog
    else:
        varnames.append('clim')
        cols.append(homog)
    nvars = obs.nvars

    def add_var(values, varname):            # PointSet.add_var (grid.py:109-130)
        if varname in varnames:
            varname += '_new'
        varnames.append(varname)
        cols.append(np.asarray(values, dtype=np.float64))

    if flag:
        add_var(flagcol, 'Flag')
        nvars += 1
    if optional_stats:
        for key, value in optional_stats.items():
            if value and key == 'lperc':
                if 'median' in local:
                    med = local['median']
                else:
                    med = _table(grids.stats_area(obs_xy, rad, lmed=True,
                                                  save=False))['median']
                with np.errstate(invalid='ignore'):
                    pdet = np.where(filled >= med, local['rperc'],
                                    local['lperc'])
                add_var(pdet, 'pdet')
                nvars += 1
            elif value:
                name = _STATS_NAMES[key]
                add_var(local[name], name)
                nvars += 1
    homogenised = gr.PointSet(obs.name + '_homogenised', obs.nodata, nvars,
                              varnames, np.column_stack(cols))
    homogenised.set_array(homogenised.array(), obs.row_index())
    return homogenised, int(det)


def _fix_columns(method, local, vline_perc):
    if method == 'mean':
        return local['mean'], None, None
    if method == 'median':
        return local['median'], None, None
    if method == 'skewness':
        return local['median'], local['mean'], local['skewness']
    return vline_perc['rperc'], vline_perc['lperc'], None


def detect(grids, obs_file, rad=0, method='mean', prob=0.95, skewness=None,
           percentile=None, flag=True, save=False, outfile=None, header=True,
           optional_stats=None):
    """Detect and correct irregularities of one candidate station.

    Same contract as the reference (homog.py:27-276): an observation is
    flagged when it lies outside the interval of probability ``prob`` centred
    in the local PDF (quantiles ``(1-prob)/2`` and ``1-(1-prob)/2`` of the
    simulated values at the station's node, both ends inclusive); flagged
    values are replaced by the local mean, median, the skewness-selected one
    of the two, or the nearest bound of the ``percentile`` interval.  Missing
    observations (``nodata`` or absent years) are first filled with the local
    mean and then tested like the others.

    Returns ``(homogenised PointSet, detected_number, missing_data)``; the
    PointSet has the candidate's columns, ``clim`` corrected, a ``Flag``
    column (observed value where corrected, ``nodata`` elsewhere) and the
    requested optional statistics.
    """
    sel = _select_stats(method, skewness, optional_stats)
    obs, obs_xy, nodatas = _prepare_obs(obs_file, grids, header)
    local = _table(grids.stats_area(obs_xy, rad, p=prob, save=save, **sel))

    meanvalues = local['mean']
    fn = 0
    if obs.array().shape[0] < grids.dz:
        obs, fn = fill_station(obs, meanvalues, grids.zi, grids.zi + grids.dz,
                               grids.cellz, header)
    _check_series(obs, grids.dz)
    if method == 'percentile' and percentile != prob:
        grids.reset_read()
        vline_perc = _table(grids.stats_area(obs_xy, rad, lperc=True,
                                             p=percentile, save=False))
    else:
        vline_perc = local
    fix_a, fix_b, fix_c = _fix_columns(method, local, vline_perc)
    clim = obs.array()[:, obs.columns().index('clim')]
    as2d = lambda a: None if a is None else np.asarray(a, np.float64)[None, :]
    filled, homog, flagcol, det = _capi.detect(
        as2d(clim), as2d(meanvalues), as2d(local['lperc']),
        as2d(local['rperc']), as2d(fix_a), as2d(fix_b), as2d(fix_c),
        method=method, thr=float(skewness or 0.0), nodata=obs.nodata,
        device=grids.device)
    homogenised, detected_number = _finish(
        obs, local, None, method, skewness,
        (filled[0], homog[0], flagcol[0], det[0]), flag, optional_stats, grids,
        obs_xy, rad)
    if save and outfile:
        homogenised.save(outfile, header)
    return homogenised, detected_number, fn + nodatas


def detect_many(grids, obs_list, rad=0, method='mean', prob=0.95,
                skewness=None, percentile=None, flag=True, header=True,
                optional_stats=None, stream=None):
    """Extension: ``detect`` for many candidate stations against ONE stack
    (the ``interface/simstats.py:204-209`` calling pattern) with one K2c and
    one K3 launch.  Returns a list of ``detect``-style tuples."""
    sel = _select_stats(method, skewness, optional_stats)
    prepared = [_prepare_obs(o, grids, header) for o in obs_list]
    locs = [p[1] for p in prepared]
    locals_ = [_table(ps) for ps in
               grids.stats_area_many(locs, rad, p=prob, stream=stream, **sel)]
    if method == 'percentile' and percentile != prob:
        vperc = [_table(ps) for ps in
                 grids.stats_area_many(locs, rad, lperc=True, p=percentile,
                                       stream=stream)]
    else:
        vperc = locals_
    S, dz = len(prepared), grids.dz
    cols = [np.empty((S, dz)) for _ in range(7)]
    fns, obs_filled = [], []
    for s, (obs, _, _) in enumerate(prepared):
        local = locals_[s]
        fn = 0
        if obs.array().shape[0] < dz:
            obs, fn = fill_station(obs, local['mean'], grids.zi,
                                   grids.zi + grids.dz, grids.cellz, header)
        _check_series(obs, dz)
        fns.append(fn)
        obs_filled.append(obs)
        fa, fb, fc = _fix_columns(method, local, vperc[s])
        cols[0][s] = obs.array()[:, obs.columns().index('clim')]
        cols[1][s] = local['mean']
        cols[2][s] = local['lperc']
        cols[3][s] = local['rperc']
        cols[4][s] = fa
        cols[5][s] = fa if fb is None else fb
        cols[6][s] = fa if fc is None else fc
    filled, homog, flagcol, det = _capi.detect(
        *cols, method=method, thr=float(skewness or 0.0),
        nodata=prepared[0][0].nodata if prepared else -999.9,
        device=grids.device, stream=stream)
    out = []
    for s, obs in enumerate(obs_filled):
        hom, dn = _finish(obs, locals_[s], None, method, skewness,
                          (filled[s], homog[s], flagcol[s], det[s]), flag,
                          optional_stats, grids, locs[s], rad)
        out.append((hom, dn, fns[s] + prepared[s][2]))
    return out


def detect_sweep(grids, obs_list, probs, rad=0, method='mean', skewness=None,
                 flag=True, header=True, optional_stats=None, stream=None):
    """Extension: ``detect_many`` for several detection probabilities at once
    (the per-network probability sweep of batch homogenisation).  The station
    columns are reduced ONCE -- a sweep over ``prob`` only changes which sorted
    ranks are read -- so one K2c launch serves every probability and only the
    tiny K3 step is repeated.  Returns ``{prob: [detect-style tuples]}``; each
    entry equals ``detect_many(..., prob=prob)``."""
    if method == 'percentile':
        raise ValueError('detect_sweep: the percentile method has its own '
                         'interval; call detect_many per probability')
    probs = [float(p) for p in probs]
    if not probs:
        return {}
    sel = _select_stats(method, skewness, optional_stats)
    prepared = [_prepare_obs(o, grids, header) for o in obs_list]
    locs = [p[1] for p in prepared]
    extra = []
    for p in probs[1:]:
        extra += [(1 - p) / 2, 1 - (1 - p) / 2]            # grid.py:700-701
    psets = grids.stats_area_many(locs, rad, p=probs[0], stream=stream,
                                  percentiles=extra or None, **sel)
    S, dz = len(prepared), grids.dz
    base, bounds, fns, filled_obs = [], [], [], []
    for s, (obs, _, _) in enumerate(prepared):
        arr = psets[s].array()
        names = psets[s].columns()
        nx = len(extra)
        core = {nm: arr[:, i] for i, nm in enumerate(names[:len(names) - nx])}
        pairs = [(core['lperc'], core['rperc'])]
        for k in range(len(probs) - 1):
            pairs.append((arr[:, arr.shape[1] - nx + 2 * k],
                          arr[:, arr.shape[1] - nx + 2 * k + 1]))
        fn = 0
        if obs.array().shape[0] < dz:
            obs, fn = fill_station(obs, core['mean'], grids.zi,
                                   grids.zi + grids.dz, grids.cellz, header)
        _check_series(obs, dz)
        base.append(core)
        bounds.append(pairs)
        fns.append(fn)
        filled_obs.append(obs)
    nodata = prepared[0][0].nodata if prepared else -999.9
    out = {}
    for k, p in enumerate(probs):
        cols = [np.empty((S, dz)) for _ in range(7)]
        locals_k = []
        for s, obs in enumerate(filled_obs):
            local = dict(base[s])
            local['lperc'], local['rperc'] = bounds[s][k]
            locals_k.append(local)
            fa, fb, fc = _fix_columns(method, local, local)
            cols[0][s] = obs.array()[:, obs.columns().index('clim')]
            cols[1][s] = local['mean']
            cols[2][s] = local['lperc']
            cols[3][s] = local['rperc']
            cols[4][s] = fa
            cols[5][s] = fa if fb is None else fb
            cols[6][s] = fa if fc is None else fc
        filled, homog, flagcol, det = _capi.detect(
            *cols, method=method, thr=float(skewness or 0.0), nodata=nodata,
            device=grids.device, stream=stream)
        res = []
        for s, obs in enumerate(filled_obs):
            work = gr.PointSet(obs.name, obs.nodata, obs.nvars, list(obs.varnames),
                               obs.array())
            work.set_array(work.array(), obs.row_index())
            hom, dn = _finish(work, locals_k[s], None, method, skewness,
                              (filled[s], homog[s], flagcol[s], det[s]), flag,
                              optional_stats, grids, locs[s], rad)
            res.append((hom, dn, fns[s] + prepared[s][2]))
        out[p] = res
    return out


def _check_series(obs, dz):
    """The filled series must have one row per grid time step.  The reference
    fails in pandas when it does not (`meanvalues.index = obs.values.index`,
    homog.py:197, raises ValueError: Length mismatch); same exception here,
    before any device buffer is sized from the wrong length."""
    rows = obs.array().shape[0]
    if rows != dz:
        raise ValueError('Length mismatch: Expected axis has {0} elements, new '
                         'values have {1} elements'.format(dz, rows))


def fill_station(pset_file, values, time_min, time_max, time_step=1,
                 header=True):
    """Expand an incomplete station series to every time step of the grid,
    filling the synthesised rows with ``values`` (homog.py:279-337).

    Rows are matched to ``arange(time_min, time_max, time_step)`` by a merge
    walk; a synthesised row copies x, y and the columns between ``time`` and
    ``clim`` (the station id) from the first row.  Returns ``(PointSet,
    number of rows added)``.
    """
    if isinstance(pset_file, gr.PointSet):
        pset = pset_file
    else:
        pset = gr.PointSet()
        pset.load(pset_file, header=header)
    names = pset.columns()
    varcol = pset.varnames.index('clim')
    times = np.arange(time_min, time_max, time_step)
    rows = pset.array()
    have = rows[:, names.index('time')]
    ncol = rows.shape[1]
    # a synthesised row: x, y of the first row, the time, the columns between
    # `time` and `clim` of the first row (station id), then the fill value
    template = np.concatenate([rows[0, :2], [0.0], rows[0, 3:varcol], [0.0]])
    fillpos = 3 + max(varcol - 3, 0)
    nrows = have.shape[0]
    pos = np.searchsorted(times, have)
    regular = (nrows == 0) or (
        bool(np.all(np.diff(have) > 0)) and int(pos.max()) < times.shape[0]
        and bool(np.all(times[pos] == have)))
    if regular:
        # the usual case -- observed years strictly increasing and all on the
        # grid's time axis: the merge walk below matches row j to its year
        filled = np.empty((times.shape[0], ncol))
        filled[:] = template[:ncol]
        filled[:, 2] = times
        filled[:, fillpos] = np.asarray(values, dtype=np.float64)[:times.shape[0]]
        if nrows:
            filled[pos] = rows
        added = times.shape[0] - nrows
    else:
        filled = np.zeros((times.shape[0], ncol))
        j = 0
        added = 0
        for i, t in enumerate(times):
            if j < nrows and t == have[j]:
                filled[i] = rows[j]
                j += 1
            else:
                template[2] = t
                template[fillpos] = values[i]
                filled[i] = template[:ncol]
                added += 1
    pset.varnames = names if len(names) == ncol else pset.varnames
    pset.set_array(filled)
    return pset, added


def list_stations(pset_file, header=True, variables=None):
    """Station ids present in a point-set (homog.py:340-380)."""
    pset = gr.loadcheck(pset_file, header)
    ids = pset.values['station'].unique()
    return [int(i) for i in ids]


def take_candidate(pset_file, station, header=True, save=False, path=None):
    """Split a network point-set into the candidate station and its
    references, dropping Flag and optional-statistics columns
    (homog.py:442-504)."""
    pset = gr.loadcheck(pset_file, header)
    is_cand = pset.values['station'] == int(station)
    drop = ['Flag', 'mean', 'median', 'std', 'pdet', 'variance', 'coefvar',
            'skewness']
    cand = pset.values[is_cand].drop(drop, axis=1, errors='ignore')
    refs = pset.values[~is_cand].drop(drop, axis=1, errors='ignore')
    names = list(cand.columns)
    cand_pset = gr.PointSet('Candidate_' + str(station), pset.nodata,
                            len(names), list(names), cand)
    refs_pset = gr.PointSet('References_' + str(station), pset.nodata,
                            len(names), list(names), refs)
    if save and path:
        import os
        base, ext = os.path.splitext(path)
        cand_pset.save(base + '_candidate' + ext, header)
        refs_pset.save(base + '_neighbours' + ext, header)
    return cand_pset, refs_pset


def update_station(stations_pset, station, header=True):
    """Write a homogenised station back into the network point-set, adding
    any new columns as NaN first (homog.py:507-545; index-aligned, like
    ``DataFrame.update``)."""
    pset = gr.loadcheck(stations_pset, header)
    if sorted(pset.varnames) != sorted(station.varnames):
        for var in [v for v in station.varnames if v not in pset.varnames]:
            pset.add_var(np.repeat(np.nan, pset.values.shape[0]), var)
    pset.values.update(station.values)
    return pset
