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


class DetectPlan(object):
    """Extension: the station path of one network prepared once and kept on
    the device, so that a step is launches only.

    ``detect`` does, per call, host work that depends on the candidates but not
    on the realizations (load / clean the point-sets, expand them to the grid's
    time axis, find the station nodes -- homog.py:169-188, 279-337,
    grid.py:793-804).  The plan does it once and uploads the result: node ids,
    observations (NaN where missing: K3 fills those with the local mean,
    homog.py:200-201).  ``enqueue`` then runs, on one stream and without a host
    round trip,

        K2c  (gsb_stats_columns_device: local PDFs at the station nodes)
        [all-gather of the per-rank table slabs, NCCL, node-sharded stacks]
        K3   (gsb_detect_table: fill / flag / substitute)
        D2H  of the [3][S][dz] result + the S detection counts (pinned)

    and ``fetch`` waits for that copy.  ``results`` assembles the reference's
    return values (point-sets) from it -- host glue, outside the step.
    """

    def __init__(self, grids, obs_list, rad=0, method='mean', prob=0.95,
                 skewness=None, percentile=None, header=True,
                 optional_stats=None, group=None):
        import torch
        self.torch = torch
        self.grids, self.rad, self.method = grids, rad, method
        self.prob, self.skewness = prob, skewness
        self.optional_stats = optional_stats
        sel = _select_stats(method, skewness, optional_stats)
        self.flags = _capi.flags_from(**sel)
        names = [nm for key, nm in gr._AREA_COLUMNS if sel.get(key)]
        names += ['lperc', 'rperc']
        self.names = names
        C = len(names)
        col = names.index
        if method == 'mean':
            c_fa, c_fb, c_fc = col('mean'), col('mean'), -1
        elif method == 'median':
            c_fa, c_fb, c_fc = col('median'), col('median'), -1
        elif method == 'skewness':
            c_fa, c_fb, c_fc = col('median'), col('mean'), col('skewness')
        else:
            c_fa, c_fb, c_fc = col('rperc'), col('lperc'), -1
        self.two_tables = method == 'percentile' and percentile != prob
        self.percentile = percentile
        self.cidx = (col('mean'), col('lperc'), col('rperc'), c_fa, c_fb, c_fc)
        if self.two_tables:          # the correction interval has its own table
            self.cidx = (col('mean'), col('lperc'), col('rperc'), 1, 0, -1)
        # ---- host preparation, once
        stack = grids._need_stack()
        dz = grids.dz
        self.prepared, self.fns, self.filled_obs, self.locs = [], [], [], []
        clim = np.empty((len(obs_list), dz))
        nans = np.full(dz, np.nan)
        for s, o in enumerate(obs_list):
            obs, xy, nodatas = _prepare_obs(o, grids, header)
            fn = 0
            if obs.array().shape[0] < dz:
                obs, fn = fill_station(obs, nans, grids.zi, grids.zi + dz,
                                       grids.cellz, header)
            _check_series(obs, dz)
            clim[s] = obs.array()[:, obs.columns().index('clim')]
            self.prepared.append((obs, xy, nodatas))
            self.fns.append(fn)
            self.locs.append(xy)
        S = self.S = len(obs_list)
        self.nodata = self.prepared[0][0].nodata if S else -999.9
        nodes, ids = [], []
        for loc in self.locs:
            node, nid = grids._area_nodes(loc, rad)
            nodes.append(node)
            ids.append(nid)
        self.nodes = nodes
        kmax = max(i.shape[1] for i in ids) if ids else 1
        ids = (np.stack([np.pad(i, ((0, 0), (0, kmax - i.shape[1])),
                                constant_values=-1) for i in ids])
               if ids else np.zeros((0, dz, 1), dtype=np.int64))
        self.K = kmax
        # ---- layer slabs of the ranks (whole time layers, gathered once)
        dev = torch.device('cuda', grids.device)
        self.dev = dev
        per = grids.dx * grids.dy
        self.group = group
        self.world = 1
        if grids.sharded:
            import torch.distributed as td
            if not (td.is_available() and td.is_initialized()):
                raise RuntimeError(
                    'this stack holds a node range only: the station tables of '
                    'the other ranks are needed (torch.distributed is not '
                    'initialised)')
            if grids.node0 % per or (stack.N % per and grids.node0 + stack.N
                                     != per * dz):
                raise ValueError('a node-sharded stack must hold whole time '
                                 'layers (node range {0}..{1}, {2} nodes per '
                                 'layer)'.format(grids.node0,
                                                 grids.node0 + stack.N, per))
            self.world = td.get_world_size(group)
            z0 = grids.node0 // per
            z1 = min(dz, -(-(grids.node0 + stack.N) // per))
            mine = torch.tensor([z0, z1], dtype=torch.int64, device=dev)
            allr = torch.empty((self.world, 2), dtype=torch.int64, device=dev)
            td.all_gather_into_tensor(allr, mine, group=group)
            ranges = [(int(a), int(b)) for a, b in allr.cpu().tolist()]
            cover = np.zeros(dz, dtype=np.int64)
            for a, b in ranges:
                cover[a:b] += 1
            if not np.all(cover == 1):
                raise ValueError('the ranks\' time-layer slabs {0} do not tile '
                                 'the {1} layers exactly once'.format(ranges, dz))
        else:
            z0, z1 = 0, dz
            ranges = [(0, dz)]
        self.z0, self.z1 = z0, z1
        width = self.width = max(b - a for a, b in ranges)
        # this rank's layers, padded to the widest slab with absent nodes
        mine_ids = np.full((S, width, kmax), -1, dtype=np.int64)
        mine_ids[:, :z1 - z0] = ids[:, z0:z1]
        zoff = np.empty(dz, dtype=np.int64)
        for r, (a, b) in enumerate(ranges):
            zoff[a:b] = r * S * width + np.arange(b - a)
        self.C = C
        f64 = torch.float64
        self.d_ids = torch.from_numpy(mine_ids).to(dev)
        self.d_zoff = torch.from_numpy(zoff).to(dev)
        self.d_clim = torch.from_numpy(clim).to(dev)
        self.d_part = torch.empty((S, width, C), dtype=f64, device=dev)
        self.d_full = (torch.empty((self.world, S, width, C), dtype=f64, device=dev)
                       if self.world > 1 else self.d_part)
        if self.two_tables:
            self.d_part2 = torch.empty((S, width, 2), dtype=f64, device=dev)
            self.d_full2 = (torch.empty((self.world, S, width, 2), dtype=f64,
                                        device=dev)
                            if self.world > 1 else self.d_part2)
        self.d_out = torch.empty((3, S, dz), dtype=f64, device=dev)
        self.d_det = torch.zeros((S,), dtype=torch.int32, device=dev)
        self.h_out = torch.empty((3, S, dz), dtype=f64).pin_memory()
        self.h_det = torch.zeros((S,), dtype=torch.int32).pin_memory()
        self.done = torch.cuda.Event()
        self._q = np.zeros(0, dtype=np.float64)

    def enqueue(self, stream=None):
        """Enqueue one station step on ``stream`` (a ``torch.cuda.Stream``;
        default: the current stream).  No host synchronisation."""
        torch = self.torch
        L = _capi.lib()
        st = stream if stream is not None else torch.cuda.current_stream(self.dev)
        h = st.cuda_stream
        stack = self.grids._need_stack()
        S, dz, C = self.S, self.grids.dz, self.C
        if S == 0:
            return
        qp = self._q.ctypes.data_as(_capi.p_f64)
        nd = self.grids.nodata
        _capi.check(L.gsb_stats_columns_device(
            stack.handle, self.d_ids.data_ptr(), S, self.width, self.K, self.flags,
            self.prob, qp, 0, nd, self.d_part.data_ptr(), h))
        if self.two_tables:
            _capi.check(L.gsb_stats_columns_device(
                stack.handle, self.d_ids.data_ptr(), S, self.width, self.K,
                _capi.PERC, self.percentile, qp, 0, nd, self.d_part2.data_ptr(), h))
        if self.world > 1:
            import torch.distributed as td
            with torch.cuda.stream(st):
                td.all_gather_into_tensor(self.d_full, self.d_part, group=self.group)
                if self.two_tables:
                    td.all_gather_into_tensor(self.d_full2, self.d_part2,
                                              group=self.group)
        fix = self.d_full2 if self.two_tables else self.d_full
        cm, cl, cr, fa, fb, fc = self.cidx
        _capi.check(L.gsb_detect_table(
            self.d_clim.data_ptr(), self.d_full.data_ptr(), fix.data_ptr(),
            self.d_zoff.data_ptr(), self.width, C, 2 if self.two_tables else C,
            cm, cl, cr, fa, fb, fc, S, dz, _capi.METHODS[self.method],
            float(self.skewness or 0.0), self.nodata, self.d_out.data_ptr(),
            self.d_det.data_ptr(), self.grids.device, h))
        with torch.cuda.stream(st):
            self.h_out.copy_(self.d_out, non_blocking=True)
            self.h_det.copy_(self.d_det, non_blocking=True)
            self.done.record(st)

    def fetch(self):
        """Wait for the last enqueued step; ``(filled, homog, flag, detected)``
        as NumPy views of the pinned result buffers."""
        self.done.synchronize()
        o = self.h_out.numpy()
        return o[0], o[1], o[2], self.h_det.numpy()

    def table(self):
        """The full ``[S][dz][C]`` station table of the last step on the host
        (columns ``self.names``)."""
        torch = self.torch
        self.done.synchronize()
        full = self.d_full.cpu().numpy().reshape(-1, self.C)
        zoff = self.d_zoff.cpu().numpy()
        rows = zoff[None, :] + np.arange(self.S)[:, None] * self.width
        return full[rows]

    def results(self, flag=True):
        """The reference's return values for the last step: a list of
        ``(homogenised PointSet, detected_number, missing_data)``."""
        filled, homog, flagcol, det = self.fetch()
        table = self.table()
        out = []
        for s, (obs, xy, nodatas) in enumerate(self.prepared):
            local = {nm: table[s, :, i] for i, nm in enumerate(self.names)}
            # each call assembles from a fresh copy of the prepared series
            fresh = gr.PointSet(obs.name, obs.nodata, obs.nvars, list(obs.varnames),
                                obs.array().copy())
            fresh.set_array(fresh.array(), obs.row_index())
            hom, dn = _finish(fresh, local, None, self.method, self.skewness,
                              (filled[s].copy(), homog[s].copy(), flagcol[s].copy(),
                               det[s]), flag, self.optional_stats, self.grids, xy,
                              self.rad)
            out.append((hom, dn, self.fns[s] + nodatas))
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


def detect_batch(networks, probs, rad=0, method='mean', skewness=None,
                 flag=True, header=True, optional_stats=None, group=None):
    """Extension (BASELINE config 5, the batch mode of
    launchers/method_classic.py:461-534 with its loop at :506): many independent
    (stack, candidate stations) problems, each with a sweep of detection
    probabilities.

    ``networks`` is a list of ``(make_grids, obs_list)``; ``make_grids(device)``
    returns the network's ``GridFiles`` on that device (called only on the rank
    that owns the network, so that 100 stacks never have to be resident at
    once) and is dumped after use.  Whole networks are dealt round-robin to the
    ranks of ``torch.distributed`` (replicas: no data-path collective); the
    results are exchanged at the end with one object gather.  Returns, on
    every rank, ``[{prob: [detect-style tuples]} for each network]``.
    """
    import torch
    import torch.distributed as td
    dist_on = td.is_available() and td.is_initialized()
    world = td.get_world_size(group) if dist_on else 1
    rank = td.get_rank(group) if dist_on else 0
    device = torch.cuda.current_device()
    mine = {}
    for i, (make_grids, obs_list) in enumerate(networks):
        if i % world != rank:
            continue
        grids = make_grids(device)
        try:
            mine[i] = detect_sweep(grids, obs_list, probs, rad=rad, method=method,
                                   skewness=skewness, flag=flag, header=header,
                                   optional_stats=optional_stats)
        finally:
            grids.dump()
    if world == 1:
        return [mine[i] for i in range(len(networks))]
    parts = [None] * world
    td.all_gather_object(parts, mine, group=group)
    merged = {}
    for part in parts:
        merged.update(part)
    return [merged[i] for i in range(len(networks))]


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
    if variables:
        pset.flush_varnames(variables)
    ids = np.unique(pset.values['station'])         # sorted, like the reference
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


def station_order(method, pset_file=None, nd=-999.9, header=True,
                  userset=None, ascending=False, md_last=True):
    """Order in which the stations of a network are homogenised
    (homog.py:548-623).  ``method``: 'random', 'sorted', 'variance' (variance
    of each station's series), 'network deviation' (|station mean - network
    mean|) or 'user' (``userset`` as given).  Missing data (``nd``) is ignored
    in the statistics; stations without any statistic go last (``md_last``) or
    first."""
    from random import shuffle
    pset = None
    if pset_file is not None:
        if isinstance(pset_file, gr.PointSet):
            pset = pset_file
        else:
            pset = gr.PointSet(psetpath=pset_file, nodata=nd, header=header)
    if method == 'user' and userset:
        return userset
    if pset is None:
        if method in ('variance', 'network deviation', 'random', 'sorted'):
            raise TypeError('Method {0} requires the stations point-set'
                            .format(method))
        raise TypeError('Method {0} not understood or invalid userset ({1}).'
                        .format(method, userset))
    stations_list = list_stations(pset, header=header)
    na_position = 'last' if md_last else 'first'
    if method == 'random':
        shuffle(stations_list)
    elif method == 'sorted':
        stations_list.sort(reverse=not ascending)
    elif method == 'variance':
        values = pset.values.replace(nd, np.nan)
        varsort = values.groupby('station', sort=False).clim.var()
        varsort = varsort.sort_values(ascending=ascending, na_position=na_position)
        stations_list = list(varsort.index)
    elif method == 'network deviation':
        values = pset.values.replace(nd, np.nan)
        stations_mean = values.groupby('station', sort=False).clim.mean()
        network_dev = ((stations_mean - values.clim.mean()).abs()
                       .sort_values(ascending=ascending, na_position=na_position))
        stations_list = list(network_dev.index)
    else:
        raise TypeError('Method {0} not understood or invalid userset ({1}).'
                        .format(method, userset))
    return stations_list


def save_output(pset_file, outfile, fformat='gsimcli', outvars=None,
                header=True, network_split=True, save_stations=False,
                keys=None, append_year=False):
    """Write homogenisation results (homog.py:626-760).

    ``fformat``: 'gsimcli' -- CSV indexed by year with the columns
    ``<station>_clim, <station>_Flag`` (+ optional statistics) per station;
    'normal' -- CSV of the point-set's columns (``outvars``: column indices);
    'gslib' -- GSLIB point-set with header.  ``save_stations`` adds
    ``<outfile>_stations.csv`` with the station ids and coordinates.
    ('cost-home' is not implemented in the reference either.)"""
    import csv
    import os
    if isinstance(pset_file, gr.PointSet):
        pset = pset_file
    else:
        pset = gr.PointSet()
        pset.load(pset_file, header)

    if fformat == 'normal':
        with open(outfile, 'w', newline='') as csvfile:
            out = csv.writer(csvfile, dialect='excel')
            if not outvars:
                out.writerow(pset.varnames)
                out.writerows(pset.values.values.tolist())
            else:
                out.writerow([pset.varnames[i] for i in outvars])
                out.writerows(pset.values.iloc[:, np.array(outvars)].values.tolist())
    elif fformat == 'gsimcli':
        stations = list_stations(pset, header)
        cols = ['clim', 'Flag']
        cols += sorted(set(pset.varnames)
                       - set(['x', 'y', 'time', 'station', 'clim', 'Flag']))
        cols = [c for c in cols if c in pset.values.columns]
        csvheader = [str(st) + '_' + var for st in stations for var in cols]
        vals = pset.values
        years = np.arange(vals['time'].min(), vals['time'].max() + 1)
        outdf = pd.DataFrame(index=years, columns=csvheader, dtype=object)
        for i, st in enumerate(stations):
            sel = vals['station'] == st
            temp = vals[sel][cols].copy()
            temp.index = vals[sel]['time']
            temp.columns = csvheader[len(cols) * i:len(cols) * (i + 1)]
            outdf.update(temp)
        if append_year:
            raise NotImplementedError
        outdf.dropna(axis=1, how='all', inplace=True)
        outdf.to_csv(outfile, index_label='year')
    elif fformat.lower() == 'gslib':
        if network_split and 'network' in pset.varnames:
            networks = pset.values['network'].unique()
            names = [v for v in pset.varnames if v != 'network']
            for nw in networks:
                part = pset.values[pset.values['network'] == nw].drop('network', axis=1)
                temp = gr.PointSet(pset.name + ' network: ' + str(nw), pset.nodata,
                                   len(names), list(names), part)
                base, ext = os.path.splitext(outfile)
                temp.save(base + '_' + str(nw) + ext, header=True)
        else:
            pset.save(outfile, header=True)
    else:
        raise ValueError('unknown output format {0!r}'.format(fformat))

    if save_stations:
        stations = list_stations(pset, header)
        stations_out = os.path.join(
            os.path.dirname(outfile),
            os.path.splitext(os.path.basename(outfile))[0] + '_stations.csv')
        rows = []
        for st in stations:
            temp = pset.values[pset.values['station'] == st]
            rows.append(temp.loc[temp.first_valid_index(), ['x', 'y']].values)
        stationsdf = pd.DataFrame(np.array(rows, dtype=float).reshape(len(stations), 2),
                                  index=stations, columns=['x', 'y'])
        if keys:
            stationsdf = stationsdf.join(pd.read_csv(keys, sep='\t', index_col=0))
        stationsdf.to_csv(stations_out, index_label='Station')
