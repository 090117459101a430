# -*- coding: utf-8 -*-
"""CPU restatement (NumPy/pandas) of gsimcli's local-PDF + detection hot path.

TEST INFRASTRUCTURE ONLY -- this is the parity oracle and the CPU baseline.
Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import it.  Nothing under ``gsimcli_b200/`` does.

Pinning: the reference ships no golden vectors for this path,
so this restatement is pinned against outputs of the reference's own source
executed through ``oracle/refshim.py`` -- see ``oracle/make_golden.py`` and
``tests/golden/*.json`` (committed), plus the Appendix-C known-answer vector.

Every function cites the reference lines it restates (paths relative to
``/root/reference/GSIMCLI``).  Arithmetic is float64 throughout, like the
reference (``np.zeros`` defaults, ``float(a)``), and third-party numerics
follow NumPy 2.3 / pandas 3.0 semantics:

* nodata      exact equality -> NaN                 (bn.replace, grid.py:681,829)
* mean        sum/n over non-NaN                    (bn.nanmean, grid.py:683,832)
* var/std     two-pass, ddof=1                      (bn.nanvar/nanstd, :689-691)
* coefvar     std/mean*100                          (grid.py:692-698,841-847)
* median      mean of the two middle values         (bn.nanmedian, :685,834)
* quantile    h=q(n-1), NumPy ``_lerp``             (pd.Series.quantile, :700,849)
* skew        adjusted Fisher-Pearson G1, pandas    (pd.Series.skew, :687,836)
              ``nanops.nanskew`` incl. its fp-error zeroing and n<3 -> NaN
* mode        EXTENSION (not in the reference): most frequent non-NaN value,
              ties -> smallest; all-unique -> minimum; empty -> NaN
"""
import numpy as np
import pandas as pd

FLAG_MEAN, FLAG_MED, FLAG_SKEW, FLAG_VAR = 1, 2, 4, 8
FLAG_STD, FLAG_COEFVAR, FLAG_PERC, FLAG_MODE = 16, 32, 64, 128

# ---------------------------------------------------------------------------
# indexing helpers (bit-exact requirement)
# ---------------------------------------------------------------------------


def coord_to_grid(coord, cells_size, first):
    """grid.py:921-954 -- world -> 1-based node, np.around (half-to-even)."""
    coord = list(coord)
    if len(coord) < len(first):
        coord.append(first[2])
    coord = np.array(coord).astype('float')
    return np.around((coord - np.array(first)) / np.array(cells_size)
                     + 1).astype('int')


def grid_to_line(coord, dims):
    """grid.py:957-979 -- 1-based file line of node (x, y, z)."""
    return (coord[0] + dims[0] * (coord[1] - 1)
            + dims[0] * dims[1] * (coord[2] - 1))


def line_zmirror(loc, dims):
    """grid.py:982-1002 -- the dz line numbers of the column at (x, y)."""
    return [grid_to_line([loc[0], loc[1], z], dims)
            for z in range(1, dims[2] + 1)]


def circle(xc, yc, r, dims=None):
    """grid.py:1156-1187 -- nodes of the disc of radius r, x-major order.

    The reference only drops negative indices (node 0 and nodes
    beyond dx/dy are kept and silently alias other lines).  With ``dims`` the
    *intended* behaviour is applied: nodes outside [1..dx] x [1..dy] are
    dropped.  ``dims=None`` reproduces the reference literally.
    """
    xx, yy = np.mgrid[xc - r:xc + r + 1, yc - r:yc + r + 1].astype('float32')
    xx[xx < 0] = np.nan
    yy[yy < 0] = np.nan
    inside = (xx - xc) ** 2 + (yy - yc) ** 2 <= r ** 2
    nodes = np.column_stack((xx[inside], yy[inside])).astype('int')
    if dims is not None:
        keep = ((nodes[:, 0] >= 1) & (nodes[:, 0] <= dims[0])
                & (nodes[:, 1] >= 1) & (nodes[:, 1] <= dims[1]))
        nodes = nodes[keep]
    return nodes


def area_lines(loc, tol, dims, first_coord, cells_size, clip=True):
    """grid.py:793-804 -- sorted K x dz matrix of 1-based line numbers."""
    node = coord_to_grid(loc, cells_size, first_coord)[:2]
    nodes = circle(node[0], node[1], tol, dims if clip else None)
    lines = np.array([line_zmirror(n, dims) for n in nodes], dtype=np.int64)
    lines = lines.reshape(len(nodes), dims[2])
    return node, np.sort(lines, axis=0)


# ---------------------------------------------------------------------------
# per-column statistics, scalar form (mirrors the reference's call sequence)
# ---------------------------------------------------------------------------


def replace_nodata(arr, nodata):
    """bn.replace(arr, nodata, nan) -- grid.py:681,829.  Exact equality in the
    precision the values are stored in."""
    arr = np.array(arr, copy=True)
    if arr.dtype == np.float32:
        mask = arr == np.float32(nodata)
        out = arr.astype(np.float64)
    else:
        out = arr.astype(np.float64)
        mask = out == nodata
    out[mask] = np.nan
    return out


def quantile_linear(sorted_valid, q):
    """pd.Series.quantile -> np.quantile(method='linear') (grid.py:699-701).

    ``h = (n-1)*q``; ``lo = floor(h)``; NumPy ``_lerp`` incl. the t>=0.5 branch.
    """
    n = len(sorted_valid)
    if n == 0:
        return np.nan
    h = (n - 1) * np.float64(q)
    if h >= n - 1:
        return np.float64(sorted_valid[-1])
    if h < 0:
        return np.float64(sorted_valid[0])
    lo = int(np.floor(h))
    t = h - lo
    a = np.float64(sorted_valid[lo])
    b = np.float64(sorted_valid[lo + 1])
    d = b - a
    if t >= 0.5:
        return b - d * (1 - t)
    return a + d * t


def skew_g1(valid):
    """pandas nanops.nanskew on the non-NaN values (grid.py:687,836)."""
    x = np.asarray(valid, dtype=np.float64)
    n = x.size
    if n < 3:
        return np.nan          # pandas: count < 3 -> NaN, whatever m2 is
    mean = x.sum() / n
    adj = x - mean
    adj2 = adj ** 2
    adj3 = adj2 * adj
    m2 = adj2.sum()
    m3 = adj3.sum()
    eps = np.finfo(np.float64).eps
    max_abs = np.abs(x).max()
    if abs(m2) < ((eps * max_abs) ** 2) * n:
        m2 = 0.0
    if abs(m3) < ((eps * max_abs) ** 3) * n:
        m3 = 0.0
    if m2 == 0:
        return 0.0
    with np.errstate(invalid='ignore', divide='ignore'):
        return (n * (n - 1) ** 0.5 / (n - 2)) * (m3 / m2 ** 1.5)


def mode_smallest(sorted_valid):
    """EXTENSION: most frequent value, ties -> smallest."""
    n = len(sorted_valid)
    if n == 0:
        return np.nan
    best_v, best_c = sorted_valid[0], 0
    i = 0
    while i < n:
        j = i
        while j + 1 < n and sorted_valid[j + 1] == sorted_valid[i]:
            j += 1
        if j - i + 1 > best_c:
            best_c, best_v = j - i + 1, sorted_valid[i]
        i = j + 1
    return np.float64(best_v)


def column_stats(arr, nodata, flags, p=0.95, percentiles=None):
    """One node's statistics over its realization values.

    Restates the body of the cell loop of ``GridFiles.stats`` (grid.py:680-701)
    and the layer loop of ``stats_area`` (grid.py:829-850).  Returns a dict.
    """
    a = replace_nodata(np.asarray(arr), nodata)
    valid = np.sort(a[~np.isnan(a)])
    n = valid.size
    out = {'n': n}
    with np.errstate(all='ignore'):
        mean = valid.sum() / n if n else np.nan
        if flags & FLAG_MEAN:
            out['mean'] = np.nanmean(a) if n else np.nan
        if flags & FLAG_MED:
            out['median'] = (np.nan if n == 0 else
                             (np.float64(valid[(n - 1) // 2])
                              + np.float64(valid[n // 2])) / 2)
        if flags & FLAG_SKEW:
            out['skewness'] = skew_g1(valid)
        var = (np.nan if n == 0 else
               ((valid - mean) ** 2).sum() / (n - 1) if n > 1 else np.nan)
        if flags & FLAG_VAR:
            out['variance'] = var
        std = np.sqrt(var)
        if flags & FLAG_STD:
            out['std'] = std
        if flags & FLAG_COEFVAR:
            out['coefvar'] = std / mean * 100
        if flags & FLAG_PERC:
            out['lperc'] = quantile_linear(valid, (1 - p) / 2)
            out['rperc'] = quantile_linear(valid, 1 - (1 - p) / 2)
        if percentiles is not None:
            out['percentiles'] = np.array(
                [quantile_linear(valid, q) for q in percentiles])
        if flags & FLAG_MODE:
            out['mode'] = mode_smallest(valid)
    return out


def column_stats_pandas(arr, nodata, flags, p=0.95):
    """The same, but through the very library calls the reference makes
    (np.nan* in place of bottleneck; ``pd.Series.skew`` / ``.quantile``).
    Used to cross-check :func:`column_stats` and as the literal CPU baseline."""
    a = replace_nodata(np.asarray(arr), nodata)
    out = {}
    with np.errstate(all='ignore'):
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            if flags & FLAG_MEAN:
                out['mean'] = np.nanmean(a)
            if flags & FLAG_MED:
                out['median'] = np.nanmedian(a)
            if flags & FLAG_SKEW:
                out['skewness'] = pd.Series(a).skew()
            if flags & FLAG_VAR:
                out['variance'] = np.nanvar(a, ddof=1)
            if flags & FLAG_STD:
                out['std'] = np.nanstd(a, ddof=1)
            if flags & FLAG_COEFVAR:
                out['coefvar'] = np.nanstd(a, ddof=1) / np.nanmean(a) * 100
            if flags & FLAG_PERC:
                qq = pd.Series(a).quantile([(1 - p) / 2, 1 - (1 - p) / 2])
                out['lperc'], out['rperc'] = qq.values
    return out


# ---------------------------------------------------------------------------
# vectorised grid statistics (axis 0 = realizations)
# ---------------------------------------------------------------------------


def _sorted_valid(stack, nodata):
    """Sort every column; NaN (nodata) last.  Returns (sorted f64, n)."""
    if stack.dtype == np.float32:
        mask = stack == np.float32(nodata)
    else:
        mask = stack == nodata
    a = stack.astype(np.float64)
    a[mask] = np.nan
    a.sort(axis=0)
    n = (~np.isnan(a)).sum(axis=0)
    return a, n


def _take(a, idx):
    return np.take_along_axis(a, idx[None, :], axis=0)[0]


def _quantile_cols(s, n, q):
    """Column-wise :func:`quantile_linear` on a sorted (NaN-last) stack."""
    h = (n - 1) * np.float64(q)
    nn = np.maximum(n, 1)
    lo = np.floor(h).astype(np.int64)
    lo = np.clip(lo, 0, nn - 1)
    hi = np.minimum(lo + 1, nn - 1)
    t = h - lo
    a = _take(s, lo)
    b = _take(s, hi)
    d = b - a
    with np.errstate(invalid='ignore'):
        res = np.where(t >= 0.5, b - d * (1 - t), a + d * t)
    last = _take(s, nn - 1)
    res = np.where(h >= n - 1, last, res)
    return np.where(n == 0, np.nan, res)


def grid_stats(stack, nodata, flags, p=0.95, percentiles=None):
    """Vectorised ``GridFiles.stats`` (grid.py:601-735) over an [R, N] stack.

    Returns a dict of float64 arrays keyed like the reference's maps
    ('meanmap', 'medianmap', 'skewmap', 'varmap', 'stdmap', 'coefvarmap',
    'percmap' [N, 2]) plus the extensions 'percentiles' [N, nq] and 'modemap'.
    Intended behaviour for SURVEY H6.1 (all N nodes are computed).
    """
    s, n = _sorted_valid(np.asarray(stack), nodata)
    R, N = s.shape
    out = {}
    # The moments are summed the way the library calls of the reference sum them
    # (np.nanmean / np.nanvar standing in for bottleneck, pandas' nanskew): over the
    # column in its ORIGINAL order, missing values set to 0, with NumPy's pairwise
    # summation -- a row of the transposed stack reduces with the very same inner
    # loop as the 1-D column.  At 1e-16 the order is visible only where the
    # reference's result is itself decided by rounding: the coefficient of
    # variation of a column whose mean cancels to ~0, and pandas' zeroing of the
    # second / third moment of a CONSTANT column (below).
    raw = np.asarray(stack)
    nd64 = np.float64(np.float32(nodata)) if raw.dtype == np.float32 else np.float64(nodata)
    vals = np.array(raw.T, dtype=np.float64, order='C')                             # [N, R]
    missT = (vals == nd64) | np.isnan(vals)
    vals[missT] = 0.0
    with np.errstate(all='ignore'):
        mean = vals.sum(axis=1) / n
        adjp = vals - mean[:, None]
        adjp[missT] = 0.0
        adjp2 = adjp ** 2
        m2 = adjp2.sum(axis=1)
        if flags & FLAG_MEAN:
            out['meanmap'] = mean
        if flags & FLAG_MED:
            nn = np.maximum(n, 1)
            med = (_take(s, (nn - 1) // 2) + _take(s, nn // 2)) / 2
            out['medianmap'] = np.where(n == 0, np.nan, med)
        if flags & FLAG_SKEW:
            # pd.Series(arr).skew() (grid.py:687) = pandas.core.nanops.nanskew on the
            # column in its ORIGINAL order, missing values set to 0: the sums are
            # NumPy's pairwise sums over the contiguous column (a row of the
            # transposed stack reduces with the very same inner loop).  The order
            # matters for one thing only: nanskew zeroes m2 / m3 below a tolerance of
            # (eps max|x|)^2 n, which is what makes the skewness of a CONSTANT column
            # exactly 0 -- a mean summed in another order misses the constant by a
            # few ulps more and lands on the other side of that tolerance (found by
            # fuzzing the oracle against the reference on float64 stacks; float32
            # stacks sum exactly in float64 and never got there).
            m2p = m2
            m3p = (adjp2 * adjp).sum(axis=1)
            eps = np.finfo(np.float64).eps
            max_abs = np.abs(vals).max(axis=1, initial=0.0)
            m2z = np.where(np.abs(m2p) < ((eps * max_abs) ** 2) * n, 0.0, m2p)
            m3z = np.where(np.abs(m3p) < ((eps * max_abs) ** 3) * n, 0.0, m3p)
            res = (n * (n - 1.0) ** 0.5 / (n - 2.0)) * (m3z / m2z ** 1.5)
            res = np.where(m2z == 0, 0.0, res)
            res = np.where(n < 3, np.nan, res)
            out['skewmap'] = res
        var = np.where(n > 1, m2 / (n - 1.0), np.nan)
        std = np.sqrt(var)
        if flags & FLAG_VAR:
            out['varmap'] = var
        if flags & FLAG_STD:
            out['stdmap'] = std
        if flags & FLAG_COEFVAR:
            out['coefvarmap'] = std / mean * 100
        if flags & FLAG_PERC:
            out['percmap'] = np.stack(
                [_quantile_cols(s, n, (1 - p) / 2),
                 _quantile_cols(s, n, 1 - (1 - p) / 2)], axis=1)
        if percentiles is not None:
            out['percentiles'] = (np.stack([_quantile_cols(s, n, q) for q in percentiles], axis=1)
                                  if len(percentiles) else np.empty((N, 0)))
        if flags & FLAG_MODE:
            idx = np.arange(R)[:, None]
            newrun = np.ones((R, N), dtype=bool)
            newrun[1:] = s[1:] != s[:-1]
            start = np.maximum.accumulate(np.where(newrun, idx, 0), axis=0)
            runlen = np.where(~np.isnan(s), idx - start + 1, 0)
            pos = runlen.argmax(axis=0)          # first longest run's end
            mode = _take(s, pos)
            out['modemap'] = np.where(n == 0, np.nan, mode)
    out['n'] = n
    return out


# ---------------------------------------------------------------------------
# station columns (stats_area) and detection
# ---------------------------------------------------------------------------

AREA_COLS = (('mean', FLAG_MEAN), ('median', FLAG_MED), ('skewness', FLAG_SKEW),
             ('variance', FLAG_VAR), ('std', FLAG_STD),
             ('coefvar', FLAG_COEFVAR))


def stats_area(stack, dims, first_coord, cells_size, nodata, loc, tol=0,
               flags=0, p=0.95, clip=True):
    """``GridFiles.stats_area`` (grid.py:737-918) on an in-memory [R, N] stack.

    Returns ``(name, DataFrame)``; columns ``x, y, z`` then, in the reference's
    order, ``mean, median, skewness, variance, std, coefvar, lperc, rperc``
    (grid.py:884-912).  ``x, y`` are GRID indices (grid.py:877-880).
    """
    stack = np.asarray(stack)
    node, lines = area_lines(loc, tol, dims, first_coord, cells_size, clip)
    dz = dims[2]
    cols = {'x': np.repeat(float(node[0]), dz),
            'y': np.repeat(float(node[1]), dz),
            'z': np.arange(first_coord[2], first_coord[2]
                           + cells_size[2] * dz)[:dz].astype(float)}
    names = [nm for nm, fl in AREA_COLS if flags & fl]
    if flags & FLAG_PERC:
        names += ['lperc', 'rperc']
    for nm in names:
        cols[nm] = np.zeros(dz)
    for layer in range(dz):
        # arr[i + j*K] = value of neighbour i in file j  (grid.py:823)
        arr = stack[:, lines[:, layer] - 1].reshape(-1)
        st = column_stats(arr, nodata, flags, p)
        for nm in names:
            cols[nm][layer] = st[nm]
    name = 'vertical line stats at (x,y) = ({0},{1})'.format(node[0], node[1])
    return name, pd.DataFrame(cols)


def fill_station(values, fillvals, time_min, time_max, time_step=1):
    """``fill_station`` (homog.py:279-337) on a DataFrame with columns
    x, y, time, station, clim [...]; includes repair R3."""
    varcol = list(values.columns).index('clim')
    timeserie = np.arange(time_min, time_max, time_step)
    filled = np.zeros((timeserie.shape[0], values.shape[1]))
    j = 0
    count = 0
    for i, itime in enumerate(timeserie):
        if j < len(values['time']) and itime == values['time'].iloc[j]:
            filled[i, :] = values.iloc[j, :]
            j += 1
        else:
            row = np.hstack([np.atleast_1d(v) for v in
                             [values.iloc[0, 0], values.iloc[0, 1], itime,
                              values.iloc[0, 3:varcol], fillvals[i]]])
            filled[i, :] = row[:values.shape[1]]
            count += 1
    return pd.DataFrame(filled, columns=values.columns), count


def detect(stack, dims, first_coord, cells_size, grid_nodata, obs, obs_nodata,
           rad=0, method='mean', prob=0.95, skewness=None, percentile=None,
           flag=True, optional_stats=None, clip=True):
    """``detect`` (homog.py:27-276) on an in-memory stack.

    ``obs`` is a DataFrame with columns x, y, time, station, clim (optionally
    Flag).  Returns ``(homogenised DataFrame, detected_number, missing_data)``.
    """
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
    sel['lmean'] = True
    sel['lperc'] = True
    keymap = dict(lmean=FLAG_MEAN, lmed=FLAG_MED, lskew=FLAG_SKEW,
                  lvar=FLAG_VAR, lstd=FLAG_STD, lcoefvar=FLAG_COEFVAR,
                  lperc=FLAG_PERC)
    flags = 0
    for k, v in sel.items():
        if v:
            flags |= keymap[k]

    obs = obs.copy()
    obs_xy = list(obs.loc[obs.first_valid_index(), ['x', 'y']])
    _, local = stats_area(stack, dims, first_coord, cells_size, grid_nodata,
                          obs_xy, rad, flags, prob, clip)
    if 'Flag' in obs.columns:
        obs = obs.drop('Flag', axis=1)
    nodatas = int(obs['clim'].isin([obs_nodata]).sum())
    obs = obs.replace(obs_nodata, np.nan)
    meanvalues = local['mean'].values
    fn = 0
    if obs.shape[0] < dims[2]:
        obs, fn = fill_station(obs, meanvalues, first_coord[2],
                               first_coord[2] + dims[2], cells_size[2])
    clim = obs['clim'].values.astype(float).copy()
    nanmask = np.isnan(clim)
    clim[nanmask] = meanvalues[nanmask]
    obs['clim'] = clim
    lperc, rperc = local['lperc'].values, local['rperc'].values
    with np.errstate(invalid='ignore'):
        hom_where = ~((clim >= lperc) & (clim <= rperc))
    detected = int(hom_where.sum())

    hom = obs.copy().dropna(axis=1)
    if method == 'mean':
        fix = local['mean'].values
    elif method == 'median':
        fix = local['median'].values
    elif method == 'skewness':
        fix = np.where(local['skewness'].values > skewness,
                       local['median'].values, local['mean'].values)
    else:
        if percentile != prob:
            _, vp = stats_area(stack, dims, first_coord, cells_size,
                               grid_nodata, obs_xy, rad, FLAG_PERC, percentile,
                               clip)
        else:
            vp = local
        with np.errstate(invalid='ignore'):
            fix = np.where(clim > rperc, vp['rperc'].values,
                           vp['lperc'].values)
    hom['clim'] = np.where(hom_where, fix, clim)
    if flag:
        hom['Flag'] = np.where(hom_where, clim, obs_nodata)
    names = dict(lmean='mean', lmed='median', lskew='skewness',
                 lvar='variance', lstd='std', lcoefvar='coefvar',
                 lperc='lperc')
    if optional_stats:
        for key, value in optional_stats.items():
            if value and key == 'lperc':
                if 'median' in local.columns:
                    med = local['median'].values
                else:
                    _, m = stats_area(stack, dims, first_coord, cells_size,
                                      grid_nodata, obs_xy, rad, FLAG_MED, 0.95,
                                      clip)
                    med = m['median'].values
                with np.errstate(invalid='ignore'):
                    hom['pdet'] = np.where(clim >= med, rperc, lperc)
            elif value:
                nm = names[key]
                hom[nm + '_new' if nm in hom.columns else nm] = \
                    local[nm].values
    return hom, detected, fn + nodatas
