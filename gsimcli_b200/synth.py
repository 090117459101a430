# -*- coding: utf-8 -*-
"""Seeded synthetic DSS realization stacks.

The reference ships no simulated maps (DSS is a closed binary), so benchmarks
and parity tests run on synthetic stacks.  The generator is a pure function of
``(seed, realization, node)`` built from 64-bit integer arithmetic only, so the
NumPy form below and the CUDA form (``gsb_synth_fill`` in
``csrc/gsb_synth.cu``) produce bit-identical values, any subset of columns can
be regenerated on the host for a parity check without ever materialising the
full stack, and every rank of a node-sharded run can fill its own slab.

Value model (all in units of 1/16 so that text ``%12.4f``, float64 and float32
agree exactly -- SURVEY H3, parity mode A):

    v16 = clamp(mu16(x, y, z) + noise16(seed, r, node), VMIN16, VMAX16)
    mu16    = 16*800 + 16*300 * tri(x/dx) * tri(y/dy) + 16*5*(z mod 100)
    noise16 = 16*600 * (u1*u2) + 16*100 * u3 - 16*200      (right-skewed)
    value   = v16 / 16,  or ``nodata`` with probability ~1e-3

The clamp creates atoms at the bounds (as DSS does, parsers/dss.py:322-325) and
the 1/16 lattice creates ties, so ``mode`` is non-trivial.
"""
import numpy as np

VMIN16 = 16 * 350
VMAX16 = 16 * 1400
NODATA_THRESH = 66          # of 65536 -> ~1.0e-3
DEFAULT_NODATA = -999.9

_M1 = np.uint64(0xBF58476D1CE4E5B9)
_M2 = np.uint64(0x94D049BB133111EB)
_GOLD = np.uint64(0x9E3779B97F4A7C15)


def _splitmix(z):
    z = (z + _GOLD)
    z = (z ^ (z >> np.uint64(30))) * _M1
    z = (z ^ (z >> np.uint64(27))) * _M2
    return z ^ (z >> np.uint64(31))


def _tri_centered(i, d):
    """Triangle wave over 0..d-1 mapped to -512..511."""
    a = (i * 2048) // d
    t = np.where(a < 1024, a, 2047 - a)
    return t - 512


def values16(seed, r, node, dims):
    """Integer lattice value (or -1 for nodata) for realization(s) ``r`` and
    0-based node(s) ``node`` (broadcastable int64 arrays)."""
    dx, dy, dz = (int(d) for d in dims)
    r = np.asarray(r, dtype=np.int64)
    node = np.asarray(node, dtype=np.int64)
    x = node % dx
    y = (node // dx) % dy
    z = node // (dx * dy)
    cx = _tri_centered(x, dx)
    cy = _tri_centered(y, dy)
    field = (((cx * cy + (1 << 18)) * 4800) >> 18) - 4800
    # the trend restarts every 100 layers, so that the 100-layer slabs of a
    # node-sharded multi-GPU run are statistically alike (weak scaling)
    mu = 12800 + field + 80 * (z % 100)
    with np.errstate(over='ignore'):
        key = (np.uint64(seed) * _M2
               + r.astype(np.uint64) * _GOLD) ^ \
              (node.astype(np.uint64) * _M1)
        h = _splitmix(_splitmix(key))
    u1 = (h & np.uint64(0xFFFF)).astype(np.int64)
    u2 = ((h >> np.uint64(16)) & np.uint64(0xFFFF)).astype(np.int64)
    u3 = ((h >> np.uint64(32)) & np.uint64(0xFFFF)).astype(np.int64)
    sel = ((h >> np.uint64(48)) & np.uint64(0xFFFF)).astype(np.int64)
    noise = ((u1 * u2 * 9600) >> 32) + ((u3 * 1600) >> 16) - 3200
    v = np.clip(mu + noise, VMIN16, VMAX16)
    return np.where(sel < NODATA_THRESH, -1, v)


def columns(seed, R, dims, nodes, nodata=DEFAULT_NODATA):
    """float32 [R, len(nodes)] -- the realization columns of the given nodes."""
    nodes = np.asarray(nodes, dtype=np.int64)
    r = np.arange(R, dtype=np.int64)[:, None]
    v = values16(seed, r, nodes[None, :], dims)
    out = (v.astype(np.float32) / np.float32(16))
    out[v < 0] = np.float32(nodata)
    return out


def stack(seed, R, dims, node0=0, node1=None, nodata=DEFAULT_NODATA):
    """float32 [R, node1-node0] slab of the stack (x fastest, then y, z)."""
    n = int(np.prod(dims))
    node1 = n if node1 is None else node1
    return columns(seed, R, dims, np.arange(node0, node1), nodata)


def to_text(values, fmt='%12.4f', eol='\r\n', header=None):
    """GSLIB ASCII form of ONE realization (1-D array): optional 3 header
    lines then one number per line.  ``'%12.4f' + CRLF`` is the 14 bytes/node
    form the reference's disk estimator assumes (interface/gui.py:1414)."""
    body = eol.join(fmt % float(v) for v in values) + eol
    if header:
        body = eol.join(header) + eol + body
    return body.encode('ascii')


def write_stack(dirpath, base, st, fmt='%12.4f', eol='\r\n', header=None):
    """Write an [R, N] stack as ``base.out, base_2.out, ...`` (the naming
    GridFiles.load expects, tools/utils.py:104-123).  Returns the first path."""
    import os
    first = os.path.join(dirpath, base + '.out')
    for r in range(st.shape[0]):
        path = first if r == 0 else os.path.join(
            dirpath, '{0}_{1}.out'.format(base, r + 1))
        with open(path, 'wb') as fh:
            fh.write(to_text(st[r], fmt, eol, header))
    return first


def stations(seed, S, R, dims, first_coord, cells_size, tol=0,
             nodata=DEFAULT_NODATA, break_frac=0.3, nodata_frac=0.05,
             missing_frac=0.05):
    """S candidate stations at interior nodes (>= tol+1 from every edge).

    Each station's observed series is one extra realization (index R) at its
    node, with +-300 step breaks over random sub-intervals, some nodata rows
    and some missing years.  Returns a list of dicts with keys ``xy`` (world
    coords, jittered inside the cell), ``node`` (1-based ix, iy) and ``obs``
    (float64 array [rows, 5]: x, y, time, station, clim)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    dx, dy, dz = dims
    out = []
    m = tol + 1
    for s in range(S):
        ix = int(rng.integers(1 + m, dx - m + 1))
        iy = int(rng.integers(1 + m, dy - m + 1))
        wx = first_coord[0] + (ix - 1 + rng.uniform(-0.3, 0.3)) * cells_size[0]
        wy = first_coord[1] + (iy - 1 + rng.uniform(-0.3, 0.3)) * cells_size[1]
        nodes = (ix - 1) + dx * (iy - 1) + dx * dy * np.arange(dz)
        v = values16(seed, np.int64(R), nodes, dims).astype(np.float64) / 16
        v[v < 0] = 800.0
        if rng.uniform() < break_frac or s == 0:
            a = int(rng.integers(0, dz))
            b = int(rng.integers(a, dz)) + 1
            v[a:b] += rng.choice([-300.0, 300.0])
        clim = v.copy()
        clim[rng.uniform(size=dz) < nodata_frac] = nodata
        keep = rng.uniform(size=dz) >= missing_frac
        keep[0] = True
        t = first_coord[2] + cells_size[2] * np.arange(dz)
        obs = np.column_stack([np.full(dz, wx), np.full(dz, wy), t,
                               np.full(dz, float(s + 1)), clim])[keep]
        out.append(dict(xy=(wx, wy), node=(ix, iy), obs=obs))
    return out
