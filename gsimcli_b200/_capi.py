# -*- coding: utf-8 -*-
"""ctypes binding of ``libgsimcli_b200.so`` (C ABI in ``include/gsimcli_b200.h``).

The library is built in-tree by ``gsimcli_b200/build.py``.  There is no CPU
fallback anywhere in this package: if the shared library is missing, or a call
needs a CUDA device and none is present, an exception is raised.

Status codes map onto the reference's exception classes:
``GSB_ENOENT`` -> ``IOError`` (grid.py:506,574), ``GSB_EINVAL`` / ``GSB_EPARSE``
-> ``ValueError`` (homog.py:153; ``float()`` on a malformed line, grid.py:823),
``GSB_ENOMEM`` -> ``MemoryError``, ``GSB_ECUDA`` / ``GSB_ENCCL`` -> ``RuntimeError``.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIBPATH = os.path.join(HERE, 'libgsimcli_b200.so')

GSB_OK, GSB_ENOENT, GSB_EINVAL, GSB_EPARSE, GSB_ECUDA, GSB_ENCCL, GSB_ENOMEM = \
    range(7)
MEAN, MEDIAN, SKEW, VAR, STD, COEFVAR, PERC, MODE = (1, 2, 4, 8, 16, 32, 64,
                                                     128)
METHODS = {'mean': 0, 'median': 1, 'skewness': 2, 'percentile': 3}
MAX_PLANES = 64
MAX_R_GRID = 1024

_lib = None

c_i64 = C.c_int64
c_vp = C.c_void_p
p_f32 = C.POINTER(C.c_float)
p_f64 = C.POINTER(C.c_double)
p_i64 = C.POINTER(C.c_int64)
p_i32 = C.POINTER(C.c_int32)

# name -> (restype, argtypes); must list every symbol the header declares
SIGNATURES = {
    'gsb_version': (C.c_int, []),
    'gsb_last_error': (C.c_int, [C.c_char_p, C.c_size_t]),
    'gsb_device_count': (C.c_int, [C.POINTER(C.c_int)]),
    'gsb_device_info': (C.c_int, [C.c_int, C.POINTER(C.c_int),
                                  C.POINTER(C.c_int), C.POINTER(C.c_int),
                                  C.POINTER(C.c_size_t), C.c_char_p,
                                  C.c_size_t]),
    'gsb_stack_create': (C.c_int, [C.c_int, c_i64, c_i64, C.POINTER(c_vp)]),
    'gsb_stack_destroy': (C.c_int, [c_vp]),
    'gsb_stack_info': (C.c_int, [c_vp, p_i64, p_i64, p_i64,
                                 C.POINTER(C.c_int)]),
    'gsb_stack_set_option': (C.c_int, [c_vp, C.c_char_p, c_i64]),
    'gsb_stack_device_ptr': (C.c_int, [c_vp, C.POINTER(c_vp)]),
    'gsb_stack_upload_f32': (C.c_int, [c_vp, c_i64, c_i64, c_i64, c_i64, c_vp,
                                       c_i64, c_vp]),
    'gsb_stack_upload_f32_async': (C.c_int, [c_vp, c_i64, c_i64, c_i64, c_i64,
                                             c_vp, c_i64, c_vp]),
    'gsb_stack_download_f32': (C.c_int, [c_vp, c_i64, c_i64, c_i64, c_i64,
                                         c_vp, c_i64, c_vp]),
    'gsb_decode_text': (C.c_int, [c_vp, c_i64, c_vp, C.c_size_t, C.c_int,
                                  C.c_int, c_i64, c_i64, c_i64, p_i64, c_vp]),
    'gsb_decode_text_device': (C.c_int, [c_vp, c_i64, c_vp, C.c_size_t,
                                         C.c_int, C.c_int, c_i64, c_i64, c_i64,
                                         c_vp, c_vp]),
    'gsb_encode_text_device': (C.c_int, [c_vp, c_i64, c_i64, c_i64, C.c_int,
                                         C.c_int, C.c_int, c_vp, c_vp]),
    'gsb_synth_fill': (C.c_int, [c_vp, C.c_uint64, c_i64, c_i64, c_i64, c_i64,
                                 C.c_float, c_vp]),
    'gsb_stats_grid': (C.c_int, [c_vp, C.c_uint32, C.c_double, p_f64, C.c_int,
                                 C.c_float, c_i64, c_i64, c_vp, c_i64, C.c_int,
                                 p_f32, c_vp]),
    'gsb_stats_grid_planes': (C.c_int, [C.c_uint32, C.c_int]),
    'gsb_stats_columns': (C.c_int, [c_vp, p_i64, C.c_int, C.c_int, C.c_int,
                                    C.c_uint32, C.c_double, p_f64, C.c_int,
                                    C.c_float, c_vp, C.c_int, c_vp]),
    'gsb_stats_columns_device': (C.c_int, [c_vp, c_vp, C.c_int, C.c_int, C.c_int,
                                          C.c_uint32, C.c_double, p_f64, C.c_int,
                                          C.c_float, c_vp, c_vp]),
    'gsb_detect_table': (C.c_int, [c_vp, c_vp, c_vp, c_vp, c_i64] + [C.c_int] * 11
                         + [C.c_double, C.c_double, c_vp, c_vp, C.c_int, c_vp]),
    'gsb_extract_columns': (C.c_int, [c_vp, p_i64, C.c_int, C.c_int, C.c_int,
                                      c_vp, c_vp]),
    'gsb_detect': (C.c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, C.c_int,
                             C.c_int, C.c_int, C.c_double, C.c_double, c_vp,
                             c_vp, c_vp, c_vp, C.c_int, C.c_int, c_vp]),
    'gsb_debug_cs_counters': (C.c_int, [c_vp, C.c_int]),
    'gsb_debug_k2_flagged': (C.c_int, [c_vp, C.POINTER(C.c_int)]),
}


def lib():
    """Load (once) and return the shared library; raise if it is not built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIBPATH):
            raise RuntimeError(
                'libgsimcli_b200.so is not built (run `python -m '
                'gsimcli_b200.build`); this package has no CPU fallback')
        handle = C.CDLL(LIBPATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def last_error():
    buf = C.create_string_buffer(512)
    lib().gsb_last_error(buf, 512)
    return buf.value.decode('utf-8', 'replace')


def check(rc):
    """Raise the reference's exception class for a non-zero status."""
    if rc == GSB_OK:
        return
    msg = last_error()
    if rc == GSB_ENOENT:
        raise IOError(msg)
    if rc in (GSB_EINVAL, GSB_EPARSE):
        raise ValueError(msg)
    if rc == GSB_ENOMEM:
        raise MemoryError(msg)
    raise RuntimeError('gsimcli_b200 (status {0}): {1}'.format(rc, msg))


def device_count():
    n = C.c_int(0)
    rc = lib().gsb_device_count(C.byref(n))
    if rc != GSB_OK:
        return 0
    return n.value


def require_device():
    if device_count() < 1:
        raise RuntimeError('gsimcli_b200 needs a CUDA device (sm_100a); none '
                           'is visible and there is no CPU fallback')


def device_info(device=0):
    sm, mj, mn = C.c_int(0), C.c_int(0), C.c_int(0)
    mem = C.c_size_t(0)
    name = C.create_string_buffer(256)
    check(lib().gsb_device_info(device, C.byref(sm), C.byref(mj), C.byref(mn),
                                C.byref(mem), name, 256))
    return dict(sm_count=sm.value, cc=(mj.value, mn.value), total_mem=mem.value,
                name=name.value.decode())


def flags_from(lmean=False, lmed=False, lskew=False, lvar=False, lstd=False,
               lcoefvar=False, lperc=False, lmode=False):
    return ((MEAN if lmean else 0) | (MEDIAN if lmed else 0)
            | (SKEW if lskew else 0) | (VAR if lvar else 0)
            | (STD if lstd else 0) | (COEFVAR if lcoefvar else 0)
            | (PERC if lperc else 0) | (MODE if lmode else 0))


def _ptr(arr):
    return C.c_void_p(arr.ctypes.data)


class Stack(object):
    """Device-resident ``[R][N]`` float32 realization stack (``gsb_stack``)."""

    def __init__(self, R, N, device=0):
        require_device()
        self.handle = c_vp()
        self.R, self.N, self.device = int(R), int(N), int(device)
        check(lib().gsb_stack_create(self.device, self.R, self.N,
                                     C.byref(self.handle)))
        pitch = c_i64(0)
        check(lib().gsb_stack_info(self.handle, None, None, C.byref(pitch),
                                   None))
        self.pitch = pitch.value

    def close(self):
        if getattr(self, 'handle', None) and _lib is not None:
            _lib.gsb_stack_destroy(self.handle)
        self.handle = None

    __del__ = close

    def set_option(self, key, value):
        check(lib().gsb_stack_set_option(self.handle, key.encode(), int(value)))

    def k2_flagged(self):
        """Tiles of the last stats_grid call redone by the counting-sort kernel."""
        n = C.c_int(0)
        check(lib().gsb_debug_k2_flagged(self.handle, C.byref(n)))
        return n.value

    def device_ptr(self):
        p = c_vp()
        check(lib().gsb_stack_device_ptr(self.handle, C.byref(p)))
        return p.value

    def upload(self, host, r0=0, n0=0, stream=None):
        host = np.ascontiguousarray(host, dtype=np.float32)
        if host.ndim == 1:
            host = host[None, :]
        nr, nn = host.shape
        check(lib().gsb_stack_upload_f32(self.handle, r0, nr, n0, nn,
                                         _ptr(host), host.shape[1], stream))

    def upload_ptr(self, ptr, nr, nn, pitch, r0=0, n0=0, stream=None):
        check(lib().gsb_stack_upload_f32(self.handle, r0, nr, n0, nn,
                                         c_vp(ptr), pitch, stream))

    def download(self, r0=0, nr=None, n0=0, nn=None, stream=None):
        nr = self.R - r0 if nr is None else nr
        nn = self.N - n0 if nn is None else nn
        out = np.empty((nr, nn), dtype=np.float32)
        check(lib().gsb_stack_download_f32(self.handle, r0, nr, n0, nn,
                                           _ptr(out), nn, stream))
        return out

    def synth_fill(self, seed, dims, node0=0, nodata=-999.9, stream=None):
        check(lib().gsb_synth_fill(self.handle, seed, dims[0], dims[1],
                                   dims[2], node0, nodata, stream))

    def decode_text(self, r, text, header_lines=0, line_width=0, first_line=0,
                    n0=0, nlines=None, stream=None):
        """K1: decode one realization's GSLIB text (bytes / buffer) into row r."""
        nlines = self.N - n0 if nlines is None else nlines
        buf = np.frombuffer(text, dtype=np.uint8)
        bad = c_i64(-1)
        rc = lib().gsb_decode_text(self.handle, r, _ptr(buf), buf.size,
                                   header_lines, line_width, first_line, n0,
                                   nlines, C.byref(bad), stream)
        check(rc)

    def stats_grid(self, flags, p=0.95, percentiles=None, nodata=-999.9, n0=0,
                   nn=None, out=None, stream=None, want_ms=False):
        """K2: full-grid statistics -> float32 ``[planes][nn]`` host array."""
        nn = self.N - n0 if nn is None else nn
        q = np.ascontiguousarray(percentiles if percentiles is not None else [],
                                 dtype=np.float64)
        nplanes = lib().gsb_stats_grid_planes(flags, q.size)
        if nplanes < 0:
            raise ValueError('too many output planes')
        if out is None:
            out = np.empty((nplanes, nn), dtype=np.float32)
        ms = C.c_float(0)
        check(lib().gsb_stats_grid(
            self.handle, flags, p, q.ctypes.data_as(p_f64), q.size, nodata, n0,
            nn, _ptr(out), out.shape[1], 0, C.byref(ms) if want_ms else None,
            stream))
        return (out, ms.value) if want_ms else out

    def stats_grid_device(self, flags, p, percentiles, nodata, n0, nn, out_ptr,
                          out_pitch, stream=None):
        """K2 with a DEVICE output pointer (enqueue only)."""
        q = np.ascontiguousarray(percentiles if percentiles is not None else [],
                                 dtype=np.float64)
        check(lib().gsb_stats_grid(
            self.handle, flags, p, q.ctypes.data_as(p_f64), q.size, nodata, n0,
            nn, c_vp(out_ptr), out_pitch, 1, None, stream))

    def stats_grid_from_host(self, host_ptr, host_pitch, flags, p, percentiles,
                             nodata, out, nslabs=8):
        """End-to-end K2 from a HOST stack: the ``[R][host_pitch]`` float32
        matrix at ``host_ptr`` (pinned memory for full PCIe speed) is uploaded
        in ``nslabs`` node slabs; the reduction of slab k and the download of
        its maps run while slab k+1 is still on the bus, so the call costs the
        upload plus one slab of compute.  ``out`` is a pinned float32
        ``[planes][N]`` torch tensor.  PyTorch supplies the streams/events only."""
        import torch
        dev = torch.device('cuda', self.device)
        q = percentiles if percentiles is not None else []
        nplanes = lib().gsb_stats_grid_planes(flags, len(q))
        if getattr(self, '_e2e_maps', None) is None or \
                self._e2e_maps.shape != (nplanes, self.N):
            self._e2e_maps = torch.empty((nplanes, self.N), dtype=torch.float32,
                                         device=dev)
            self._e2e_up = torch.cuda.Stream(device=dev)
            self._e2e_run = torch.cuda.Stream(device=dev)
        up, run, d_maps = self._e2e_up, self._e2e_run, self._e2e_maps
        cur = torch.cuda.current_stream(dev)
        up.wait_stream(cur)
        run.wait_stream(cur)
        step = -(-self.N // nslabs)
        step = (step + 31) // 32 * 32             # whole 32-node tiles per slab
        for a in range(0, self.N, step):
            b = min(self.N, a + step)
            check(lib().gsb_stack_upload_f32_async(
                self.handle, 0, self.R, a, b - a, c_vp(host_ptr + 4 * a),
                host_pitch, c_vp(up.cuda_stream)))
            ev = torch.cuda.Event()
            ev.record(up)
            run.wait_event(ev)
            self.stats_grid_device(flags, p, percentiles, nodata, a, b - a,
                                   d_maps.data_ptr() + 4 * a, self.N,
                                   run.cuda_stream)
            with torch.cuda.stream(run):
                # one contiguous copy per plane (a strided 2-D tensor copy to
                # pinned memory goes through a kernel writing over PCIe)
                for k in range(nplanes):
                    out[k, a:b].copy_(d_maps[k, a:b], non_blocking=True)
        cur.wait_stream(run)
        cur.wait_stream(up)
        return out

    def stats_columns(self, node_ids, flags, p=0.95, percentiles=None,
                      nodata=-999.9, stream=None):
        """K2c: node_ids int64 ``[S][dz][K]`` (local, -1 absent) -> float64
        ``[S][dz][C]`` table."""
        ids = np.ascontiguousarray(node_ids, dtype=np.int64)
        S, dz, K = ids.shape
        q = np.ascontiguousarray(percentiles if percentiles is not None else [],
                                 dtype=np.float64)
        ncols = lib().gsb_stats_grid_planes(flags, q.size)
        table = np.empty((S, dz, ncols), dtype=np.float64)
        check(lib().gsb_stats_columns(
            self.handle, ids.ctypes.data_as(p_i64), S, dz, K, flags, p,
            q.ctypes.data_as(p_f64), q.size, nodata, _ptr(table), 0, stream))
        return table

    def extract_columns(self, node_ids, stream=None):
        ids = np.ascontiguousarray(node_ids, dtype=np.int64)
        S, dz, K = ids.shape
        vals = np.empty((S, dz, K, self.R), dtype=np.float32)
        check(lib().gsb_extract_columns(self.handle, ids.ctypes.data_as(p_i64),
                                        S, dz, K, _ptr(vals), stream))
        return vals


def encode_values(values, width, prec, crlf=False, left=False, device=0):
    """Fixed-width text records of a float32 vector, formatted on the device
    (``gsb_encode_text_device``): uint8 array ``[n][width + eol]``."""
    import torch
    vals = np.ascontiguousarray(values, dtype=np.float32).reshape(1, -1)
    n = vals.shape[1]
    rec = width + (2 if crlf else 1)
    stack = Stack(1, n, device)
    try:
        stack.upload(vals)
        dtext = torch.empty(n * rec, dtype=torch.uint8,
                            device=torch.device('cuda', device))
        check(lib().gsb_encode_text_device(
            stack.handle, 0, 0, n, width, prec, (1 if crlf else 0) | (2 if left else 0),
            dtext.data_ptr(), None))
        out = dtext.cpu().numpy().reshape(n, rec)
    finally:
        stack.close()
    return out


def detect(clim, mean, lperc, rperc, fix_a, fix_b=None, fix_c=None,
           method='mean', thr=0.0, nodata=-999.9, device=0, stream=None):
    """K3 on host arrays ``[S][dz]`` (float64).  Returns
    ``(filled, homog, flag, detected[S])``."""
    require_device()
    arrs = [np.ascontiguousarray(a, dtype=np.float64)
            for a in (clim, mean, lperc, rperc, fix_a)]
    S, dz = arrs[0].shape
    fb = None if fix_b is None else np.ascontiguousarray(fix_b, np.float64)
    fc = None if fix_c is None else np.ascontiguousarray(fix_c, np.float64)
    # the library reads S*dz doubles from every array: a shorter one would be
    # over-read.  The reference fails the same way, in pandas
    # (`meanvalues.index = obs.values.index`, homog.py:197): ValueError.
    for a in arrs[1:] + [x for x in (fb, fc) if x is not None]:
        if a.shape != (S, dz):
            raise ValueError(
                'Length mismatch: Expected axis has {0} elements, new values '
                'have {1} elements'.format(a.shape[-1] if a.ndim else 0, dz))
    filled = np.empty((S, dz))
    homog = np.empty((S, dz))
    flag = np.empty((S, dz))
    det = np.zeros(S, dtype=np.int32)
    check(lib().gsb_detect(
        _ptr(arrs[0]), _ptr(arrs[1]), _ptr(arrs[2]), _ptr(arrs[3]),
        _ptr(arrs[4]), None if fb is None else _ptr(fb),
        None if fc is None else _ptr(fc), S, dz, METHODS[method], thr, nodata,
        _ptr(filled), _ptr(homog), _ptr(flag), _ptr(det), 0, device, stream))
    return filled, homog, flag, det
