# -*- coding: utf-8 -*-
"""Benchmark of the local-PDF + detection hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Workload (BASELINE.json configs[2], the configuration the metric's target is
quoted on; it fits one B200): 1000 realizations of a 200x200x100 grid, full
percentile set 5..95 step 5 + mode + moments + median, plus detection with
median correction at 20 candidate stations.  One "step" = one pass of the hot
path over the whole stack: K2 (full-grid reduction), K2c (station columns) and
K3 (detect/correct).  With N > 1 GPUs every rank holds its own block of 100
time slices of a 200x200x(100 N) grid (weak scaling, node sharding; the only
exchange is the all-gather of the per-station tables).

Keys of the JSON line: see the task contract.  ``value`` = node-realizations/s
with the stack resident in HBM; ``e2e`` = the same through the public API with
the float32 stack in pinned HOST memory (H2D of the stack in node slabs, each
slab's reduction and the D2H of its maps overlapped with the next upload -- all
inside the timed region); ``roofline`` for the dominant kernel (K2) against
the measured HBM peak; ``cpu_baseline`` = the NumPy oracle on a bounded sample
on the host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

R = int(os.environ.get('GSB_BENCH_R', 1000))
DIMS1 = [int(x) for x in os.environ.get('GSB_BENCH_DIMS', '200,200,100').split(',')]
FIRST = [0.0, 0.0, 1900]
CELLS = [1000.0, 1000.0, 1]
NSTATIONS = 20
SEED = 1003
ND = -999.9
QS = [round(0.05 * k, 2) for k in range(1, 20)]
METRIC = 'node-realizations/sec for local-PDF stats + detection'
UNIT = 'node-realizations/s'


def workload_name(n_gpus):
    d = DIMS1
    return ('{0} realizations of {1}x{2}x{3} grid per GPU ({4} GPU(s), '
            'node-sharded by time slices), mean/var/std/coefvar/skew/median + '
            'percentiles 5..95 step 5 + mode, detection + median correction at '
            '{5} stations').format(R, d[0], d[1], d[2], n_gpus, NSTATIONS)


def peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        with open(path) as fh:
            return float(json.load(fh)['hbm_gbs']), 'measured (MEASURED_PEAKS.json)'
    return 6650.0, 'fallback (B200_PROFILING.md)'


def measured_traffic(r, nn):
    """DRAM bytes per launch of the K2 kernel (dram__bytes_read.sum +
    dram__bytes_write.sum) from the committed ``ncu --set full`` capture of
    this very workload (profiles/k2_traffic.json), else None."""
    path = os.path.join(ROOT, 'profiles', 'k2_traffic.json')
    try:
        with open(path) as fh:
            for rec in json.load(fh):
                if rec['R'] == r and rec['nodes'] == nn:
                    return rec['dram_bytes_per_launch']
    except (OSError, ValueError, KeyError):
        pass
    return None


class ClockSampler(threading.Thread):
    """Samples nvidia-smi clocks / throttle reasons during the timed region."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.reasons = set()
        self.stop_flag = threading.Event()
        self.max_mhz = None

    def run(self):
        q = ('clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,'
             'clocks_event_reasons.hw_thermal_slowdown,'
             'clocks_event_reasons.sw_thermal_slowdown,'
             'clocks_event_reasons.sw_power_cap')
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown',
                 'sw_power_cap']
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(
                    ['nvidia-smi', '-i', str(self.index), '--query-gpu=' + q,
                     '--format=csv,noheader,nounits'], capture_output=True,
                    text=True, timeout=5).stdout.strip().split(',')
                self.samples.append(float(out[0]))
                self.max_mhz = float(out[1])
                for nm, v in zip(names, out[2:]):
                    if 'Active' in v and 'Not' not in v:
                        self.reasons.add(nm)
            except Exception:
                pass
            self.stop_flag.wait(0.2)

    def summary(self):
        if not self.samples:
            return {'sm_mhz': None, 'sm_max_mhz': self.max_mhz, 'reasons': []}
        return {'sm_mhz': float(np.median(self.samples)),
                'sm_max_mhz': self.max_mhz, 'reasons': sorted(self.reasons)}


# ---------------------------------------------------------------- CPU legs
def _oracle_chunk(args):
    seed, dims, nodes = args
    from gsimcli_b200 import synth
    from oracle import np_oracle as O
    cols = synth.columns(seed, R, dims, nodes, ND)
    flags = (O.FLAG_MEAN | O.FLAG_MED | O.FLAG_SKEW | O.FLAG_VAR | O.FLAG_STD
             | O.FLAG_COEFVAR | O.FLAG_MODE)
    t0 = time.perf_counter()
    res = O.grid_stats(cols, ND, flags, 0.95, QS)
    return time.perf_counter() - t0, float(np.nansum(res['medianmap']))


def cpu_sample(dims, nnodes, procs, chunk=8192):
    """Time the oracle (``oracle/np_oracle.py``, the NumPy port of the
    reference's per-node statistics) on ``nnodes`` columns of the workload,
    in chunks spread over ``procs`` processes.  The realization values are
    generated in memory outside the timed region (the reference additionally
    parses text).  Returns (node-realizations/s, seconds of oracle time)."""
    import multiprocessing as mp
    rng = np.random.default_rng(0)
    nodes = np.sort(rng.choice(int(np.prod(dims)), size=nnodes, replace=False))
    nchunks = max(max(procs, 1) * 2, (nnodes + chunk - 1) // chunk)
    chunks = [c for c in np.array_split(nodes, nchunks) if len(c)]
    if procs > 1:
        with mp.get_context('fork').Pool(procs) as pool:
            times = pool.map(_oracle_chunk, [(SEED, dims, c) for c in chunks])
        # oracle compute time scheduled over `procs` workers
        busy = sum(t for t, _ in times)
        wall = max(busy / procs, max(t for t, _ in times))
    else:
        times = [_oracle_chunk((SEED, dims, c)) for c in chunks]
        wall = sum(t for t, _ in times)
    return nnodes * R / wall, wall


def run_reference(args):
    rank = int(os.environ.get('RANK', 0))
    if rank != 0:
        return
    cores = len(os.sched_getaffinity(0))
    dims = list(DIMS1)
    nnodes = int(os.environ.get('GSB_REF_NODES', 65536 * max(1, min(cores, 32))))
    vals, walls = [], []
    for i in range(args.warmup + args.steps):
        v, w = cpu_sample(dims, nnodes, cores)
        if i >= args.warmup:
            vals.append(v)
            walls.append(w)
    value = float(np.mean(vals))
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT,
        'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': float(np.mean(walls)) * 1e3, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64',
        'data': 'synthetic',
        'config': {'workload': workload_name(args.gpus)},
        'cpu_baseline': {
            'value': value, 'unit': UNIT, 'cores': cores, 'kind': 'port',
            'sample': ('{0} node columns x {1} realizations per step, NumPy '
                       'oracle (oracle/np_oracle.py) over {2} processes; the '
                       'reference itself is Python 2 and cannot run on this '
                       'box').format(nnodes, R, cores)},
        'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0,
                'd2h_bytes_per_step': 0},
    }
    print(json.dumps(line))


# ------------------------------------------------------------------ GPU arm
def run_gpu(args):
    import pandas as pd
    import torch
    import torch.distributed as td
    from gsimcli_b200 import _capi, synth, dist
    from gsimcli_b200.tools import grid as gr, homog as hmg

    world = int(os.environ.get('WORLD_SIZE', 1))
    rank = int(os.environ.get('RANK', 0))
    local = int(os.environ.get('LOCAL_RANK', 0))
    if not torch.cuda.is_available():
        raise RuntimeError('bench.py needs a CUDA device; there is no CPU path')
    torch.cuda.set_device(local)
    if world > 1:
        td.init_process_group('nccl', device_id=torch.device('cuda', local))

    dims = [DIMS1[0], DIMS1[1], DIMS1[2] * world]
    n0, n1 = dist.slab_range(dims, rank, world)
    nn = n1 - n0
    grids = gr.GridFiles.from_synthetic(SEED, R, dims, FIRST, CELLS, ND,
                                        device=local,
                                        node_range=(n0, n1) if world > 1 else None)
    stack = grids._stack
    flags = (_capi.MEAN | _capi.MEDIAN | _capi.SKEW | _capi.VAR | _capi.STD
             | _capi.COEFVAR | _capi.MODE)
    nplanes = _capi.lib().gsb_stats_grid_planes(flags, len(QS))
    out_pitch = (nn + 3) // 4 * 4
    d_out = torch.empty((nplanes, out_pitch), dtype=torch.float32, device='cuda')
    stations = synth.stations(SEED, NSTATIONS, R, dims, FIRST, CELLS)
    cols = ['x', 'y', 'time', 'station', 'clim']

    def obs_list():
        return [gr.PointSet('cand%d' % i, ND, 5, list(cols),
                            np.array(s['obs'], dtype=np.float64))
                for i, s in enumerate(stations)]

    # K2 runs on its own stream and leaves a few SMs free; the station path
    # (K2c + K3, tiny) runs on a second stream concurrently with it, so its
    # host-side pandas glue overlaps the full-grid kernel
    stack.set_option('k2_sm_reserve', int(os.environ.get('GSB_K2_RESERVE', 4)))
    s_grid = torch.cuda.Stream()
    s_stat = torch.cuda.Stream()
    ev_k0 = torch.cuda.Event(enable_timing=True)
    ev_k1 = torch.cuda.Event(enable_timing=True)
    kernel_ms = []

    def step_resident(timed):
        """Hot path with the stack resident in HBM."""
        cur = torch.cuda.current_stream()
        s_grid.wait_stream(cur)
        s_stat.wait_stream(cur)
        if timed:
            ev_k0.record(s_grid)
        stack.stats_grid_device(flags, 0.95, QS, ND, 0, nn, d_out.data_ptr(),
                                out_pitch, s_grid.cuda_stream)
        if timed:
            ev_k1.record(s_grid)
        res = hmg.detect_many(grids, obs_list(), method='median', prob=0.95,
                              stream=s_stat.cuda_stream)
        cur.wait_stream(s_grid)
        cur.wait_stream(s_stat)
        if timed:
            ev_k1.synchronize()
            kernel_ms.append(ev_k0.elapsed_time(ev_k1))
        return res

    def barrier():
        if world > 1:
            td.barrier()
        torch.cuda.synchronize()

    def timed_loop(fn, steps):
        barrier()
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            out = fn()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device='cuda')
            td.all_reduce(t, op=td.ReduceOp.MAX)
            ms = float(t.item())
        return ms, out

    sampler = ClockSampler(local)
    for _ in range(args.warmup):
        step_resident(False)
    sampler.start()
    ms_total, res = timed_loop(lambda: step_resident(True), args.steps)
    units_per_step = float(R) * nn * world
    value = units_per_step * args.steps / (ms_total * 1e-3)
    detected = int(sum(r[1] for r in res))

    # ---- e2e: the float32 stack in pinned host memory, maps back to host
    e2e = None
    fast = bool(os.environ.get('GSB_BENCH_FAST'))   # developer runs: resident leg only
    try:
        if fast:
            raise RuntimeError('skipped (GSB_BENCH_FAST)')
        pitch = stack.pitch
        host = torch.empty((R, pitch), dtype=torch.float32, pin_memory=True)
        dview = torch.empty(0)
        import ctypes
        # one-off D2H of the resident synthetic stack (setup, not timed)
        _capi.check(_capi.lib().gsb_stack_download_f32(
            stack.handle, 0, R, 0, nn, ctypes.c_void_p(host.data_ptr()), pitch,
            None))
        del dview
        maps = torch.empty((nplanes, nn), dtype=torch.float32, pin_memory=True)
        maps_np = maps.numpy()

        stream = torch.cuda.current_stream().cuda_stream

        def step_e2e():
            # upload, reduction and map download pipelined over node slabs
            stack.stats_grid_from_host(host.data_ptr(), pitch, flags, 0.95, QS,
                                       ND, maps)
            res = hmg.detect_many(grids, obs_list(), method='median',
                                  prob=0.95, stream=stream)
            torch.cuda.current_stream().synchronize()      # maps are on the host
            return res

        e2e_steps = max(1, min(args.steps, 3))
        step_e2e()
        ms_e2e, res2 = timed_loop(step_e2e, e2e_steps)
        assert int(sum(r[1] for r in res2)) == detected
        # the pipelined host path must reproduce the resident maps bit for bit
        same = (maps.to('cuda').view(torch.int32) == d_out[:, :nn].view(torch.int32))
        assert bool(same.all()), 'e2e maps differ from the resident-stack maps'
        del same
        e2e = {'value': units_per_step * e2e_steps / (ms_e2e * 1e-3),
               'unit': UNIT, 'h2d_bytes_per_step': int(R * nn * 4 * world),
               'd2h_bytes_per_step': int(nplanes * nn * 4 * world),
               'steps': e2e_steps, 'ms_per_step': ms_e2e / e2e_steps,
               'input': 'float32 stack in pinned host memory, uploaded in 8 node '
                        'slabs overlapped with K2 and the map download'}
        del host, maps
    except (RuntimeError, MemoryError) as exc:     # e.g. cannot pin 16 GB
        e2e = {'value': None, 'unit': UNIT, 'error': str(exc)[:200]}
    sampler.stop_flag.set()
    sampler.join(timeout=2)

    if rank == 0:
        peak, peak_src = peaks()
        k_ms = float(np.mean(kernel_ms))
        alg_bytes = 4.0 * R * nn + 4.0 * nplanes * nn
        achieved = alg_bytes / (k_ms * 1e-3) / 1e9
        cores = len(os.sched_getaffinity(0))
        cpu_nodes = int(os.environ.get('GSB_CPU_NODES', 262144))
        cpu_val, cpu_wall = cpu_sample(dims, 4096 if fast else cpu_nodes, 1)
        line = {
            'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world,
            'steps': args.steps, 'warmup': args.warmup,
            'ms_per_step': ms_total / args.steps, 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32',
            'data': 'synthetic',
            'config': {'workload': workload_name(world),
                       'l2': 'inputs larger than L2 ({0:.1f} GB stack per GPU)'
                       .format(R * nn * 4 / 1e9),
                       'detected': detected},
            'e2e': e2e,
            'gpu_launches': 3 * args.steps,
            'roofline': {'bound': 'hbm', 'achieved': achieved, 'peak': peak,
                         'unit': 'GB/s', 'frac': achieved / peak,
                         'frac_of_nominal_8TBs': achieved / 8000.0,
                         'traffic': measured_traffic(R, nn), 'kernel': 'k2_csort_kernel',
                         'kernel_ms': k_ms, 'algorithmic_bytes': alg_bytes,
                         'peak_source': peak_src},
            'cpu_baseline': {
                'value': cpu_val, 'unit': UNIT, 'cores': 1, 'kind': 'port',
                'host_cores_available': cores,
                'sample': ('{0} node columns x {1} realizations, NumPy oracle '
                           '(oracle/np_oracle.py), {2:.1f} s').format(
                               cpu_nodes, R, cpu_wall)},
            'clocks': sampler.summary(),
        }
        print(json.dumps(line))
    grids.dump()
    if world > 1:
        td.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200')
    args = ap.parse_args()
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == '__main__':
    main()
