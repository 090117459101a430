# -*- coding: utf-8 -*-
"""SURVEY section 4 tier (c): the node-sharded path on 1 / 2 / 4 / 8 GPUs must
give byte-identical maps, station tables and detection flags, equal to the
oracle.  One rank per GPU (torch.distributed.run, NCCL); the only exchange is
the all-gather of the station tables.  Needs >= 2 GPUs (skipped otherwise)."""
import json
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _run(world, out):
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1',
           '--nproc-per-node', str(world), '--master-addr', '127.0.0.1',
           '--master-port', str(_free_port()),
           os.path.join(ROOT, 'tests', 'mgpu_worker.py'), out]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    with open(out) as fh:
        return json.load(fh)


@pytest.mark.skipif(_ngpus() < 2, reason='needs at least 2 GPUs')
def test_sharded_results_identical_across_world_sizes(tmp_path):
    worlds = [w for w in (1, 2, 4, 8) if w <= _ngpus()]
    res = [_run(w, str(tmp_path / ('w%d.json' % w))) for w in worlds]
    for r in res:
        assert r['grid_ok'] and r['stations_ok'] and r['tables_identical_on_all_ranks'], r
    for key in ('maps_crc', 'flags_crc', 'homog_crc', 'table_crc', 'detected'):
        assert len({r[key] for r in res}) == 1, (key, res)
