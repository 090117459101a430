# -*- coding: utf-8 -*-
"""The two helpers of ``GSIMCLI/tools/utils.py`` the hot path depends on."""
import os


def skip_lines(file_id, nlines):
    """Advance a file object by ``nlines`` lines (utils.py:66-78)."""
    for _ in range(nlines):
        file_id.readline()


def filename_indexing(file_id, n):
    """``base.ext`` -> ``base_<n>.ext``: the naming of DSS realization files,
    ``<out>_sim.out, <out>_sim_2.out, ...`` (utils.py:104-123)."""
    base, ext = os.path.splitext(file_id)
    return '{0}_{1}{2}'.format(base, n, ext)
