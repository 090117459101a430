# -*- coding: utf-8 -*-
"""The helper of ``GSIMCLI/tools/utils.py`` the hot path depends on.
(``skip_lines``, utils.py:66-78, has no counterpart: header lines are skipped by
the K1 decoder, ``gsb_decode_text(header_lines=...)``.)"""
import os


def filename_indexing(file_id, n):
    """``base.ext`` -> ``base_<n>.ext``: the naming of DSS realization files,
    ``<out>_sim.out, <out>_sim_2.out, ...`` (utils.py:104-123)."""
    base, ext = os.path.splitext(file_id)
    return '{0}_{1}{2}'.format(base, n, ext)
