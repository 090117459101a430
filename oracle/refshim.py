# -*- coding: utf-8 -*-
"""Import-time shim that executes the reference's OWN hot-path sources.

TEST INFRASTRUCTURE ONLY.  Nothing under ``gsimcli_b200/`` may import this
module; it is used by ``oracle/make_golden.py`` (to produce the committed
fixtures under ``tests/golden/``) and by the ``refshim``-marked tests that run
only where ``/root/reference`` exists (this build container, never the GPU box).

The reference (iled/gsimcli) is Python 2.7 and depends on ``bottleneck``,
neither of which exists here.  Nothing under ``/root/reference`` is modified or
copied: ``tools/utils.py``, ``tools/grid.py`` and ``tools/homog.py`` are read
as text, six py2->py3 textual substitutions and three one-line repairs of
genuine HEAD breakage are applied in memory, and the result is exec'd into
fresh modules.  See SURVEY.md section 8(c) / Appendix B for the rationale of
every substitution:

  py2->py3 : ``except X, e`` / ``print x`` / ``xrange`` / ``.iteritems()`` /
             ``raw_input`` / ``isinstance(f, file)``
  R1 grid.py:99      ``values = values or np.zeros((0, 0))`` raises for any
                     ndarray/DataFrame -> ``None`` test.
  R2 homog.py:204-5  ``Series.between(Series, Series)`` with differently
                     labelled indices -> positional ``.values``.
  R3 homog.py:330-2  ragged row list rejected by modern NumPy -> ``np.hstack``.

``bottleneck`` is replaced by a NumPy-backed stub (its nan* functions are
documented drop-ins for ``np.nan*``; only summation order differs).
"""
import os
import re
import sys
import types

import numpy as np
import pandas  # noqa: F401  (must be imported BEFORE the bottleneck stub)

REF_ROOT = os.environ.get('GSIMCLI_REFERENCE', '/root/reference')
REF = os.path.join(REF_ROOT, 'GSIMCLI')


def available():
    """True when the reference checkout can be read (build container only)."""
    return os.path.isfile(os.path.join(REF, 'tools', 'grid.py'))


def _install_bottleneck_stub():
    bn = types.ModuleType('bottleneck')

    def _replace(a, old, new):  # in-place, like bn.replace
        if old != old:
            a[np.isnan(a)] = new
        else:
            a[a == old] = new

    bn.replace = _replace
    bn.nanmean = np.nanmean
    bn.nanmedian = np.nanmedian
    bn.nanvar = lambda a, ddof=0: np.nanvar(a, ddof=ddof)
    bn.nanstd = lambda a, ddof=0: np.nanstd(a, ddof=ddof)
    bn.__version__ = '0.0-stub'
    sys.modules['bottleneck'] = bn


def _py3ify(src):
    src = re.sub(r'except (\w+), (\w+):', r'except \1 as \2:', src)
    src = re.sub(r'^(\s*)print (.+)$', r'\1print(\2)', src, flags=re.M)
    src = (src.replace('xrange', 'range')
              .replace('.iteritems()', '.items()')
              .replace('raw_input', 'input'))
    src = src.replace('isinstance(self.files[0], file)',
                      'hasattr(self.files[0], "readline")')
    # R1
    src = src.replace('values = values or np.zeros((0, 0))',
                      'values = np.zeros((0, 0)) if values is None else values')
    # R2
    src = (src.replace("between(local_stats['lperc'],",
                       "between(local_stats['lperc'].values,")
              .replace("local_stats['rperc'])\n    detected_number",
                       "local_stats['rperc'].values)\n    detected_number"))
    # R3
    src = (src.replace('filled[i, :] = [pset.values.iloc[0, 0]',
                       'filled[i, :] = np.hstack([np.atleast_1d(_v) for _v in '
                       '[pset.values.iloc[0, 0]')
              .replace('][:pset.values.shape[1]]',
                       ']])[:pset.values.shape[1]]'))
    return src


def _load(name, path):
    mod = types.ModuleType(name)
    mod.__file__ = path
    sys.modules[name] = mod
    with open(path, encoding='utf-8') as fh:
        src = fh.read()
    exec(compile(_py3ify(src), path, 'exec'), mod.__dict__)
    return mod


_CACHE = {}


def load_reference():
    """Return ``(utils, grid, homog)`` modules built from the reference text."""
    if _CACHE:
        return _CACHE['utils'], _CACHE['grid'], _CACHE['homog']
    if not available():
        raise RuntimeError('reference checkout not found at ' + REF)
    _install_bottleneck_stub()
    # the shimmed modules live under a private package name so they can never
    # shadow gsimcli_b200.tools
    tools = types.ModuleType('tools')
    tools.__path__ = []
    sys.modules['tools'] = tools
    utils = _load('tools.utils', os.path.join(REF, 'tools', 'utils.py'))
    grid = _load('tools.grid', os.path.join(REF, 'tools', 'grid.py'))
    homog = _load('tools.homog', os.path.join(REF, 'tools', 'homog.py'))
    tools.utils, tools.grid, tools.homog = utils, grid, homog
    _CACHE.update(utils=utils, grid=grid, homog=homog)
    return utils, grid, homog
