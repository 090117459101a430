# -*- coding: utf-8 -*-
import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (B200)')
    config.addinivalue_line(
        'markers', 'refshim: runs the reference source (build container only)')


def load_golden(name):
    with open(os.path.join(GOLDEN, name + '.json')) as fh:
        return json.load(fh)


GOLDEN_CASES = ['kat_appendix_c', 'dyadic_even', 'dyadic_odd', 'decimals_lf',
                'header3_stats']


@pytest.fixture(params=GOLDEN_CASES)
def golden(request):
    return load_golden(request.param)
