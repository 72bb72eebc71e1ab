import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, 'ofa-for-super-resolution_b200'), os.path.join(ROOT, 'oracle'), ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')


@pytest.fixture(scope='session')
def golden():
    import json
    import numpy as np
    g = os.path.join(ROOT, 'tests', 'golden')
    arrays = np.load(os.path.join(g, 'reference_outputs.npz'))
    with open(os.path.join(g, 'reference_bookkeeping.json')) as f:
        book = json.load(f)
    return arrays, book
