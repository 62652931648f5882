import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')


@pytest.fixture
def rg():
    # the reference seeds its `rg` fixture from --random-seed (python/glow/conftest.py:54-61)
    return np.random.default_rng(int(os.environ.get('GLOW_TEST_SEED', '20261019')))


@pytest.fixture(scope='session')
def golden_dir():
    return GOLDEN
