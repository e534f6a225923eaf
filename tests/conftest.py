import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

SCENES = ['scene%d' % i for i in range(11)]


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')


def scene_path(name):
    return os.path.join(ROOT, 'scenes', name + '.json')


@pytest.fixture(scope='session')
def built():
    """libpt_cuda.so and the oracle, built in-tree (idempotent)."""
    import __graft_entry__ as g
    g.build()
    return True


@pytest.fixture(scope='session')
def ptlib(built):
    import pathtracer_b200 as pt
    return pt


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)
