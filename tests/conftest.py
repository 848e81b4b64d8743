import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: test needs a CUDA device (run on the B200 box with -m gpu)')


@pytest.fixture(scope='session')
def golden():
    import numpy as np
    return np.load(os.path.join(ROOT, 'tests', 'golden', 'export_golden.npz'))


@pytest.fixture(scope='session')
def robot():
    from jogramop_framework_b200.ingest import RobotModel
    return RobotModel.load(os.path.join(ROOT, 'jogramop_framework_b200', 'data', 'robot_mobile_panda.npz'))


@pytest.fixture(scope='session')
def scene_loader():
    from jogramop_framework_b200.ingest import SceneModel
    cache = {}

    def load(sid, variant=''):
        key = (sid, variant)
        if key not in cache:
            cache[key] = SceneModel.load(os.path.join(ROOT, 'jogramop_framework_b200', 'data',
                                                      f'scene_{sid:03d}{variant}.npz'))
        return cache[key]
    return load
