import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: test needs a CUDA device (B200); run with -m gpu')
    config.addinivalue_line('markers', 'slow: longer duplicate / calibration runs, executed with CATB_SLOW_TESTS=1')


def pytest_collection_modifyitems(config, items):
    """The default CPU suite is sized to a few minutes: tests marked `slow` repeat, at greater length or in a second
    numerical mode, what a faster test of the same code already establishes; CATB_SLOW_TESTS=1 runs them too."""
    if os.environ.get('CATB_SLOW_TESTS', '0') == '1':
        return
    skip = pytest.mark.skip(reason='set CATB_SLOW_TESTS=1')
    for item in items:
        if 'slow' in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope='session')
def golden_dir():
    return os.path.join(ROOT, 'tests', 'golden')
