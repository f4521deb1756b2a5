import os
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

GOLDEN = os.path.join(REPO, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')


def golden_data(rel):
    return os.path.join(GOLDEN, 'data', rel)


def golden_gen(rel):
    return os.path.join(GOLDEN, 'gen', rel)


@pytest.fixture(scope='session')
def oracle():
    from oracle import khmer_oracle
    return khmer_oracle


@pytest.fixture(scope='session', autouse=True)
def _build_native():
    """Make sure both shared libraries exist (the GPU box receives them prebuilt)."""
    import __graft_entry__ as entry
    entry.build(quiet=True)
