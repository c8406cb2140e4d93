import pathlib
import sys
import zlib

import numpy as np
import pytest

ROOT = pathlib.Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: test needs a CUDA device (B200)')


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        have = torch.cuda.is_available()
    except Exception:  # pragma: no cover
        have = False
    if have:
        return
    skip = pytest.mark.skip(reason='no CUDA device')
    for item in items:
        if 'gpu' in item.keywords:
            item.add_marker(skip)


@pytest.fixture
def rng(request):
    """ per-test deterministic generator seeded from the test id (like the reference's tests/conftest.py:37-42) """
    seed = zlib.crc32(request.node.nodeid.encode())
    return np.random.default_rng(seed)
