import json
import os
import sys
from pathlib import Path

import numpy as np
import pytest

REPO = Path(__file__).resolve().parent.parent
GOLDEN = REPO / 'tests' / 'golden'
if str(REPO) not in sys.path:
    sys.path.insert(0, str(REPO))


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a real B200 (run with -m gpu under gpurun)')


def pytest_collection_modifyitems(config, items):
    """`gpu` tests are skipped (not failed) when no CUDA device is visible, so a bare `pytest tests/` works anywhere."""
    import torch

    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason='no CUDA device')
    for item in items:
        if 'gpu' in item.keywords:
            item.add_marker(skip)


def load_json(name):
    return json.loads((GOLDEN / name).read_text())


def load_npz(name):
    return np.load(GOLDEN / name)


def decode_entry(e):
    """tests/golden history entry -> numpy array / python value / None (see oracle/gen_golden.py: enc)."""
    if e is None:
        return None
    if 'py' in e:
        return e['py']
    return np.asarray(e['data'], dtype=e['dtype']).reshape(e['shape'])


@pytest.fixture
def free_port():
    import socket

    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
# a protocol bug must fail a test after a minute, not hold the GPU for libdmlb's default ten (spawned ranks inherit this)
os.environ.setdefault('DMLB_PEER_TIMEOUT', '90')
