import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, 'tests', 'golden')
REFERENCE = '/root/reference'


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if 'gpu' in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope='session')
def golden_dir():
    return GOLDEN


@pytest.fixture(scope='session')
def decode_cases():
    data = np.load(os.path.join(GOLDEN, 'decode_cases.npz'))
    with open(os.path.join(GOLDEN, 'decode_cases.json')) as fh:
        meta = json.load(fh)
    return meta, data


@pytest.fixture(scope='session')
def read_basecalls():
    with open(os.path.join(GOLDEN, 'reads_basecalls.json')) as fh:
        return {r['name']: r for r in json.load(fh)}


@pytest.fixture(scope='session')
def reads_daq():
    return np.load(os.path.join(GOLDEN, 'reads_daq.npz'))


@pytest.fixture(scope='session')
def pretrained():
    """The pretrained network rebuilt from the committed weight fixture (no reference needed)."""
    from sloika_b200 import zoo
    with open(os.path.join(GOLDEN, 'pretrained_arch.json')) as fh:
        arch = json.load(fh)
    weights = dict(np.load(os.path.join(GOLDEN, 'pretrained_weights.npz')))
    return zoo.from_weights(arch, weights)


def scaled_signal(daq_file, name):
    """pA signal of a bundled read from the DAQ fixture: (daq + offset) * range / digitisation."""
    offset, rng, digi = daq_file[name + '_scaling']
    return (daq_file[name] + offset) * (rng / digi)


needs_reference = pytest.mark.skipif(not os.path.isdir(REFERENCE), reason="/root/reference not mounted")


@pytest.fixture(scope='session')
def forward_cases():
    """Outputs of the reference's own layers.py / conv.py (tools/make_golden_forward.py): (meta, arrays)."""
    data = np.load(os.path.join(GOLDEN, 'forward_cases.npz'))
    with open(os.path.join(GOLDEN, 'forward_cases.json')) as fh:
        meta = json.load(fh)
    return meta, data


def case_weights(data, name):
    prefix = name + '/w/'
    return {k[len(prefix):]: data[k] for k in data.files if k.startswith(prefix)}
