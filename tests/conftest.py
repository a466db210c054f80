"""pytest configuration: registers the ``gpu`` marker and puts the project dirs on sys.path."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG_DIR = os.path.join(ROOT, 'recbole-cdr_b200')
for p in (ROOT, PKG_DIR):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN_DIR = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: test needs a CUDA device (run on the B200 box with -m gpu)')
    config.addinivalue_line('markers', 'unvalidated: gpu test of a kernel written while no GPU was reachable and not yet run '
                                       'on hardware (its logic is covered by the CPU emulator tests); skipped unless '
                                       'XDR_RUN_UNVALIDATED=1 -- the first gpurun call of the next session runs them')


def pytest_collection_modifyitems(config, items):
    if os.environ.get('XDR_RUN_UNVALIDATED', '0') != '1':
        skip_unv = pytest.mark.skip(reason='kernel not yet validated on hardware (set XDR_RUN_UNVALIDATED=1 to run)')
        for item in items:
            if 'unvalidated' in item.keywords:
                item.add_marker(skip_unv)
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:  # pragma: no cover
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason='no CUDA device in this container')
    for item in items:
        if 'gpu' in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope='session')
def golden_dir():
    return GOLDEN_DIR
