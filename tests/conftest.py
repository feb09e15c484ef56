import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "src"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box: pytest -m gpu)")
    config.addinivalue_line("markers", "multigpu: needs >= 2 CUDA devices")


def pytest_collection_modifyitems(config, items):
    import torch
    have = torch.cuda.is_available()
    n = torch.cuda.device_count() if have else 0
    for it in items:
        if "gpu" in it.keywords and not have:
            it.add_marker(pytest.mark.skip(reason="no CUDA device"))
        if "multigpu" in it.keywords and n < 2:
            it.add_marker(pytest.mark.skip(reason="needs >= 2 GPUs"))


@pytest.fixture
def repo_root():
    return ROOT
