import os
import sys

import pytest

os.environ.setdefault("PP_CUDNN_BENCHMARK", "0")  # short test runs: no cuDNN autotuning (pixelpick_b200/args.py)

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden():
    import numpy as np
    return np.load(os.path.join(ROOT, "tests", "golden", "query_golden.npz"))


@pytest.fixture(scope="session", autouse=True)
def _built_library():
    """The C-ABI library must exist for every test run (CPU runs only load it and look up symbols)."""
    import __graft_entry__ as g
    if not os.path.exists(os.path.join(ROOT, "pixelpick_b200", "csrc", "libpixelpick_b200.so")):
        g.build()
