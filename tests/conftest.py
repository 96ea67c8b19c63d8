import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")
if GOLDEN not in sys.path:
    sys.path.insert(0, GOLDEN)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (sm_100) GPU; run with -m gpu on the GPU box")


def pytest_collection_modifyitems(config, items):
    # `-m gpu` tests must fail loudly (not skip) when the CUDA path is unusable on a GPU box;
    # on a CPU-only box they are deselected by `-m "not gpu"`.
    pass


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN
