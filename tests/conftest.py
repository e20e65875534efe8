import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (sm_100a) GPU; run with -m gpu under gpurun")


@pytest.fixture(scope="session")
def golden():
    import torch
    return torch.load(os.path.join(ROOT, "tests", "golden", "bidatenet_golden.pt"))


@pytest.fixture(scope="session")
def built_lib():
    """Build (if stale) and load the C-ABI library; CPU tests only check that it loads and exports."""
    import __graft_entry__ as g
    g.build()
    from fabric_b200 import _lib
    return _lib.load()


@pytest.fixture(autouse=True)
def _gpu_quiesce(request):
    """GPU tests start and end with an idle device: no test inherits another test's in-flight work (side-stream copies,
    deferred frees), so a failure always belongs to the test that reports it."""
    if request.node.get_closest_marker("gpu") is None:
        yield
        return
    import torch
    if torch.cuda.is_available():
        torch.cuda.synchronize()
    yield
    if torch.cuda.is_available():
        torch.cuda.synchronize()
