import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "unvalidated: GPU test of code written after the round's GPU budget was spent; it has "
                            "never run on a GPU and is skipped unless PPNP_TEST_UNVALIDATED=1 (tools/gpu_calls/r02_first_call.sh sets it)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Build the CUDA library and the C oracle once per session (nvcc cross-compiles without a GPU)."""
    import __graft_entry__ as g
    g.build()


def pytest_collection_modifyitems(config, items):
    import torch
    if os.environ.get("PPNP_TEST_UNVALIDATED") != "1":
        hold = pytest.mark.skip(reason="not yet validated on a GPU (written after the round's GPU budget was spent): "
                                       "set PPNP_TEST_UNVALIDATED=1 to run it")
        for it in items:
            if "unvalidated" in it.keywords:
                it.add_marker(hold)
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)
