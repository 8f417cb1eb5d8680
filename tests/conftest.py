import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "unvalidated: GPU test of code that has not run on a GPU yet; skipped unless "
                            "PPNP_TEST_UNVALIDATED=1.  Nothing carries it at the moment (round 2 ran every such case).")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Build the CUDA library and the C oracle once per session (nvcc cross-compiles without a GPU)."""
    import __graft_entry__ as g
    g.build()


# Hot-path rows of SURVEY section 8 first (a: kernels behind the C ABI, b: the shim boundary, e: multi-GPU), the
# "next" rows (f) after them: under `pytest -x` a failure in a widening row must not hide the hot path.
_ORDER = ["test_abi", "test_oracle_golden", "test_plan", "test_plan_properties", "test_gpu_parity", "test_gpu_variants",
          "test_gpu_tc", "test_gpu_exact", "test_gpu_batch", "test_shim", "test_gpu_dropin", "test_gpu_bench", "test_dist_cpu",
          "test_gpu_dist", "test_bench_cpu"]


def _rank(item):
    name = os.path.splitext(os.path.basename(str(item.fspath)))[0]
    return _ORDER.index(name) if name in _ORDER else len(_ORDER)


def pytest_collection_modifyitems(config, items):
    import torch
    items.sort(key=_rank)          # stable: the order inside a file is kept
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
