import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu under gpurun)")


def pytest_collection_modifyitems(config, items):
    """GPU tests fail loudly (never skip silently) when selected on a box without CUDA; when they
    are not explicitly selected with -m gpu and no GPU is present they are skipped."""
    import torch
    if torch.cuda.is_available():
        return
    markexpr = config.getoption("-m") or ""
    if "gpu" in markexpr and "not gpu" not in markexpr:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden():
    import numpy as np

    def load(name):
        return np.load(os.path.join(GOLDEN, name), allow_pickle=False)
    return load


@pytest.fixture(autouse=True)
def _bf16_contractions_by_default(request):
    """GPU tests written against the bf16-operand policy (the oracle's policy="bf16") pin it; the fp32-faithful
    mode that the drop-ins select outside torch.autocast has its own tests (tests/test_gpu_fp32_mode.py), which
    set the precision themselves."""
    import torch
    if "gpu" not in request.keywords or not torch.cuda.is_available():
        yield
        return
    from dinox_b200 import losshead
    prev = losshead.set_contraction_precision("bf16")
    yield
    losshead.set_contraction_precision(prev)
