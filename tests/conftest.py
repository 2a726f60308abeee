import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    """`gpu` tests need a CUDA device AND the built extension: skip them (instead of failing at the first CUDA call)
    on a box without either, so a plain `pytest tests` works everywhere."""
    reason = None
    try:
        import torch
        if not torch.cuda.is_available():
            reason = "needs a CUDA device"
    except Exception as e:  # pragma: no cover
        reason = f"torch unavailable: {e}"
    if reason is None:
        lib = os.path.join(ROOT, "continuous-time-diffusion-models-for-discrete-data_b200", "libctdd_b200.so")
        if not os.path.exists(lib):
            reason = "libctdd_b200.so is not built (python -c 'import __graft_entry__ as g; g.build()')"
    if reason is None:
        return
    skip = pytest.mark.skip(reason=reason)
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden():
    import numpy as np
    d = os.path.join(ROOT, "tests", "golden")
    return {name[:-4]: np.load(os.path.join(d, name)) for name in os.listdir(d) if name.endswith(".npz")}
