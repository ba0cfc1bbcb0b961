import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")
    config.addinivalue_line(
        "markers", "needs_reference: imports the unmodified reference from /root/reference (build container only)"
    )


def pytest_collection_modifyitems(config, items):
    have_ref = os.path.isfile("/root/reference/src/environment/yard.py")
    skip_ref = pytest.mark.skip(reason="/root/reference not present on this box")
    for item in items:
        if "needs_reference" in item.keywords and not have_ref:
            item.add_marker(skip_ref)


@pytest.fixture(scope="session")
def golden_traces():
    import numpy as np

    return np.load(os.path.join(GOLDEN, "traces.npz"))
