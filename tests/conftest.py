import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "reference: needs the read-only reference tree at /root/reference")


@pytest.fixture(scope="session")
def golden_pc():
    import json
    import numpy as np
    with open(os.path.join(GOLDEN_DIR, "pc_transform.json")) as f:
        meta = json.load(f)
    return meta, np.load(os.path.join(GOLDEN_DIR, "pc_transform.npz"))


@pytest.fixture(scope="session")
def golden_small():
    import numpy as np
    return np.load(os.path.join(GOLDEN_DIR, "small_cases.npz"))
