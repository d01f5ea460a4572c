import json
import os
import sys
import warnings

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
warnings.filterwarnings("ignore", category=DeprecationWarning)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")


@pytest.fixture(scope="session")
def golden():
    """Vectors produced by the reference itself (oracle/make_golden.py)."""
    return dict(np.load(os.path.join(GOLDEN, "reference_vectors.npz")))


@pytest.fixture(scope="session")
def golden_vad():
    with open(os.path.join(GOLDEN, "vad_state_machines.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def gpu():
    """Loads libosb200 and requires a device; every -m gpu test goes through the C ABI."""
    from open_speech_b200 import _native as N

    N.require_gpu()
    N.check(N.lib().osb_init(0))
    return N
