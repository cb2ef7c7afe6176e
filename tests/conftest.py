import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

TESTS_DIR = os.path.dirname(os.path.abspath(__file__))
if TESTS_DIR not in sys.path:
    sys.path.insert(0, TESTS_DIR)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle_lib():
    from oracle import oracle_py as O
    O.build(ref=os.path.isdir("/root/reference"))
    return O


@pytest.fixture(scope="session", params=["random20k", "repeat30k"])
def golden(request):
    z = np.load(os.path.join(GOLDEN, request.param + ".npz"))
    g = {k: z[k] for k in z.files}
    g["name"] = request.param
    g["n_opts"] = sum(1 for k in z.files if k.startswith("opt") and k[3:].isdigit())
    return g


@pytest.fixture(scope="session")
def cuda_lib():
    """Builds (if stale) and loads the CUDA library; GPU tests fail loudly when it is unusable."""
    from compseed_b200 import build as B
    B.build()
    import compseed_b200 as cs
    cs.load_library()
    assert cs.device_count() > 0, "no CUDA device: gpu-marked tests must run on the B200 box"
    return cs
