import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "ref: needs /root/reference (build container only)")


@pytest.fixture(scope="session")
def golden():
    return np.load(os.path.join(GOLDEN, "reference_vectors.npz"))


@pytest.fixture(scope="session")
def etg_default(golden):
    """W,b of Opt_with_points(Footheight=0.1, Steplength=0.05) as produced by the reference's own train.py code."""
    return golden["opt_w0"], golden["opt_b0"]


@pytest.fixture(scope="session")
def etg_stable():
    """A gentler gait (Footheight 0.03, Steplength 0.02) that walks >1000 steps open-loop: long-horizon drift tests."""
    from paddlerobotics_b200.etg import ETG_layer, Opt_with_points
    layer = ETG_layer(0.5, 0.026, 20, 0.04, np.array([-np.pi / 2, 0]), 0.2, 0.5)
    w, b, _ = Opt_with_points(ETG=layer, ETG_T=0.5, Footheight=0.03, Steplength=0.02)
    return w, b
