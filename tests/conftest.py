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


@pytest.fixture(scope="session")
def O():
    """The CPU oracle (test infrastructure; built on demand with oracle/Makefile)."""
    from oracle import oracle
    oracle.lib()
    return oracle


@pytest.fixture(scope="session")
def lib_built():
    """libplane_ransac.so, built in-tree if it is not there yet (nvcc cross-compiles without a GPU)."""
    from dialog_b200 import build
    return build.build()


@pytest.fixture(scope="session")
def double_shadow():
    xyz = np.load(os.path.join(GOLDEN, "double_shadow_xyz.npy"))
    assert xyz.shape == (991, 3) and xyz.dtype == np.float32
    return xyz


@pytest.fixture(scope="session")
def scene2():
    from dialog_b200 import synth
    return synth.three_planes_scene()


@pytest.fixture(scope="session")
def scene3():
    from dialog_b200 import synth
    return synth.indoor_scene()
