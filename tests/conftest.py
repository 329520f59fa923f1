import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import pyoracle
    pyoracle.build()
    return pyoracle


@pytest.fixture(scope="session")
def synth():
    import slam_module_b200 as sm
    return sm.synth


@pytest.fixture(scope="session")
def golden():
    import numpy as np
    d = ROOT / "tests" / "golden"
    return {"cv2": np.load(d / "golden_cv2.npz"), "ref": np.load(d / "golden_ref_leaves.npz")}


@pytest.fixture(scope="session")
def slamgpu():
    """The ctypes binding of the CUDA library.  GPU tests fail (not skip) when it is missing."""
    from slam_module_b200 import slamgpu as sg
    sg.lib()
    return sg
