import os
import shutil
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on a B200)")


@pytest.fixture(scope="session")
def flowops_lib():
    """libflowops.so, (re)built in-tree when nvcc is present; otherwise the prebuilt file is used."""
    from ir2rgb_b200 import _lib, build
    if shutil.which("nvcc") or os.path.exists("/usr/local/cuda/bin/nvcc"):
        build.build()
    return _lib.load()


@pytest.fixture(scope="session")
def c_oracle():
    from oracle import c_oracle as co
    co.build()
    return co


@pytest.fixture(scope="session")
def golden_native():
    import numpy as np
    path = os.path.join(GOLDEN, "native_ops_ref_sm100.npz")
    if not os.path.exists(path):
        pytest.skip("tests/golden/native_ops_ref_sm100.npz not generated yet (tests/golden/make_golden_gpu.py)")
    return np.load(path)


@pytest.fixture(scope="session")
def golden_resample():
    import numpy as np
    return np.load(os.path.join(GOLDEN, "resample_cpu.npz"))
