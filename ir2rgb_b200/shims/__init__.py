"""ctypes stand-ins for the reference's three pybind11 modules ("Option B").

`correlation_cuda`, `resample2d_cuda` and `channelnorm_cuda` here have the call signatures of the compiled
modules the reference's wrappers import (correlation_cuda.cc:169-172, resample2d_cuda.cc:25-28,
channelnorm_cuda.cc:28-31) and forward to libflowops.so.  A maintainer who wants to keep the reference's own
`correlation.py` / `resample2d.py` / `channelnorm.py` untouched puts these on the import path under those names:

    import ir2rgb_b200.shims as shims
    shims.install()            # sys.modules["correlation_cuda"] = shims.correlation_cuda, ...

`tests/test_shims_gpu.py` loads the reference's unmodified wrappers on top of them.
"""
import sys

from . import channelnorm_cuda, correlation_cuda, resample2d_cuda

NAMES = ("correlation_cuda", "resample2d_cuda", "channelnorm_cuda")


def install():
    """Register the shims under the module names the reference's wrappers import."""
    for name in NAMES:
        sys.modules[name] = globals()[name]


def uninstall():
    for name in NAMES:
        if sys.modules.get(name) is globals()[name]:
            del sys.modules[name]
