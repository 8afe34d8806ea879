import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu under gpurun)")
    # a fresh checkout has no built artefacts (they are git-ignored): build the product library and the C oracle once
    # (nvcc cross-compiles without a GPU; about a minute).  The product itself never builds or falls back at import.
    import shutil
    import subprocess
    lib = os.path.join(ROOT, "supersdr_b200", "libssdr_b200.so")
    if not os.path.isfile(lib) and shutil.which("nvcc") and shutil.which("make"):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "supersdr_b200", "csrc"), "-j8"], stdout=subprocess.DEVNULL)
    if not os.path.isfile(os.path.join(ROOT, "oracle", "_build", "libssdr_oracle.so")) and shutil.which("make"):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")], stdout=subprocess.DEVNULL)


def _has_gpu():
    try:
        import ctypes
        cuda = ctypes.CDLL("libcuda.so.1")
        n = ctypes.c_int(0)
        return cuda.cuInit(0) == 0 and cuda.cuDeviceGetCount(ctypes.byref(n)) == 0 and n.value > 0
    except OSError:
        return False


HAS_GPU = _has_gpu()


def pytest_collection_modifyitems(config, items):
    if HAS_GPU:
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container (run under gpurun)")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def ssdr():
    import supersdr_b200
    supersdr_b200.init()
    return supersdr_b200
