import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "knode-cosserat_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    """`pytest tests` on a machine without a CUDA device (or without the built library) skips the gpu-marked tests instead
    of failing them; the GPU box runs them with -m gpu."""
    try:
        import torch
        have_gpu = torch.cuda.is_available()
    except Exception:
        have_gpu = False
    have_lib = os.path.exists(os.path.join(PKG, "libknode_cosserat_b200.so"))
    if have_gpu and have_lib:
        return
    skip = pytest.mark.skip(reason="no CUDA device" if not have_gpu else "libknode_cosserat_b200.so not built")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden():
    import numpy as np

    class _G:
        def __init__(self):
            self._c = {}

        def __getitem__(self, name):
            if name not in self._c:
                self._c[name] = np.load(os.path.join(GOLDEN, name + ".npz"))
            return self._c[name]

    return _G()
