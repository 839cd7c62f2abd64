import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


class Golden:
    def __init__(self):
        self.z = np.load(os.path.join(ROOT, "tests", "golden", "attack_update.npz"))

    def case(self, name):
        out = {}
        for k in self.z.files:
            if k.startswith(name + "/"):
                v = self.z[k]
                out[k.split("/", 1)[1]] = torch.from_numpy(v) if v.ndim > 0 else v.item()
        assert out, name
        return out


@pytest.fixture(scope="session")
def golden():
    return Golden()


@pytest.fixture(scope="session")
def built_lib():
    """Build (if needed) and load libb2attack.so."""
    from eval_driving_safety_b200 import build, _lib
    build.build()
    return _lib.load()
