import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


class Golden(object):
    """A tests/golden/*.npz cell fixture with the CellResult interface of oracles.py."""

    def __init__(self, name):
        d = np.load(os.path.join(GOLDEN, name))
        self.xyz = d["xyz"]
        self.w = d["w"] if "w" in d.files else None
        self.m = d["m"]
        self.mg = d["mg"]
        self.n = self.m.shape[0]
        ptr, ids = d["face_ptr"], d["face_ids"]
        self.nfaces = np.diff(ptr).astype(np.uint32)
        self.fcap = 64
        self.face_nbr = np.zeros((self.n, 64), np.uint32)
        self.face_area = np.zeros((self.n, 64)) if "face_area" in d.files else None
        for i in range(self.n):
            self.face_nbr[i, : self.nfaces[i]] = ids[ptr[i]: ptr[i + 1]]
            if self.face_area is not None:
                self.face_area[i, : self.nfaces[i]] = d["face_area"][ptr[i]: ptr[i + 1]]
        self.extra = {k: d[k] for k in d.files}


@pytest.fixture(scope="session")
def port():
    import oracles
    return oracles.Port()


@pytest.fixture(scope="session")
def reference():
    import oracles
    if not oracles.have_reference():
        pytest.skip("oracle/_ref/libgeoref.so not built (needs /root/reference)")
    return oracles.Reference()
