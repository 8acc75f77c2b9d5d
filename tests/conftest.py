import os
import sys

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
for p in (REPO, os.path.join(HERE, "golden")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


class Golden:
    """Lazy access to tests/golden/*.npz (outputs of the reference itself, see make_golden.py)."""

    def __init__(self):
        self._files = {}

    def __call__(self, name):
        if name not in self._files:
            self._files[name] = np.load(os.path.join(HERE, "golden", name + ".npz"))
        return self._files[name]

    def coo(self, fname, prefix):
        d = self(fname)
        return d[prefix + "_row"], d[prefix + "_col"], d[prefix + "_val"], tuple(int(v) for v in d[prefix + "_shape"])

    def pt_coo(self, fname, prefix, device="cpu", coalesced=True):
        d = self(fname)
        t = torch.sparse_coo_tensor(torch.from_numpy(d[prefix + "_idx"]), torch.from_numpy(d[prefix + "_val"]),
                                    tuple(int(v) for v in d[prefix + "_shape"]), is_coalesced=coalesced)
        return t.to(device)


@pytest.fixture(scope="session")
def golden():
    return Golden()


def assert_close(actual, expected, rtol, atol, what=""):
    actual = np.asarray(actual, dtype=np.float64)
    expected = np.asarray(expected, dtype=np.float64)
    assert actual.shape == expected.shape, "%s: shape %s vs %s" % (what, actual.shape, expected.shape)
    err = np.abs(actual - expected)
    tol = atol + rtol * np.abs(expected)
    bad = err > tol
    if bad.any():
        i = np.unravel_index(np.argmax(err - tol), err.shape)
        raise AssertionError("%s: %d / %d elements out of tolerance (rtol %g atol %g); worst at %s: got %r expected %r"
                             % (what, bad.sum(), bad.size, rtol, atol, i, actual[i], expected[i]))
