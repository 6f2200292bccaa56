import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(autouse=True, scope="session")
def _true_fp32_contractions():
    """parity is stated against the reference's fp32 arithmetic (its CPU path is the pinned oracle): the convolutions
    and matmuls the path does not own must not run in TF32 (torch's cuDNN default) while they are compared with it"""
    try:
        import torch
        torch.backends.cudnn.allow_tf32 = False
        torch.backends.cuda.matmul.allow_tf32 = False
    except Exception:  # pragma: no cover
        pass
    yield


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:  # pragma: no cover
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


class Golden:
    """tests/golden/<name>.npz with keys '<case>.<field>'"""

    def __init__(self, name):
        self.z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))

    def cases(self):
        return sorted({k.split(".", 1)[0] for k in self.z.files})

    def case(self, c):
        pre = c + "."
        return {k[len(pre):]: self.z[k] for k in self.z.files if k.startswith(pre)}

    def __getitem__(self, k):
        return self.z[k]


def golden(name):
    return Golden(name)


def assert_close(a, b, rtol=1e-5, atol=None, what=""):
    """relative 1e-5 in fp32 (north_star); atol scales with the tensor's magnitude for near-zero entries"""
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, f"{what}: shape {a.shape} vs {b.shape}"
    if atol is None:
        atol = rtol * (np.abs(b).max() if b.size else 0.0) + 1e-30
    err = np.abs(a - b)
    bad = err > (atol + rtol * np.abs(b))
    assert not bad.any(), f"{what}: {bad.sum()} / {a.size} mismatches, max err {err.max():.3e} (ref max {np.abs(b).max():.3e})"


def assert_exact(a, b, what=""):
    a = np.asarray(a); b = np.asarray(b)
    assert a.shape == b.shape, f"{what}: shape {a.shape} vs {b.shape}"
    assert np.array_equal(a, b), f"{what}: {(a != b).sum()} / {a.size} elements differ (bit-exact required)"
