"""pytest configuration: registers the `gpu` marker and shared fixtures/helpers."""
from __future__ import annotations

import hashlib
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")

# parity gate of BASELINE.json's north_star: |x-y| <= 1e-14 + 1e-10 |y| elementwise
ATOL = 1e-14
RTOL = 1e-10


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:  # pragma: no cover
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def sha(a: np.ndarray) -> str:
    a = np.ascontiguousarray(a) + 0.0  # fold -0.0 into +0.0
    return hashlib.sha256(a.tobytes()).hexdigest()


def assert_parity(x, y, what=""):
    """north_star gate; also requires identical shape and dtype."""
    x = np.asarray(x); y = np.asarray(y)
    assert x.shape == y.shape, f"{what}: shape {x.shape} != {y.shape}"
    assert x.dtype == y.dtype, f"{what}: dtype {x.dtype} != {y.dtype}"
    err = np.abs(x - y)
    tol = ATOL + RTOL * np.abs(y)
    bad = ~(err <= tol)
    assert not bad.any(), f"{what}: {bad.sum()} of {bad.size} elements outside 1e-10 rel/1e-14 abs; max err {np.nanmax(err)}"


def random_triple(n, batch=(), seed=None):
    """The reference's synthetic-triple recipe (tests/test_math/test_lattice/test_vanilla.py:24-35), restated."""
    rng = np.random.RandomState(seed)
    A = rng.random((*batch, n, n)) + 1j * rng.random((*batch, n, n))
    A = A + np.swapaxes(A, -1, -2)
    A /= np.abs(np.linalg.eigvals(A)).max() + 0.2
    b = rng.random((*batch, n)) + 1j * rng.random((*batch, n))
    c = rng.random(batch) + 1j * rng.random(batch)
    return A, b, c


@pytest.fixture(scope="session")
def golden():
    return np.load(os.path.join(GOLDEN, "vanilla_golden.npz"))
