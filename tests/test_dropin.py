"""CPU test of the drop-in install: with /root/reference present (build container only; the GPU
box has no reference tree), `dropin.install()` is applied to the LIVE unmodified reference and its own callers
(CircuitComponent.fock_array, math.hermite_renormalized*) are driven through our operator mirror.  The GPU strategies are replaced
BY THE TEST with the oracle (no device here), so what is checked is the wiring: same signatures, same results, calls really routed."""
import numpy as np
import pytest

import oracle
from oracle import refimport

pytestmark = pytest.mark.skipif(not refimport.available(), reason="reference tree not present (GPU box)")


@pytest.fixture()
def routed(monkeypatch):
    refimport.install_shims(with_lab=True)
    from mrmustard_b200 import dropin, strategies
    calls = {"vanilla": 0, "batch": 0}

    def vanilla(shape, A, b, c, out=None):
        calls["vanilla"] += 1
        return oracle.vanilla(shape, A, b, c, out=out)

    def batch(shape, A, b, c, stable=False, out=None):
        calls["batch"] += 1
        return oracle.vanilla_batch(shape, A, b, c, stable, out)

    monkeypatch.setattr(strategies, "vanilla_numba", vanilla)
    monkeypatch.setattr(strategies, "vanilla_batch_numba", batch)
    yield dropin, calls
    dropin.uninstall()


def test_fock_array_routes_through_dropin(routed):
    dropin, calls = routed
    from mrmustard.lab import BSgate, Sgate
    u = BSgate((0, 1), theta=0.5, phi=0.2) >> Sgate(0, r=0.3) >> Sgate(1, r=0.2)
    want = np.asarray(u.fock_array((6, 6, 6, 6)))            # stock numba path
    dropin.install()
    got = np.asarray(u.fock_array((6, 6, 6, 6)))
    assert calls["vanilla"] == 1
    assert got.shape == want.shape and np.array_equal(got, want)
    dropin.uninstall()
    assert np.array_equal(np.asarray(u.fock_array((6, 6, 6, 6))), want) and calls["vanilla"] == 1   # restored


def test_manager_batched_entry_routes_through_dropin(routed):
    dropin, calls = routed
    from mrmustard import math
    rng = np.random.RandomState(4)
    A = rng.random((3, 2, 2)) + 1j * rng.random((3, 2, 2)); A = (A + A.transpose(0, 2, 1)) / 4
    b = rng.random((3, 2)) + 1j * rng.random((3, 2)); c = rng.random(3) + 0j
    want = np.asarray(math.hermite_renormalized(A, b, c, (5, 4)))
    dropin.install()
    got = np.asarray(math.hermite_renormalized(A, b, c, (5, 4)))
    assert calls["batch"] == 1 and np.array_equal(got, want)
    out = np.zeros((3, 5, 4), dtype=np.complex128)
    res = math.hermite_renormalized(A, b, c, (5, 4), out=out)      # the manager returns a reshaped view of `out` when batched
    assert np.shares_memory(res, out) and np.array_equal(out, want)
