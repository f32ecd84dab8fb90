"""CPU tests of the host-side mirror (mrmustard_b200.backend): batching/broadcast/reshape semantics of
BackendManager.hermite_renormalized (backend_manager.py:643-727).  The GPU strategies are replaced BY THE
TEST with the oracle (test-only injection) so the host logic can run without a device."""
import numpy as np
import pytest

import oracle
from conftest import random_triple


@pytest.fixture()
def backend(monkeypatch):
    from mrmustard_b200 import backend as be, strategies
    monkeypatch.setattr(strategies, "vanilla_numba", lambda shape, A, b, c, out=None: oracle.vanilla(shape, A, b, c, out=out))
    monkeypatch.setattr(strategies, "stable_numba", lambda shape, A, b, c, out=None: oracle.stable(shape, A, b, c, out=out))
    monkeypatch.setattr(strategies, "vanilla_batch_numba",
                        lambda shape, A, b, c, stable=False, out=None: oracle.vanilla_batch(shape, A, b, c, stable, out))
    monkeypatch.setattr(strategies, "binomial", oracle.binomial)
    return be


@pytest.mark.parametrize("stable", [True, False])
def test_manager_semantics(backend, stable):
    A, b, c = random_triple(2, (), seed=673)
    G = backend.hermite_renormalized(A, b, c, (3, 3), stable=stable)
    assert G.shape == (3, 3)
    out_arr = np.zeros((3, 3), dtype=np.complex128)
    assert backend.hermite_renormalized(A, b, c, (3, 3), stable=stable, out=out_arr) is out_arr
    A, b, c = random_triple(2, (2, 1), seed=673)
    shape = (4, 5)
    G = backend.hermite_renormalized(A[0, 0], b, c[0, 0], shape, stable=stable)
    assert G.shape == (2, 1, *shape)
    assert np.array_equal(G[1, 0], oracle.vanilla(shape, A[0, 0], b[1, 0], complex(c[0, 0]), stable=stable))
    G = backend.hermite_renormalized(A, b, c, shape, stable=stable)
    assert G.shape == (2, 1, *shape)
    assert np.array_equal(G[1, 0], oracle.vanilla(shape, A[1, 0], b[1, 0], complex(c[1, 0]), stable=stable))
    out_arr = np.zeros((2, 1, *shape), dtype=np.complex128)
    G = backend.hermite_renormalized(A, b, c, shape, stable=stable, out=out_arr)
    assert np.array_equal(out_arr[0, 0], oracle.vanilla(shape, A[0, 0], b[0, 0], complex(c[0, 0]), stable=stable))


def test_manager_errors(backend):
    A, b, c = random_triple(2, (2, 1), seed=673)
    with pytest.raises(ValueError):
        backend.hermite_renormalized(A, b, c, (4, 5), out=np.zeros((2, 1, 3, 5), complex))
    with pytest.raises(ValueError):
        backend.hermite_renormalized(A, b[:1], c, (4, 5))
    with pytest.raises(ValueError):
        backend.hermite_renormalized(A, b, c[:1], (4, 5))


def test_stable_setting_is_honoured(backend):
    A, b, c = random_triple(2, (), seed=1)
    backend.settings.STABLE_FOCK_CONVERSION = True
    try:
        G = backend.hermite_renormalized(A, b, c, (6, 6))
    finally:
        backend.settings.STABLE_FOCK_CONVERSION = False
    assert np.array_equal(G, oracle.stable((6, 6), A, b, complex(c)))


def test_binomial_defaults(backend, golden):
    # backend_numpy.py:419-420: max_l2 or AUTOSHAPE_PROBABILITY ; global_cutoff or sum(shape)-len(shape)+1
    A, b, c = golden["bin_A"], golden["bin_b"], complex(golden["bin_c"])
    G = backend.hermite_renormalized_binomial(A, b, c, (6, 6), None, None)
    want = oracle.binomial((6, 6), A, b, c, 0.99999, 11)[0]
    assert np.array_equal(G, want)
