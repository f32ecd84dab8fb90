"""Lattice + derived-variable contraction: oracle vs the reference's golden vectors (CPU) and
the CUDA path vs both (GPU).  Tolerance: the north_star gate |x - y| <= 1e-14 + 1e-10 |y| (the reference's einsum order is
not pinned)."""
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def gc():
    return np.load(os.path.join(HERE, "golden", "contract_golden.npz"))


def close(x, y):
    return np.all(np.abs(x - y) <= 1e-14 + 1e-10 * np.abs(y))


def test_oracle_contract_vs_reference_golden(gc):
    import oracle
    for name in gc["cases"]:
        A, b, c, F = gc[f"{name}_A"], gc[f"{name}_b"], gc[f"{name}_c"], gc[f"{name}_F"]
        core, der = tuple(gc[f"{name}_core"]), tuple(gc[f"{name}_der"])
        batch = b.shape[:-1]
        if not batch:
            assert close(oracle.vanilla_contract(core, der, A, b, c), F), name
        else:
            for idx in np.ndindex(*batch):
                assert close(oracle.vanilla_contract(core, der, A[idx], b[idx], c[idx]), F[idx]), (name, idx)


@pytest.mark.gpu
def test_gpu_contract_vs_reference_golden(gc):
    import mrmustard_b200 as mm
    for name in gc["cases"]:
        A, b, c, F = gc[f"{name}_A"], gc[f"{name}_b"], gc[f"{name}_c"], gc[f"{name}_F"]
        got = mm.hermite_renormalized_contracted(A, b, c, tuple(gc[f"{name}_core"]))
        assert got.shape == F.shape and got.dtype == np.complex128
        assert close(got, F), name


@pytest.mark.gpu
def test_gpu_contract_vs_materialised_lattice():
    """Size-independent property at a larger size: contracting on the device equals contracting the materialised lattice."""
    import mrmustard_b200 as mm
    from conftest import random_triple
    A, b, _ = random_triple(4, (), seed=21)
    rng = np.random.RandomState(3)
    core, der = (30, 31), (6, 50)
    c = rng.standard_normal(der) + 1j * rng.standard_normal(der)
    G = mm.strategies.vanilla_numba(core + der, A, b, 1.0)
    want = np.einsum("abk,k->ab", G.reshape(core + (-1,)), c.reshape(-1))
    assert close(mm.hermite_renormalized_contracted(A, b, c, core), want)
    Ab, bb = np.stack([A, 0.7 * A, A]), np.stack([b, b, 0.5 * b])
    cb = np.stack([c, 2 * c, c])
    got = mm.hermite_renormalized_contracted(Ab, bb, cb, core)
    assert close(got[0], want) and close(got[2], np.einsum("abk,k->ab", mm.strategies.vanilla_numba(core + der, A, 0.5 * b, 1.0).reshape(core + (-1,)), c.reshape(-1)))


@pytest.mark.gpu
def test_gpu_contract_errors():
    import mrmustard_b200 as mm
    A = np.eye(3) * 0.1; b = np.ones(3) * 0.1
    with pytest.raises(ValueError):
        mm.hermite_renormalized_contracted(A, b, np.ones((2, 2, 2, 2)), (3,))   # too many derived axes
    with pytest.raises(ValueError):
        mm.hermite_renormalized_contracted(A, b, np.ones((2,)), (3, 3, 3, 3))   # more core axes than variables
