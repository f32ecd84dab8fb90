"""Pins the diagonal / 1-leftover-mode oracle (oracle/diagonal.py) to the reference's golden vectors (CPU)."""
import numpy as np
import pytest

from conftest import GOLDEN, assert_parity
from oracle import diagonal as od


@pytest.fixture(scope="module")
def gd():
    import os
    return np.load(os.path.join(GOLDEN, "diagonal_golden.npz"))


def test_diagonal_cases(gd):
    for name in gd["diag_cases"]:
        if name == "d3b":
            continue  # (18,19,20) is covered on the GPU; the pure-Python oracle takes too long here
        A, b, c = gd[f"{name}_A"], gd[f"{name}_b"], complex(gd[f"{name}_c"])
        cut = tuple(int(x) for x in gd[f"{name}_cut"])
        A2, b2 = od.reorder_AB_bargmann(A, b)
        G = od.diagonal(A2, b2, c, cut)
        assert_parity(G, gd[f"{name}_G"], name)


def test_diagonal_b_batched(gd):
    A2, b2 = od.reorder_AB_bargmann(gd["db_A"], gd["db_b"])
    G = od.diagonal(A2, b2, complex(gd["db_c"]), tuple(int(x) for x in gd["db_cut"]))
    assert G.shape == gd["db_G"].shape            # batch on the last axis
    assert_parity(G, gd["db_G"], "b-batched diagonal")


def test_leftover_cases(gd):
    for name in gd["leftover_cases"]:
        if name == "l3b":
            continue  # large; GPU only
        A, b, c = gd[f"{name}_A"], gd[f"{name}_b"], complex(gd[f"{name}_c"])
        oc, pnr = int(gd[f"{name}_oc"]), tuple(int(x) for x in gd[f"{name}_pnr"])
        A2, b2 = od.reorder_AB_bargmann(A, b)
        G = od.leftover(A2, b2, c, (oc + 1,) + tuple(p + 1 for p in pnr))
        assert_parity(G, gd[f"{name}_Gcompact"], name + " vs compactFock")
        assert np.allclose(G, gd[f"{name}_G"], rtol=1e-9, atol=1e-12), name + " vs numpy backend (fast_diagonal)"
        F = od.fast_diagonal(A, b, c, oc, pnr)
        assert F.shape == gd[f"{name}_Gfast"].shape
        assert np.allclose(F, gd[f"{name}_Gfast"], rtol=1e-9, atol=1e-12)
