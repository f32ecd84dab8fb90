"""Pins the C oracle (oracle/hermite_oracle.c) to the reference's own output (CPU, no GPU needed).

Golden vectors come from the unmodified numba strategies (tests/golden/gen_golden.py); the known-answer
test restates the reference's tests/test_math/test_special.py:24-36.
"""
import numpy as np
import pytest
from scipy.special import eval_hermite, factorial

import oracle
from conftest import assert_parity, random_triple, sha


def test_kat_hermite_polynomials():
    # reference: tests/test_math/test_special.py:24-36
    x = np.arange(-1, 1, 0.1)
    A = -np.ones((1, 1), dtype=complex)
    vals = np.array([oracle.vanilla((5,), 2 * A, 2 * np.array([x0], dtype=complex), 1) for x0 in x]).T
    expected = np.array([eval_hermite(i, x) / np.sqrt(factorial(i)) for i in range(5)])
    assert np.allclose(vals, expected)
    vals_s = np.array([oracle.stable((5,), 2 * A, 2 * np.array([x0], dtype=complex), 1) for x0 in x]).T
    assert np.allclose(vals_s, expected)


def test_random_cases_bit_exact(golden):
    for name in golden["random_cases"]:
        A, b, c = golden[f"{name}_A"], golden[f"{name}_b"], complex(golden[f"{name}_c"])
        shape = tuple(golden[f"{name}_shape"])
        G = oracle.vanilla(shape, A, b, c)
        assert G.shape == shape and G.dtype == np.complex128
        assert np.array_equal(G, golden[f"{name}_G"]), name
        assert np.array_equal(oracle.stable(shape, A, b, c), golden[f"{name}_Gs"]), name
        dA, db, dc = oracle.vanilla_vjp(golden[f"{name}_G"], c, golden[f"{name}_g"])
        assert_parity(dA, golden[f"{name}_dA"], name + " dA")
        assert_parity(db, golden[f"{name}_db"], name + " db")
        assert_parity(np.asarray(dc), golden[f"{name}_dc"], name + " dc")


def test_batch_cases(golden):
    for name in golden["batch_cases"]:
        A, b, c = golden[f"{name}_A"], golden[f"{name}_b"], golden[f"{name}_c"]
        shape = tuple(golden[f"{name}_shape"])
        assert np.array_equal(oracle.vanilla_batch(shape, A, b, c), golden[f"{name}_G"]), name
        assert np.array_equal(oracle.vanilla_batch(shape, A, b, c, stable=True), golden[f"{name}_Gs"]), name
        dA, db, dc = oracle.vanilla_batch_vjp(golden[f"{name}_G"], c, golden[f"{name}_g"])
        assert_parity(dA, golden[f"{name}_dA"], name + " dA")
        assert_parity(db, golden[f"{name}_db"], name + " db")
        assert_parity(dc, golden[f"{name}_dc"], name + " dc")


def test_out_is_written_in_place():
    # reference contract: core.py:73 `out.ravel()`; tests/test_math/test_lattice/test_vanilla.py:185
    A, b, c = random_triple(2, (), seed=3)
    out = np.full((4, 5), 7 + 7j)
    ret = oracle.vanilla((4, 5), A, b, complex(c), out=out)
    assert ret is out
    assert np.array_equal(out, oracle.vanilla((4, 5), A, b, complex(c)))


def test_cfg1(golden):
    G = oracle.vanilla((200,), golden["cfg1_A"], golden["cfg1_b"], complex(golden["cfg1_c"]))
    assert np.array_equal(G, golden["cfg1_G"])
    Gs = oracle.stable((200,), golden["cfg1_A"], golden["cfg1_b"], complex(golden["cfg1_c"]))
    assert np.array_equal(Gs, golden["cfg1_G_stable"])


def test_cfg2_small_and_full(golden):
    A, b, c = golden["cfg2_A"], golden["cfg2_b"], complex(golden["cfg2_c"])
    assert np.array_equal(oracle.vanilla((12,) * 4, A, b, c), golden["cfg2_G12"])
    assert np.array_equal(oracle.stable((12,) * 4, A, b, c), golden["cfg2_G12_stable"])
    G = oracle.vanilla((50,) * 4, A, b, c)
    assert sha(G) == str(golden["cfg2_G50_sha"])       # bit-exact on the ill-conditioned config
    assert np.array_equal(G.ravel()[::9973], golden["cfg2_G50_sample"])
    Gs = oracle.stable((50,) * 4, A, b, c)
    assert sha(Gs) == str(golden["cfg2_G50_stable_sha"])
    G = oracle.vanilla((50,) * 4, golden["cfg2r_A"], golden["cfg2r_b"], complex(golden["cfg2r_c"]))
    assert sha(G) == str(golden["cfg2r_G50_sha"])


def test_cfg5_forward_and_vjp(golden):
    A, b, c = golden["cfg5_A"], golden["cfg5_b"], complex(golden["cfg5_c"])
    G8 = oracle.vanilla((8,) * 4, A, b, c)
    assert np.array_equal(G8, golden["cfg5_G8"])
    dA, db, dc = oracle.vanilla_vjp(G8, c, golden["cfg5_g8"])
    assert_parity(dA, golden["cfg5_dA8"]); assert_parity(db, golden["cfg5_db8"])
    assert_parity(np.asarray(dc), golden["cfg5_dc8"])
    G = oracle.vanilla((40,) * 4, A, b, c)
    assert sha(G) == str(golden["cfg5_G40_sha"])
    g = np.random.RandomState(1).standard_normal(G.shape) + 0j
    dA, db, dc = oracle.vanilla_vjp(G, c, g)
    assert_parity(dA, golden["cfg5_dA40"]); assert_parity(db, golden["cfg5_db40"])
    assert_parity(np.asarray(dc), golden["cfg5_dc40"])


def test_cfg4_small(golden):
    G = oracle.vanilla((3,) * 8, golden["cfg4_A"], golden["cfg4_b"], complex(golden["cfg4_c"]))
    assert np.array_equal(G, golden["cfg4_G3"])


def test_cfg3_full_batch(golden):
    A, b, c = random_triple(2, (65536,), seed=673)
    assert sha(np.concatenate([A.ravel(), b.ravel(), c.ravel()])) == str(golden["cfg3_in_sha"])
    G = oracle.vanilla_batch((40, 40), A, b, c)
    assert np.array_equal(G[:4], golden["cfg3_G_first4"])
    assert np.array_equal(G.ravel()[::1000003], golden["cfg3_G_sample"])
    for i, h in enumerate(golden["cfg3_chunk_sha"]):
        assert sha(G[i * 4096:(i + 1) * 4096]) == str(h)
    g = np.random.RandomState(1).standard_normal((64, 40, 40)) + 0j
    dA, db, dc = oracle.vanilla_batch_vjp(G[:64].copy(), c[:64].copy(), g)
    assert_parity(dA, golden["cfg3_dA64"]); assert_parity(db, golden["cfg3_db64"]); assert_parity(dc, golden["cfg3_dc64"])
    Gs = oracle.vanilla_batch((40, 40), A[:64].copy(), b[:64].copy(), c[:64].copy(), stable=True)
    assert sha(Gs) == str(golden["cfg3_Gs64_sha"])


@pytest.mark.parametrize("tag", ["a", "b", "c", "d"])
def test_binomial(golden, tag):
    c0, c1, max_l2, gc = golden[f"bin_{tag}_args"]
    G, norm = oracle.binomial((int(c0), int(c1)), golden["bin_A"], golden["bin_b"], complex(golden["bin_c"]), max_l2, int(gc))
    assert np.array_equal(G, golden[f"bin_{tag}_G"])
    assert norm == float(golden[f"bin_{tag}_norm"])


def test_binomial_3d(golden):
    G, norm = oracle.binomial((4, 3, 5), golden["bin3_A"], golden["bin3_b"], complex(golden["bin3_c"]), 1e9, 10)
    assert np.array_equal(G, golden["bin3_G"])
    assert norm == float(golden["bin3_norm"])


@pytest.mark.parametrize("tag", ["dg", "sg"])
def test_stable_cutoff_1000(golden, tag):
    # reference test_vanilla_stable (tests/test_math/test_lattice/test_lattice_functions.py:137-149)
    A, b, c = golden[f"st_{tag}_A"], golden[f"st_{tag}_b"], complex(golden[f"st_{tag}_c"])
    G = oracle.stable((1000, 1000), A, b, c)
    assert sha(G) == str(golden[f"st_{tag}_sha"])
    if tag == "dg":
        assert np.allclose(G.ravel()[::7919], golden["st_dg_closed_form_sample"])
    else:
        assert np.max(np.abs(G)) < 1
