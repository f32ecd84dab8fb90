"""GPU parity tests for the VJP (vanilla_vjp_numba / vanilla_batch_vjp_numba) through the C ABI.

Gate: |x-y| <= 1e-14 + 1e-10 |y| elementwise against the oracle and the reference's golden vectors
(the sums are re-associated on the GPU, so bit-exactness is not expected; BASELINE.json north_star).
"""
import numpy as np
import pytest

from conftest import assert_parity, random_triple

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def S():
    from mrmustard_b200 import strategies
    return strategies


@pytest.fixture(scope="module")
def O():
    import oracle
    return oracle


def _check(got, want, what):
    dA, db, dc = got
    wA, wb, wc = want
    assert_parity(dA, wA, what + " dLdA")
    assert_parity(db, wb, what + " dLdb")
    assert_parity(np.asarray(dc), np.asarray(wc), what + " dLdc")


def test_golden_random_cases(S, golden):
    for name in golden["random_cases"]:
        got = S.vanilla_vjp_numba(golden[f"{name}_G"], complex(golden[f"{name}_c"]), golden[f"{name}_g"])
        assert isinstance(got[2], complex)
        _check(got, (golden[f"{name}_dA"], golden[f"{name}_db"], golden[f"{name}_dc"]), name)
        # structural facts of gradients.py:82: dLdA is symmetric
        assert np.array_equal(got[0], got[0].T)


def test_golden_batch_cases(S, golden):
    for name in golden["batch_cases"]:
        got = S.vanilla_batch_vjp_numba(golden[f"{name}_G"], golden[f"{name}_c"], golden[f"{name}_g"])
        _check(got, (golden[f"{name}_dA"], golden[f"{name}_db"], golden[f"{name}_dc"]), name)


@pytest.mark.parametrize("n", [2, 3])
def test_finite_differences(S, n):
    """The reference's own VJP test (tests/test_math/test_lattice/test_vanilla.py:43-87), on the GPU path."""
    eps = 1e-9
    A, b, c = random_triple(n, (), seed=673)
    shape = (4,) * n
    G = S.vanilla_numba(shape, A, b, complex(c))
    dLdG = np.random.RandomState(7).standard_normal(G.shape)
    dLdA, dLdb, dLdc = S.vanilla_vjp_numba(G, complex(c), dLdG + 0j)
    fd_c = np.sum(dLdG * (S.vanilla_numba(shape, A, b, complex(c) + eps) - G) / eps)
    assert np.allclose(dLdc, fd_c)
    fd_b = np.zeros(n, complex)
    for i in range(n):
        bp = b.copy(); bp[i] += eps
        fd_b[i] = np.sum(dLdG * (S.vanilla_numba(shape, A, bp, complex(c)) - G) / eps)
    assert np.allclose(dLdb, fd_b)
    fd_A = np.zeros((n, n), complex)
    for i in range(n):
        for j in range(n):
            Ap = A.copy(); Ap[i, j] += eps
            fd_A[i, j] = np.sum(dLdG * (S.vanilla_numba(shape, Ap, b, complex(c)) - G) / eps)
    assert np.allclose(dLdA, (fd_A + fd_A.T) / 2)


@pytest.mark.parametrize("shape", [(1,), (5,), (1, 2, 3), (3, 1, 1), (7, 9), (6, 5, 4, 3), (3,) * 7, (2,) * 9, (2,) * 10,
                                   (64, 65), (20, 5, 5, 20)])
def test_vs_oracle_shapes(S, O, shape):
    A, b, c = random_triple(len(shape), (), seed=3 + len(shape))
    G = O.vanilla(shape, A, b, complex(c))
    g = np.random.RandomState(11).standard_normal(shape) + 1j * np.random.RandomState(12).standard_normal(shape)
    _check(S.vanilla_vjp_numba(G, complex(c), g), O.vanilla_vjp(G, complex(c), g), str(shape))


def test_cfg5_full_size(S, O, golden):
    A, b, c = golden["cfg5_A"], golden["cfg5_b"], complex(golden["cfg5_c"])
    G8 = golden["cfg5_G8"]
    _check(S.vanilla_vjp_numba(G8, c, golden["cfg5_g8"]), (golden["cfg5_dA8"], golden["cfg5_db8"], golden["cfg5_dc8"]), "cfg5 (8,)*4")
    G = S.vanilla_numba((40,) * 4, A, b, c)
    g = np.random.RandomState(1).standard_normal(G.shape) + 0j
    _check(S.vanilla_vjp_numba(G, c, g), (golden["cfg5_dA40"], golden["cfg5_db40"], golden["cfg5_dc40"]), "cfg5 (40,)*4")


def test_cfg3_batched(S, O, golden):
    A, b, c = random_triple(2, (65536,), seed=673)
    A, b, c = A[:64].copy(), b[:64].copy(), c[:64].copy()
    G = S.vanilla_batch_numba((40, 40), A, b, c)
    g = np.random.RandomState(1).standard_normal((64, 40, 40)) + 0j
    _check(S.vanilla_batch_vjp_numba(G, c, g), (golden["cfg3_dA64"], golden["cfg3_db64"], golden["cfg3_dc64"]), "cfg3[:64]")
    # a larger slice against the oracle, exercising grid.x = batch > 65535/.. paths
    A, b, c = random_triple(2, (70000,), seed=5)
    G = O.vanilla_batch((6, 7), A, b, c)
    g = np.random.RandomState(2).standard_normal(G.shape) + 1j * np.random.RandomState(3).standard_normal(G.shape)
    _check(S.vanilla_batch_vjp_numba(G, c, g), O.vanilla_batch_vjp(G, c, g), "B=70000")


@pytest.mark.parametrize("shape", [(40, 40), (5, 2), (9, 31), (6, 33), (4, 100), (3, 256), (3, 257), (2, 17), (17, 3)])
def test_row_walk_batched_2d(S, O, shape):
    # batches >= 256 of 2-index lattices go through the warp-synchronous row walk (k_vjp_lanes); 260 leaves a ragged last warp
    A, b, c = random_triple(2, (260,), seed=31)
    G = O.vanilla_batch(shape, A, b, c)
    g = np.random.RandomState(4).standard_normal(G.shape) + 1j * np.random.RandomState(5).standard_normal(G.shape)
    _check(S.vanilla_batch_vjp_numba(G, c, g), O.vanilla_batch_vjp(G, c, g), f"row walk {shape}")


def test_linearity_in_cotangent(S):
    """Size-independent property: the VJP is linear in dLdG."""
    A, b, c = random_triple(3, (), seed=21)
    G = S.vanilla_numba((16, 15, 14), A, b, complex(c))
    r = np.random.RandomState(4)
    g1 = r.standard_normal(G.shape) + 1j * r.standard_normal(G.shape)
    g2 = r.standard_normal(G.shape) + 1j * r.standard_normal(G.shape)
    a1 = S.vanilla_vjp_numba(G, complex(c), g1)
    a2 = S.vanilla_vjp_numba(G, complex(c), g2)
    a12 = S.vanilla_vjp_numba(G, complex(c), 2.0 * g1 - 3.0 * g2)
    for x1, x2, x12 in zip(a1, a2, a12):
        assert np.allclose(2.0 * np.asarray(x1) - 3.0 * np.asarray(x2), np.asarray(x12), rtol=1e-10, atol=1e-12)


@pytest.mark.parametrize("shape,batch", [((40, 40, 40), 1), ((9, 10, 33, 41), 1), ((5, 4, 6, 7, 37), 2), ((3, 2, 3, 2, 20, 70), 1),
                                         ((33, 33, 5, 4), 3), ((12, 12, 12, 12), 1), ((2, 1, 130, 129), 1)])
def test_more_shapes_vs_oracle(shape, batch):
    """Lattices with three to six indices, ragged last dims, small batches: the VJP against the oracle."""
    import oracle
    from mrmustard_b200 import strategies as S
    D = len(shape)
    rng = np.random.RandomState(D * 7 + batch)
    if batch == 1:
        A, b, c = random_triple(D, (), seed=50 + D)
        G = oracle.vanilla(shape, A * 0.6, b * 0.7, complex(c))
        g = rng.standard_normal(shape) + 1j * rng.standard_normal(shape)
        got, want = S.vanilla_vjp_numba(G, complex(c), g), oracle.vanilla_vjp(G, complex(c), g)
    else:
        A, b, c = random_triple(D, (batch,), seed=60 + D)
        G = oracle.vanilla_batch(shape, A * 0.6, b * 0.7, c)
        g = rng.standard_normal(G.shape) + 1j * rng.standard_normal(G.shape)
        got, want = S.vanilla_batch_vjp_numba(G, c, g), oracle.vanilla_batch_vjp(G, c, g)
    for x, y, nm in zip(got, want, ("dLdA", "dLdb", "dLdc")):
        assert_parity(np.asarray(x, dtype=np.complex128), np.asarray(y, dtype=np.complex128), f"{shape} {nm} vs oracle")


@pytest.mark.parametrize("shape,nseg", [((6, 7, 16, 17), None), ((5, 9, 20, 13), "3"), ((4, 4, 33, 31), "4"), ((9, 5, 16, 16), "1"),
                                         ((12, 11, 19, 23), "5")])
def test_plane_tile_vjp_vs_oracle(S, O, monkeypatch, shape, nseg):
    """k_vjp_planes (TMA-staged plane tiles of one 4-index lattice, the cfg5 kernel) forced onto small lattices: every plane role
    (k0 - 1, k0 - 2, k1 - 1, k1 - 2 absent at the low edges), ragged k1 segments, several tasks per CTA; gate 1e-10 / 1e-14."""
    from mrmustard_b200 import _lib
    monkeypatch.setenv("MMH_VJP_PLANES_MIN_N", "1")
    if nseg:
        monkeypatch.setenv("MMH_VJP_NSEG", nseg)
    A, b, c = random_triple(4, (), seed=41)
    G = O.vanilla(shape, A, b, complex(c))
    g = np.random.RandomState(21).standard_normal(shape) + 1j * np.random.RandomState(22).standard_normal(shape)
    n0 = _lib.launch_count()
    got = S.vanilla_vjp_numba(G, complex(c), g)
    assert _lib.launch_count() == n0 + 2
    _check(got, O.vanilla_vjp(G, complex(c), g), f"planes {shape}")
    monkeypatch.setenv("MMH_NO_VJP_PLANES", "1")
    ref = S.vanilla_vjp_numba(G, complex(c), g)
    for x, y in zip(got, ref):
        assert np.allclose(np.asarray(x), np.asarray(y), rtol=1e-11, atol=1e-14)
