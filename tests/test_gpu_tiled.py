"""GPU parity tests of the tiled multi-CTA march (k_march_tiled): one lattice cut into tile-owner CTAs that
exchange one-cell halos through L2 with release/acquire progress counters.  Small lattices are forced through
that path (MMH_FORCE_TILED, MMH_TILE_G are test hooks read by the library at call time) with many tile-grid
shapes, and must be bit-identical to the oracle."""
import numpy as np
import pytest

from conftest import random_triple, sha

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def S():
    from mrmustard_b200 import strategies
    return strategies


@pytest.fixture(scope="module")
def O():
    import oracle
    return oracle


CASES = [
    ((9, 8, 7, 6), ["1,1,1", "2,2,2", "3,1,2", "1,1,3", "2,4,1", "4,3,3"]),
    ((7, 20, 19), ["1,1", "2,2", "5,1", "1,6", "7,7"]),
    ((12, 33), ["1", "2", "5", "11"]),
    ((5, 6, 5, 4, 3), ["2,2,2", "3,1,1", "1,2,3"]),
    ((3, 4, 3, 2, 3, 2), ["2,2,1", "1,3,2"]),
    ((2, 3, 2, 3, 2, 3, 2, 3), ["2,1,2"]),
    ((6, 1, 5, 1, 4), ["1,2,1", "1,5,1"]),
]


@pytest.mark.parametrize("shape,grids", CASES)
def test_forced_tiled_vs_oracle(S, O, monkeypatch, shape, grids):
    A, b, c = random_triple(len(shape), (), seed=7 + len(shape))
    want = O.vanilla(shape, A, b, complex(c))
    monkeypatch.setenv("MMH_FORCE_TILED", "1")
    for g in grids:
        monkeypatch.setenv("MMH_TILE_G", g)
        got = S.vanilla_numba(shape, A, b, complex(c))
        assert np.array_equal(got, want), f"shape {shape} tile grid {g}"
    monkeypatch.delenv("MMH_TILE_G")
    assert np.array_equal(S.vanilla_numba(shape, A, b, complex(c)), want)   # planner's own choice


def test_default_path_large_lattices(S, O, golden):
    # default planner on lattices large enough for the multi-CTA path
    for shape, seed in [((40, 41, 42), 1), ((24, 25, 26, 27), 2), ((300, 300), 3), ((8,) * 6, 4), ((50, 3000), 5)]:
        A, b, c = random_triple(len(shape), (), seed=seed)
        assert np.array_equal(S.vanilla_numba(shape, A, b, complex(c)), O.vanilla(shape, A, b, complex(c))), shape
    A, b, c = golden["cfg2_A"], golden["cfg2_b"], complex(golden["cfg2_c"])
    for _ in range(3):   # repeated launches reuse/rotate the flag words
        assert sha(S.vanilla_numba((50,) * 4, A, b, c)) == str(golden["cfg2_G50_sha"])


def test_giant_panel_fallback(S, O):
    # stage-0 panel of 1.77 M points: too large for the tiled march -> per-step launches (k_panel_step)
    shape = (3, 11, 11, 11, 11, 11, 11)
    A, b, c = random_triple(len(shape), (), seed=9)
    assert np.array_equal(S.vanilla_numba(shape, A, b, complex(c)), O.vanilla(shape, A, b, complex(c)))


def test_batch_of_large_lattices_pipelined(S, O):
    """A batch of lattices that each take the multi-kernel single-lattice path: consecutive lattices are pipelined behind each other
    (programmatic launch chain, rotating exchange-buffer slots, full stream order every 4th lattice) -- 6 lattices cross that boundary."""
    shape = (24, 25, 26, 27)
    A, b, c = random_triple(4, (6,), seed=31)
    G = S.vanilla_batch_numba(shape, A, b, c)
    for l in range(6):
        assert np.array_equal(G[l], O.vanilla(shape, A[l], b[l], complex(c[l]))), l
    G2 = S.vanilla_batch_numba(shape, A, b, c)   # second call reuses the slots
    assert np.array_equal(G, G2)


ROWS_CASES = [
    ((9, 8, 7, 6), ["1,1,1", "2,2,2", "3,1,2", "1,1,3", "2,4,1", "4,3,3"]),   # stage 0: three panel dims, stage 1: two
    ((7, 20, 19), ["1,1", "2,2", "5,1", "1,6", "7,7"]),                         # stage 0: two panel dims
    ((5, 6, 5, 4, 3), ["2,2,2", "3,1,1", "1,2,3"]),                             # stages 1 and 2 (stage 0 has four panel dims: tiled2)
    ((3, 13, 12, 11), ["3,3,2", "2,1,4"]),                                      # ragged boxes, rows longer than one chunk
]


@pytest.mark.parametrize("shape,grids", ROWS_CASES)
def test_forced_row_lane_march_vs_oracle(S, O, monkeypatch, shape, grids):
    """k_march_rows (mmh_rows.cu: row-owning compute lanes, service warps for halo import / drain) forced onto small lattices with
    many box grids and every instantiated cells-per-lane count: bit-identical to the oracle."""
    from mrmustard_b200 import _lib
    A, b, c = random_triple(len(shape), (), seed=11 + len(shape))
    want = O.vanilla(shape, A, b, complex(c))
    monkeypatch.setenv("MMH_FORCE_TILED", "1")
    for g in grids:
        monkeypatch.setenv("MMH_ROWS_G", g)
        for R in ("2", "3", "4", "5", "6"):
            monkeypatch.setenv("MMH_ROWS_R", R)
            n0 = _lib.launch_count()
            got = S.vanilla_numba(shape, A, b, complex(c))
            assert _lib.launch_count() > n0
            assert np.array_equal(got, want), f"shape {shape} box grid {g} R {R}"


def test_row_lane_march_default_and_reuse(S, O, golden, monkeypatch):
    """The planner's own choice on lattices whose boxes are large enough for k_march_rows, repeated launches on the same exchange
    buffer (self-cleaning sentinel), against the golden cfg2 lattice and against k_march_tiled2 (MMH_NO_ROWS)."""
    A, b, c = golden["cfg2_A"], golden["cfg2_b"], complex(golden["cfg2_c"])
    for _ in range(3):
        assert sha(S.vanilla_numba((50,) * 4, A, b, c)) == str(golden["cfg2_G50_sha"])
    for shape, seed in [((20, 47, 49, 51), 5), ((12, 60, 33, 40), 6)]:
        A, b, c = random_triple(4, (), seed=seed)
        got = S.vanilla_numba(shape, A, b, complex(c))
        monkeypatch.setenv("MMH_NO_ROWS", "1")
        ref = S.vanilla_numba(shape, A, b, complex(c))
        monkeypatch.delenv("MMH_NO_ROWS")
        assert np.array_equal(got.view(np.int64), ref.view(np.int64)), shape
        # the recurrence only reads lower indices, so the corner of the large lattice is the small lattice of the same triple
        assert np.array_equal(got[:3, :9, :9, :9], O.vanilla((3, 9, 9, 9), A, b, complex(c))), shape


def test_cluster_variant_of_the_tiled_march(S, O, golden, monkeypatch):
    """MMH_CLUSTER=1: tile grids of <= 16 tiles run as one thread-block cluster and push their halo cells into the consumer tile's
    shared memory (st.shared::cluster) instead of through the L2 exchange buffer; same sentinel protocol, bit-identical results."""
    monkeypatch.setenv("MMH_CLUSTER", "1")
    A, b, c = golden["cfg2_A"], golden["cfg2_b"], complex(golden["cfg2_c"])
    for _ in range(2):
        assert sha(S.vanilla_numba((50,) * 4, A, b, c)) == str(golden["cfg2_G50_sha"])   # stage 1: 12 tiles in one cluster
    monkeypatch.setenv("MMH_FORCE_TILED", "1")
    for shape, grids in [((9, 8, 7, 6), ["2,2,2", "2,4,1", "4,2,2"]), ((7, 20, 19), ["2,2", "4,4", "1,6"]), ((12, 33), ["2", "11"])]:
        A, b, c = random_triple(len(shape), (), seed=7 + len(shape))
        want = O.vanilla(shape, A, b, complex(c))
        for g in grids:
            monkeypatch.setenv("MMH_TILE_G", g)
            assert np.array_equal(S.vanilla_numba(shape, A, b, complex(c)), want), f"shape {shape} tile grid {g}"
