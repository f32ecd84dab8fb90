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


# ---- launch plans of the batched lane / box kernels (host only, through the mmh_debug_plan export) ------------------
def _plan(what, shape, stage=0):
    import ctypes
    from mrmustard_b200 import _lib
    out = (ctypes.c_int * 6)()
    rc = _lib.lib.mmh_debug_plan(what, len(shape), _lib.shape_array(shape), stage, out)
    return rc, list(out)


@pytest.mark.parametrize("n1", [1, 2, 3, 5, 16, 17, 31, 32, 33, 40, 63, 64, 100, 129, 255, 256])
def test_lane_plan_covers_the_row(n1):
    # mmh_lanes.cu: a lane owns R consecutive positions, ln lanes hold a row, a warp marches Lw rows
    rc, (R, ln, Lw, *_) = _plan(0, (7, n1))
    assert rc == 0
    assert 2 <= R <= 8 and ln * R >= n1 > (ln - 1) * R
    assert 1 <= Lw == 32 // ln


def test_lane_plan_rejects_rows_beyond_a_warp():
    assert _plan(0, (7, 257))[0] != 0      # 8 positions x 32 lanes is the widest row
    assert _plan(0, (3, 40))[1][:3] == [5, 8, 4]   # cfg3: five positions per lane, four lattices per warp, no idle lane


@pytest.mark.parametrize("shape", [(20, 20, 20, 20), (30, 30, 30, 30), (50, 50, 50, 50), (64, 64, 64), (5, 40, 41), (4, 1100),
                                   (3, 6, 6, 6, 6), (7, 2, 3, 300), (9, 33, 1, 37), (12, 2000, 3)])
def test_box_plan_invariants(shape):
    # mmh_box.cu: boxes of <= 1024 points over the first <= 3 panel dims, two points per thread, halo <= 2 cells per thread,
    # the shared-memory panel buffer holds box + halo + zero cell + trash cell
    rc, (g0, g1, g2, T, nt, ls) = _plan(1, shape, 0)
    assert rc == 0
    g = [g0, g1, g2]
    panel = shape[1:]
    assert nt == min(3, len(panel))
    inner = int(np.prod(panel[nt:]))
    e = [-(-panel[m] // g[m]) if m < nt else 1 for m in range(3)]
    assert all(g[m] == 1 for m in range(nt, 3)) and all(1 <= g[m] <= panel[m] for m in range(nt))
    TS = inner * e[0] * e[1] * e[2]
    HC = sum(TS // e[m] for m in range(3) if g[m] > 1)
    assert TS <= 1024 and T % 32 == 0 and 64 <= T <= 512 and 2 * T >= TS and HC <= 2 * T
    assert ls == TS + HC + 2
    # the last index of the lattice stays whole whenever a row of it fits a box (write granularity)
    if inner == 1 and panel[nt - 1] <= 1024:
        assert g[nt - 1] == 1


def test_box_plan_limits():
    assert _plan(1, (4, 5, 6, 7, 8, 9, 10), 0)[0] != 0      # more than four panel dims: not boxed
    assert _plan(1, (4, 2000, 2000), 1)[0] == 0             # stage 1: a 1-D panel of 2000 points, two boxes
    assert _plan(1, (4, 2000, 2000), 1)[1][:3] == [2, 1, 1]
    assert _plan(1, (4, 3, 2000), 5)[0] != 0                # bad stage


# ---- single-lattice planners of round 2 (mmh_rows.cu, mmh_stable_boxes.cu), host only ---------------------------------------
def test_row_lane_plan_for_cfg2_and_thresholds():
    # cfg2 stage 0: 125 boxes of 10x10x10 on the row-lane march, two cells per lane, five lanes per row, conflict-free row stride
    rc, (rows, g0, g1, g2, rc_, rs) = _plan(2, (50, 50, 50, 50), 0)
    assert rc == 0 and rows == 1 and (g0, g1, g2) == (5, 5, 5)
    R, C = divmod(rc_, 1000)
    assert (R, C) == (2, 5) and rs >= R * C and rs % 8 == C % 8
    # stage 1 (a 50 x 50 panel) and lattices with boxes below 700 cells stay on the tiled march
    assert _plan(2, (50, 50, 50, 50), 1)[1][0] == 0
    assert _plan(2, (40, 40, 40, 40), 0)[1][0] == 0
    # a panel beyond 148 x 1024 cells has no box plan at all (per-step launches)
    assert _plan(2, (56, 56, 56, 56), 0)[0] != 0


@pytest.mark.parametrize("shape", [(8, 45, 45, 45), (3, 50, 49, 48), (20, 47, 49, 51), (6, 44, 52, 50), (5, 53, 46, 47)])
def test_row_lane_plan_invariants(shape):
    rc, (rows, g0, g1, g2, rc_, rs) = _plan(2, shape, 0)
    assert rc == 0
    if not rows:
        return
    R, C = divmod(rc_, 1000)
    e = [-(-shape[1 + m] // g) for m, g in enumerate((g0, g1, g2))]
    assert g0 * g1 * g2 <= 148 and e[0] * e[1] * e[2] >= 700
    assert C * R >= e[2] > (C - 1) * R                      # the chunks of a row cover it without an idle chunk
    assert e[0] * e[1] * C <= 512                           # compute lanes of one CTA
    assert rs >= R * C and rs % 8 == C % 8                  # row stride continues the bank groups from row to row


@pytest.mark.parametrize("shape", [(50, 50, 50, 50), (100, 100, 100), (1000, 1000), (7, 6, 9, 8), (62, 63)])
def test_stable_box_plan(shape):
    D = len(shape)
    rc, (E, n0, n1, n2, n3, smem) = _plan(3, shape)
    assert rc == 0 and E == {4: 6, 3: 14, 2: 62}[D]
    nb = [n0, n1, n2, n3]
    assert nb[:4 - D] == [1] * (4 - D)
    assert all(nb[4 - D + j] == -(-shape[j] // E) for j in range(D))
    assert (E + 2) ** D == 4096 and smem < 200 * 1024       # the extended box (two-deep lower halo) is 4096 cells = 64 KB
    assert _plan(3, (5,) * 5)[0] != 0                        # five indices: level wavefront
