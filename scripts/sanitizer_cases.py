"""Small, representative invocations of every kernel family for compute-sanitizer (memcheck / racecheck / synccheck):
    compute-sanitizer --tool racecheck python scripts/sanitizer_cases.py
Every case is checked against the CPU oracle, so a sanitizer-clean run is also a parity run.  Sizes are chosen so that the
forced tile grids, the box march, the lane march (ragged last warp), the VJP kernels, the gates, the rolling diagonal sweep and
the contraction all execute within a few minutes under the sanitizer's ~50x slowdown."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from conftest import random_triple
import oracle
from oracle import gates as og
import mrmustard_b200 as mm
from mrmustard_b200 import strategies as S, fock

def check(name, ok):
    print(f"{name}: {'ok' if ok else 'MISMATCH'}", flush=True)
    assert ok, name

# tiled march (k_march_tiled2): forced tile grids, halo exchange through the sentinel buffer
os.environ["MMH_FORCE_TILED"] = "1"
for shape, grid in [((9, 8, 7, 6), "2,2,2"), ((7, 20, 19), "2,2"), ((5, 6, 5, 4, 3), "2,2,2")]:
    os.environ["MMH_TILE_G"] = grid
    A, b, c = random_triple(len(shape), (), seed=7 + len(shape))
    check(f"tiled {shape} grid {grid}", np.array_equal(S.vanilla_numba(shape, A, b, complex(c)), oracle.vanilla(shape, A, b, complex(c))))
del os.environ["MMH_TILE_G"]
# row-lane march (k_march_rows): forced box grids, compute / service warp hand-offs, halo exchange through the sentinel buffer
for shape, grid, R in [((9, 8, 7, 6), "2,2,2", "2"), ((7, 20, 19), "2,2", "3"), ((3, 13, 12, 11), "3,3,2", "5")]:
    os.environ["MMH_ROWS_G"], os.environ["MMH_ROWS_R"] = grid, R
    A, b, c = random_triple(len(shape), (), seed=11 + len(shape))
    check(f"rows {shape} grid {grid} R {R}", np.array_equal(S.vanilla_numba(shape, A, b, complex(c)), oracle.vanilla(shape, A, b, complex(c))))
del os.environ["MMH_FORCE_TILED"], os.environ["MMH_ROWS_G"], os.environ["MMH_ROWS_R"]
# default single-lattice path with stage overlap (sentinel-validated panel 0) on a lattice large enough for it
shape = (16, 17, 18, 19)
A, b, c = random_triple(4, (), seed=2)
check(f"single lattice {shape}", np.array_equal(S.vanilla_numba(shape, A, b, complex(c)), oracle.vanilla(shape, A, b, complex(c))))
# lane march with fused chain (k_march_lanes), ragged last warp; stage march; box march
for shape, B in [((9, 31), 260), ((3, 4, 40), 260), ((2, 3, 5, 7), 300)]:
    A, b, c = random_triple(len(shape), (B,), seed=29)
    check(f"lanes {shape} x {B}", np.array_equal(S.vanilla_batch_numba(shape, A, b, c), oracle.vanilla_batch(shape, A, b, c)))
shape = (14, 13, 12, 11)
A, b, c = random_triple(4, (40,), seed=17)
check(f"box march {shape} x 40", np.array_equal(S.vanilla_batch_numba(shape, A, b, c), oracle.vanilla_batch(shape, A, b, c)))
# stable rule (cooperative level wavefront) and one-CTA kernels
shape = (9, 8, 7, 6)
A, b, c = random_triple(4, (), seed=5)
check("stable", np.array_equal(S.stable_numba(shape, A, b, complex(c)), oracle.vanilla(shape, A, b, complex(c), stable=True)))
# stable rule on a wavefront of boxes (k_stable_boxes): ticket order, ready flags, ragged boxes in 2 / 3 / 4 indices
os.environ["MMH_STABLE_BOXES_MIN_N"] = "1"
for shp in [(7, 6, 9, 8), (15, 14, 29), (70, 130)]:
    A_, b_, c_ = random_triple(len(shp), (), seed=17 + len(shp))
    check(f"stable boxes {shp}", np.array_equal(S.stable_numba(shp, A_, b_, complex(c_)), oracle.vanilla(shp, A_, b_, complex(c_), stable=True)))
del os.environ["MMH_STABLE_BOXES_MIN_N"]
# VJP on TMA-staged plane tiles (k_vjp_planes): bulk copies + mbarrier transaction counts, several tasks per CTA
os.environ["MMH_VJP_PLANES_MIN_N"], os.environ["MMH_VJP_NSEG"] = "1", "3"
shp = (5, 9, 20, 13)
A_, b_, c_ = random_triple(4, (), seed=41)
G_ = oracle.vanilla(shp, A_, b_, complex(c_))
g_ = np.random.RandomState(21).standard_normal(shp) + 1j * np.random.RandomState(22).standard_normal(shp)
got, want = S.vanilla_vjp_numba(G_, complex(c_), g_), oracle.vanilla_vjp(G_, complex(c_), g_)
check("vjp planes", all(np.allclose(x, y, rtol=1e-10, atol=1e-14) for x, y in zip(got, want)))
del os.environ["MMH_VJP_PLANES_MIN_N"], os.environ["MMH_VJP_NSEG"]
# VJP: partial + finish, and the lane row walk
G = oracle.vanilla(shape, A, b, complex(c))
g = np.random.RandomState(1).standard_normal(shape) + 0j
got, want = S.vanilla_vjp_numba(G, complex(c), g), oracle.vanilla_vjp(G, complex(c), g)
check("vjp", all(np.allclose(x, y, rtol=1e-10, atol=1e-14) for x, y in zip(got, want)))
A2, b2, c2 = random_triple(2, (260,), seed=3)
G2 = oracle.vanilla_batch((9, 31), A2, b2, c2)
g2 = np.random.RandomState(2).standard_normal(G2.shape) + 0j
got, want = S.vanilla_batch_vjp_numba(G2, c2, g2), oracle.vanilla_batch_vjp(G2, c2, g2)
check("vjp lanes", all(np.allclose(x, y, rtol=1e-10, atol=1e-14) for x, y in zip(got, want)))
# gates
check("squeezer", np.array_equal(S.squeezer((12, 9), 0.4, 0.7) + 0.0, og.squeezer((12, 9), 0.4, 0.7) + 0.0))
check("beamsplitter", np.array_equal(S.beamsplitter((5, 4, 6, 3), 0.5, 0.2) + 0.0, og.beamsplitter((5, 4, 6, 3), 0.5, 0.2) + 0.0))
check("stable_beamsplitter", np.array_equal(S.stable_beamsplitter((5, 4, 6, 3), 0.5, 0.2) + 0.0, og.stable_beamsplitter((5, 4, 6, 3), 0.5, 0.2) + 0.0))
check("displacement", np.allclose(S.displacement((9, 12), 0.3 + 0.4j), og.displacement((9, 12), 0.3 + 0.4j), rtol=1e-10, atol=1e-14))
# compactFock: full layout and rolling level buffers
from oracle import diagonal as od
gd = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "diagonal_golden.npz"))
A, b, c = gd["d3_A"], gd["d3_b"], complex(gd["d3_c"])
for roll in ("0", "1"):
    os.environ["MMH_DIAG_ROLLING"] = roll
    check(f"diagonal rolling={roll}", np.allclose(mm.hermite_renormalized_diagonal(A, b, c, (5, 5, 5)), gd["d3_G"], rtol=1e-10, atol=1e-14))
del os.environ["MMH_DIAG_ROLLING"]
# Fock-space contraction
rng = np.random.RandomState(0)
a1 = rng.standard_normal((9, 8, 9, 8)) + 1j * rng.standard_normal((9, 8, 9, 8)); a2 = rng.standard_normal((7, 10)) + 0j
check("contract", np.allclose(fock.contract(a1, [0, 1, 2, 3], a2, [2, 3], [0, 1]), np.einsum("abcd,cd->ab", a1[:, :, :7, :], a2[:, :8]), rtol=1e-12))
print("all cases ok")
