"""GPU parity tests for the forward path (vanilla / stable / batched / binomial) through the C ABI.

Every comparison is against the CPU oracle (oracle/, pinned to the reference by tests/test_oracle_golden.py)
on identical input bytes, and against the committed golden vectors of the reference itself.  The forward
results are required to be BIT-IDENTICAL (np.array_equal); the north_star tolerance (1e-10 rel / 1e-14 abs)
is implied.  SURVEY.md H1 explains why bit-exactness is the practical gate on cfg2.
"""
import ctypes

import numpy as np
import pytest
from scipy.special import eval_hermite, factorial

from conftest import assert_parity, random_triple, sha

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def S():
    from mrmustard_b200 import strategies
    return strategies


@pytest.fixture(scope="module")
def O():
    import oracle
    return oracle


def test_kat_hermite_polynomials(S):
    # reference known-answer test: tests/test_math/test_special.py:24-36
    x = np.arange(-1, 1, 0.1)
    A = -np.ones((1, 1), dtype=complex)
    for fn in (S.vanilla_numba, S.stable_numba):
        vals = np.array([fn((5,), 2 * A, 2 * np.array([x0], dtype=complex), 1) for x0 in x]).T
        expected = np.array([eval_hermite(i, x) / np.sqrt(factorial(i)) for i in range(5)])
        assert np.allclose(vals, expected)


def test_golden_random_cases_bit_exact(S, golden):
    for name in golden["random_cases"]:
        A, b, c = golden[f"{name}_A"], golden[f"{name}_b"], complex(golden[f"{name}_c"])
        shape = tuple(int(s) for s in golden[f"{name}_shape"])
        G = S.vanilla_numba(shape, A, b, c)
        assert G.shape == shape and G.dtype == np.complex128
        assert np.array_equal(G, golden[f"{name}_G"]), name
        Gs = S.stable_numba(shape, A, b, c)
        assert np.array_equal(Gs, golden[f"{name}_Gs"]), name


def test_golden_batch_cases_bit_exact(S, golden):
    for name in golden["batch_cases"]:
        A, b, c = golden[f"{name}_A"], golden[f"{name}_b"], golden[f"{name}_c"]
        shape = tuple(int(s) for s in golden[f"{name}_shape"])
        assert np.array_equal(S.vanilla_batch_numba(shape, A, b, c), golden[f"{name}_G"]), name
        assert np.array_equal(S.vanilla_batch_numba(shape, A, b, c, True), golden[f"{name}_Gs"]), name


@pytest.mark.parametrize("shape", [(1,), (2,), (200,), (1, 1), (1, 7), (7, 1), (33, 2), (3, 1, 4), (2, 2, 2, 2, 2, 2, 2),
                                   (5, 4, 3, 2, 3), (17, 19), (64, 64), (9, 8, 7), (128, 130), (2, 300)])
@pytest.mark.parametrize("stable", [False, True])
def test_vs_oracle_shapes(S, O, shape, stable):
    A, b, c = random_triple(len(shape), (), seed=sum(shape) + len(shape))
    want = O.vanilla(shape, A, b, complex(c), stable=stable)
    got = (S.stable_numba if stable else S.vanilla_numba)(shape, A, b, complex(c))
    assert got.shape == want.shape
    assert np.array_equal(got, want)


def test_out_contract(S, O):
    # core.py:73: `out` is written in place and returned (tests/test_math/test_lattice/test_vanilla.py:185)
    A, b, c = random_triple(2, (), seed=673)
    out = np.full((3, 3), 5 - 2j)
    G = S.vanilla_numba((3, 3), A, b, complex(c), out=out)
    assert G is out
    assert np.array_equal(out, O.vanilla((3, 3), A, b, complex(c)))
    with pytest.raises(ValueError):
        S.vanilla_numba((3, 3), A, b, complex(c), out=np.zeros((4, 4), complex))
    with pytest.raises(ValueError):
        S.vanilla_numba((3, 0), A, b, complex(c))


def test_cfg1(S, golden):
    A, b, c = golden["cfg1_A"], golden["cfg1_b"], complex(golden["cfg1_c"])
    assert np.array_equal(S.vanilla_numba((200,), A, b, c), golden["cfg1_G"])
    assert np.array_equal(S.stable_numba((200,), A, b, c), golden["cfg1_G_stable"])


def test_cfg2_bit_exact_full_size(S, O, golden):
    A, b, c = golden["cfg2_A"], golden["cfg2_b"], complex(golden["cfg2_c"])
    assert np.array_equal(S.vanilla_numba((12,) * 4, A, b, c), golden["cfg2_G12"])
    assert np.array_equal(S.stable_numba((12,) * 4, A, b, c), golden["cfg2_G12_stable"])
    G = S.vanilla_numba((50,) * 4, A, b, c)
    assert G.shape == (50,) * 4
    assert sha(G) == str(golden["cfg2_G50_sha"])
    assert np.array_equal(G.ravel()[::9973], golden["cfg2_G50_sample"])
    assert np.array_equal(G, O.vanilla((50,) * 4, A, b, c))
    Gs = S.stable_numba((50,) * 4, A, b, c)
    assert sha(Gs) == str(golden["cfg2_G50_stable_sha"])
    G = S.vanilla_numba((50,) * 4, golden["cfg2r_A"], golden["cfg2r_b"], complex(golden["cfg2r_c"]))
    assert sha(G) == str(golden["cfg2r_G50_sha"])


def test_cfg5_forward_full_size(S, golden):
    A, b, c = golden["cfg5_A"], golden["cfg5_b"], complex(golden["cfg5_c"])
    assert np.array_equal(S.vanilla_numba((8,) * 4, A, b, c), golden["cfg5_G8"])
    G = S.vanilla_numba((40,) * 4, A, b, c)
    assert sha(G) == str(golden["cfg5_G40_sha"])


def test_cfg4_eight_modes(S, O, golden):
    A, b, c = golden["cfg4_A"], golden["cfg4_b"], complex(golden["cfg4_c"])
    assert np.array_equal(S.vanilla_numba((3,) * 8, A, b, c), golden["cfg4_G3"])
    shape = (5, 4, 5, 4, 5, 4, 5, 4)
    assert np.array_equal(S.vanilla_numba(shape, A, b, c), O.vanilla(shape, A, b, c))


def test_cfg3_full_batch_bit_exact(S, golden):
    A, b, c = random_triple(2, (65536,), seed=673)
    assert sha(np.concatenate([A.ravel(), b.ravel(), c.ravel()])) == str(golden["cfg3_in_sha"])
    G = S.vanilla_batch_numba((40, 40), A, b, c)
    assert G.shape == (65536, 40, 40)
    assert np.array_equal(G[:4], golden["cfg3_G_first4"])
    assert np.array_equal(G.ravel()[::1000003], golden["cfg3_G_sample"])
    for i, h in enumerate(golden["cfg3_chunk_sha"]):
        assert sha(G[i * 4096:(i + 1) * 4096]) == str(h), f"chunk {i}"
    Gs = S.vanilla_batch_numba((40, 40), A[:64].copy(), b[:64].copy(), c[:64].copy(), True)
    assert sha(Gs) == str(golden["cfg3_Gs64_sha"])


def test_batched_medium_lattices_vs_oracle(S, O):
    # batch small, lattice large enough for the cooperative all-SM kernel
    A, b, c = random_triple(3, (3,), seed=5)
    shape = (30, 31, 32)
    assert np.array_equal(S.vanilla_batch_numba(shape, A, b, c), O.vanilla_batch(shape, A, b, c))
    assert np.array_equal(S.vanilla_batch_numba(shape, A, b, c, True), O.vanilla_batch(shape, A, b, c, stable=True))


@pytest.mark.parametrize("shape", [(40, 40), (5, 2), (9, 31), (6, 33), (4, 100), (3, 256), (3, 257), (1, 17), (17, 1),
                                   (3, 4, 40), (2, 3, 5, 7), (200,), (5,), (3,)])
def test_lane_march_same_bits_as_stage_march(S, O, shape, monkeypatch):
    # stage D-2 of a batch >= 256 goes through the warp-synchronous kernel (mmh_lanes.cu); 257 + 3 triples leave a ragged
    # last warp.  Oracle on a sample, the shared-memory stage march (MMH_NO_LANES) on everything.
    A, b, c = random_triple(len(shape), (260,), seed=29)
    G = S.vanilla_batch_numba(shape, A, b, c)
    for l in (0, 1, 127, 255, 256, 259):
        assert np.array_equal(G[l], O.vanilla(shape, A[l], b[l], c[l])), l
    monkeypatch.setenv("MMH_NO_LANES", "1")
    assert sha(G) == sha(S.vanilla_batch_numba(shape, A, b, c))


def test_batches_of_medium_lattices_both_schedules(S, O, monkeypatch):
    # lattices above the one-CTA shared-memory size: a few dozen go one CTA per lattice, a handful go through the pipelined
    # all-SM path (forward_impl's cost rule); both schedules must give the oracle's bits
    shape = (14, 13, 12, 11)
    A, b, c = random_triple(4, (40,), seed=17)
    want = O.vanilla_batch(shape, A, b, c)
    assert np.array_equal(S.vanilla_batch_numba(shape, A, b, c), want)
    assert np.array_equal(S.vanilla_batch_numba(shape, A[:3].copy(), b[:3].copy(), c[:3].copy()), want[:3])
    for thr in ("1", "1000000"):
        monkeypatch.setenv("MMH_PER_CTA_BATCH", thr)
        assert np.array_equal(S.vanilla_batch_numba(shape, A, b, c), want), thr


@pytest.mark.parametrize("shape,batch", [((20, 20, 20, 20), 3), ((5, 40, 41), 7), ((3, 6, 6, 6, 6, 6), 4), ((3, 6, 6, 6, 6), 4), ((4, 1100), 5),
                                         ((7, 2, 3, 300), 3), ((9, 33, 1, 37), 6)])
def test_box_march_vs_oracle(S, O, shape, batch, monkeypatch):
    # panels above 1024 points in the one-CTA-per-lattice schedule: the box march (mmh_box.cu); MMH_PER_CTA_BATCH=1 forces
    # that schedule for any batch
    monkeypatch.setenv("MMH_PER_CTA_BATCH", "1")
    A, b, c = random_triple(len(shape), (batch,), seed=41)
    G = S.vanilla_batch_numba(shape, A, b, c)
    assert np.array_equal(G, O.vanilla_batch(shape, A, b, c))
    monkeypatch.setenv("MMH_BOX_R", "4")     # panel history in shared memory, four slots per thread
    assert np.array_equal(G, S.vanilla_batch_numba(shape, A, b, c))
    monkeypatch.setenv("MMH_NO_BOX", "1")
    assert np.array_equal(G, S.vanilla_batch_numba(shape, A, b, c))


@pytest.mark.parametrize("tag", ["a", "b", "c", "d"])
def test_binomial_golden(S, golden, tag):
    c0, c1, max_l2, gc = golden[f"bin_{tag}_args"]
    G, norm = S.binomial((int(c0), int(c1)), golden["bin_A"], golden["bin_b"], complex(golden["bin_c"]), max_l2, int(gc))
    assert np.array_equal(G, golden[f"bin_{tag}_G"])      # same early-stop level, same values
    assert abs(norm - float(golden[f"bin_{tag}_norm"])) <= 1e-12 * abs(norm)


def test_binomial_3d_and_vs_vanilla(S, O, golden):
    G, norm = S.binomial((4, 3, 5), golden["bin3_A"], golden["bin3_b"], complex(golden["bin3_c"]), 1e9, 10)
    assert np.array_equal(G, golden["bin3_G"])
    # reference test_vanillaNumba_vs_binomial (tests/test_math/test_lattice/test_lattice_functions.py:117-134)
    A, b, c = golden["bin_A"], golden["bin_b"], complex(golden["bin_c"])
    ket_vanilla = S.vanilla_numba((10, 10), A, b, c)[:5, :5]
    ket_binomial = S.binomial((5, 5), A, b, c, max_l2=0.9999, global_cutoff=12)[0][:5, :5]
    assert np.allclose(ket_vanilla, ket_binomial)


def test_device_pointer_entry_points(golden):
    """mmh_forward / mmh_forward_batched with device pointers on a non-default stream (torch = plumbing only)."""
    import torch
    from mrmustard_b200 import _lib
    A, b, c = golden["cfg2_A"], golden["cfg2_b"], golden["cfg2_c"].reshape(1)
    dev = torch.device("cuda:0")
    dA = torch.from_numpy(np.ascontiguousarray(A)).to(dev)
    db = torch.from_numpy(np.ascontiguousarray(b)).to(dev)
    dc = torch.from_numpy(np.ascontiguousarray(c)).to(dev)
    shape = (12,) * 4
    dG = torch.empty(shape, dtype=torch.complex128, device=dev)
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        _lib.check(_lib.lib.mmh_forward(4, _lib.shape_array(shape), dA.data_ptr(), db.data_ptr(), dc.data_ptr(),
                                        dG.data_ptr(), 0, ctypes.c_void_p(st.cuda_stream)))
    st.synchronize()
    assert np.array_equal(dG.cpu().numpy(), golden["cfg2_G12"])
    assert _lib.launch_count() > 0


def test_reference_manager_semantics(O):
    """hermite_renormalized batching/broadcast semantics (reference tests test_vanilla.py:165-238)."""
    import mrmustard_b200 as mm
    for stable in (True, False):
        A, b, c = random_triple(2, (), seed=673)
        G = mm.hermite_renormalized(A, b, c, (3, 3), stable=stable)
        assert G.shape == (3, 3)
        out_arr = np.zeros((3, 3), dtype=np.complex128)
        G = mm.hermite_renormalized(A, b, c, (3, 3), stable=stable, out=out_arr)
        assert out_arr is G
        assert np.allclose(G, mm.hermite_renormalized(A, b, c, (3, 3), stable=stable))
        A, b, c = random_triple(2, (2, 1), seed=673)
        shape = (4, 5)
        G = mm.hermite_renormalized(A[0, 0], b, c[0, 0], shape, stable=stable)   # b-batched
        assert G.shape == (2, 1, *shape)
        assert np.array_equal(G[1, 0], O.vanilla(shape, A[0, 0], b[1, 0], complex(c[0, 0]), stable=stable))
        out_arr = np.zeros((2, 1, *shape), dtype=np.complex128)
        G = mm.hermite_renormalized(A, b, c, shape, stable=stable, out=out_arr)  # fully batched with out
        assert G.shape == (2, 1, *shape)
        assert np.array_equal(out_arr[1, 0], O.vanilla(shape, A[1, 0], b[1, 0], complex(c[1, 0]), stable=stable))
        with pytest.raises(ValueError):
            mm.hermite_renormalized(A, b, c, shape, out=np.zeros((2, 1, 3, 5), complex))
        with pytest.raises(ValueError):
            mm.hermite_renormalized(A, b[:1], c, shape)


@pytest.mark.parametrize("tag", ["dg", "sg"])
def test_stable_cutoff_1000(S, golden, tag):
    """Reference test_vanilla_stable (test_lattice_functions.py:137-149): Dgate(4+4j) / Sgate(r=4) at cutoff 1000 with the
    stable rule, bit-identical to stable_numba, equal to the closed-form displacement, bounded for the squeezer."""
    A, b, c = golden[f"st_{tag}_A"], golden[f"st_{tag}_b"], complex(golden[f"st_{tag}_c"])
    G = S.stable_numba((1000, 1000), A, b, c)
    assert sha(G) == str(golden[f"st_{tag}_sha"])
    assert np.array_equal(G.ravel()[::7919], golden[f"st_{tag}_sample"])
    if tag == "dg":
        assert np.allclose(G.ravel()[::7919], golden["st_dg_closed_form_sample"])
    else:
        assert np.max(np.abs(G)) < 1


def test_underflowing_and_overflowing_lattices_bit_exact(S, O):
    """Amplitudes that run through the scaled-division ranges (|x| < 2^-899 down to subnormals and exact zeros; |x| > 2^900 up to the
    largest finite doubles): the quotient must stay the correctly rounded IEEE quotient of the reference on every path (single CTA,
    chain, tiled, batched), bit for bit."""
    def triple(D, scale, seed=11):
        rng = np.random.RandomState(seed)
        A = rng.random((D, D)) + 1j * rng.random((D, D)); A = (A + A.T) * scale
        return A, (rng.random(D) + 1j * rng.random(D)) * scale, 0.7 - 0.2j

    # (Once an amplitude overflows to inf the reference's nan pattern is NOT reproduced: numba divides complex by a real promoted to
    #  complex, which turns inf + i y into inf + i nan, the kernels divide the components.  The cases below stay finite.)
    for shape, scale, kind in [((400, 500), 0.05, "tiny"), ((900,), 0.02, "tiny"), ((1200, 300), 0.08, "tiny"),
                               ((225, 225), 9.0, "huge"), ((420,), 40.0, "huge")]:
        A, b, c = triple(len(shape), scale)
        want = O.vanilla(shape, A, b, c)
        got = S.vanilla_numba(shape, A, b, c)
        mag = np.abs(want)
        assert np.isfinite(want).all(), shape
        if kind == "tiny":
            assert ((mag < 2.0 ** -899) & (mag > 0)).any(), shape          # the case really exercises the scaled range
        else:
            assert (mag > 2.0 ** 900).any(), shape
        assert np.array_equal(got, want), shape
    A, b, c = random_triple(2, (40,), seed=2)
    A = A * 0.05; b = b * 0.05
    want = O.vanilla_batch((300, 200), A, b, c)
    assert ((np.abs(want) < 2.0 ** -899) & (np.abs(want) > 0)).any()
    assert np.array_equal(S.vanilla_batch_numba((300, 200), A, b, c), want)


@pytest.mark.parametrize("shape", [(7, 6, 9, 8), (13, 5, 14, 7), (6, 6, 6, 6), (20, 33, 17), (15, 14, 29), (70, 130), (62, 63), (5, 1, 9, 13)])
def test_stable_box_wavefront_vs_oracle(S, O, monkeypatch, shape):
    """k_stable_boxes (mmh_stable_boxes.cu: persistent CTAs take boxes of the lattice in a topological order, two-deep halos from the
    lattice, local level wavefront in shared memory) forced onto small lattices with ragged boxes in 2, 3 and 4 indices: bit-identical
    to the oracle's stable rule, also when called twice in a row (flags and ticket are reset per call)."""
    from mrmustard_b200 import _lib
    monkeypatch.setenv("MMH_STABLE_BOXES_MIN_N", "1")
    A, b, c = random_triple(len(shape), (), seed=17 + len(shape))
    want = O.vanilla(shape, A, b, complex(c), stable=True)
    for _ in range(2):
        n0 = _lib.launch_count()
        got = S.stable_numba(shape, A, b, complex(c))
        assert _lib.launch_count() == n0 + 1
        assert np.array_equal(got, want), shape


def test_stable_box_wavefront_batched_and_default(S, O, golden):
    # a batch of lattices large enough for the box wavefront by default (one launch per lattice)
    shape = (24, 25, 26, 27)
    A, b, c = random_triple(4, (2,), seed=23)
    G = S.vanilla_batch_numba(shape, A, b, c, True)
    for l in range(2):
        assert np.array_equal(G[l], O.vanilla(shape, A[l], b[l], complex(c[l]), stable=True)), l
