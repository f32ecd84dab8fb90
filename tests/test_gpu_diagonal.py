"""GPU parity tests for the compactFock diagonal / one-leftover-mode paths (through the C ABI).

Gate: 1e-10 relative / 1e-14 absolute against the reference's golden vectors (the reference evaluates A[i] @ G_in
with BLAS, so summation order — and therefore the last bits — are not defined by the reference itself)."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, assert_parity, random_triple

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gd():
    return np.load(os.path.join(GOLDEN, "diagonal_golden.npz"))


@pytest.fixture(scope="module")
def mm():
    import mrmustard_b200
    return mrmustard_b200


def test_diagonal_golden(mm, gd):
    for name in gd["diag_cases"]:
        A, b, c = gd[f"{name}_A"], gd[f"{name}_b"], complex(gd[f"{name}_c"])
        cut = tuple(int(x) for x in gd[f"{name}_cut"])
        G = mm.hermite_renormalized_diagonal(A, b, c, cut)
        assert_parity(G, gd[f"{name}_G"], name)
        A2, b2 = mm.backend.reorder_AB_bargmann(A, b)
        G2 = mm.hermite_renormalized_diagonal(A2, b2, c, cut, reorderedAB=False)
        assert np.array_equal(G, G2)


def test_diagonal_b_batched(mm, gd):
    # reference test_diagonalbatchNumba_vs_diagonalNumba (tests/test_math/test_lattice/test_lattice_functions.py:66-84)
    G = mm.hermite_renormalized_diagonal(gd["db_A"], gd["db_b"], complex(gd["db_c"]), tuple(int(x) for x in gd["db_cut"]))
    assert G.shape == gd["db_G"].shape
    assert_parity(G, gd["db_G"], "b-batched")
    G0 = mm.hermite_renormalized_diagonal(gd["db_A"], gd["db_b"][:, 0].copy(), complex(gd["db_c"]), tuple(int(x) for x in gd["db_cut"]))
    assert np.allclose(G0, G[..., 0], rtol=1e-12, atol=1e-15)


def test_diagonal_equals_diagonal_of_vanilla(mm, gd):
    # reference test_compactFock_diagonal (tests/test_math/test_compactFock.py:19-44)
    A, b, c = gd["d3_A"], gd["d3_b"], complex(gd["d3_c"])
    cut = (5, 5, 5)
    G_ref = mm.hermite_renormalized(A, b, c, cut * 2)
    ref = np.array([G_ref[tuple(list(i) + list(i))] for i in np.ndindex(*cut)]).reshape(cut)
    assert np.allclose(ref, mm.hermite_renormalized_diagonal(A, b, c, cut))


def test_leftover_golden(mm, gd):
    for name in gd["leftover_cases"]:
        A, b, c = gd[f"{name}_A"], gd[f"{name}_b"], complex(gd[f"{name}_c"])
        oc, pnr = int(gd[f"{name}_oc"]), tuple(int(x) for x in gd[f"{name}_pnr"])
        G = mm.hermite_renormalized_1leftoverMode(A, b, c, oc, pnr)
        assert G.shape == (oc + 1, oc + 1) + tuple(p + 1 for p in pnr)
        assert_parity(np.ascontiguousarray(G), gd[f"{name}_Gcompact"], name + " vs compactFock")
        assert np.allclose(G, gd[f"{name}_G"], rtol=1e-9, atol=1e-12), name + " vs numpy backend"
        F = mm.strategies.fast_diagonal(A, b, c, oc, pnr)
        assert np.allclose(F, gd[f"{name}_Gfast"], rtol=1e-9, atol=1e-12), name + " fast_diagonal layout"


def test_leftover_equals_diagonals_of_vanilla(mm, gd):
    # reference test_compactFock_1leftover (tests/test_math/test_compactFock.py:47-71)
    A, b, c = gd["l3_A"], gd["l3_b"], complex(gd["l3_c"])
    G_left = mm.hermite_renormalized_1leftoverMode(A, b, c, output_cutoff=3, pnr_cutoffs=(1, 2))
    G_ref = mm.hermite_renormalized(A, b, c, (4, 2, 3, 4, 2, 3))
    expected = np.diagonal(np.diagonal(G_ref, axis1=1, axis2=4), axis1=1, axis2=3)
    assert np.allclose(expected, G_left)


def test_validation_errors(mm):
    A = np.eye(4, dtype=complex) * 0.1
    b = np.zeros(4, complex)
    with pytest.raises(TypeError):
        mm.strategies.hermite_multidimensional_diagonal([[0.1]], b, 1.0, (3,))
    with pytest.raises(ValueError):
        mm.strategies.hermite_multidimensional_diagonal(A[:, :3].copy(), b, 1.0, (3, 3))
    with pytest.raises(ValueError):
        mm.strategies.hermite_multidimensional_diagonal(A + np.triu(np.ones((4, 4)), 1), b, 1.0, (3, 3))
    with pytest.raises(ValueError):
        mm.strategies.hermite_multidimensional_diagonal(A, b, 1.0, (3, 3, 3))
    with pytest.raises(ValueError):
        mm.strategies.hermite_multidimensional_1leftoverMode(np.eye(2, dtype=complex), np.zeros(2, complex), 1.0, (3,))


def test_diagonal_jacobians_golden(mm, gd):
    """grad_hermite_multidimensional_diagonal vs the reference's forward-mode Jacobians (diagonal_grad.py)."""
    for name in gd["grad_cases"]:
        A, b, c = gd[f"{name}_A"], gd[f"{name}_b"], complex(gd[f"{name}_c"])
        cut = tuple(int(x) for x in gd[f"{name}_cut"])
        A2, b2 = mm.backend.reorder_AB_bargmann(A, b)
        dG0, dA, dB = mm.strategies.grad_hermite_multidimensional_diagonal(np.ascontiguousarray(A2), b2, c, np.empty(cut, complex))
        assert_parity(dG0, gd[f"{name}_dG0"], name + " dG0")
        assert_parity(dA, gd[f"{name}_dA"], name + " dA")
        assert_parity(dB, gd[f"{name}_dB"], name + " dB")


def test_diagonal_vjp_finite_differences(mm, gd):
    A, b, c = gd["d2_A"], gd["d2_b"], complex(gd["d2_c"])
    cut = (6, 7)
    A2, b2 = (np.ascontiguousarray(x) for x in mm.backend.reorder_AB_bargmann(A, b))
    g = np.random.RandomState(0).standard_normal(cut) + 1j * np.random.RandomState(1).standard_normal(cut)
    dLdA, dLdB, dLdC = mm.strategies.hermite_renormalized_diagonal_vjp(A2, b2, c, cut, g)
    f = lambda A_, b_, c_: np.sum(g * mm.strategies.hermite_multidimensional_diagonal(A_, b_, c_, cut))
    f0, eps = f(A2, b2, c), 1e-7
    assert np.isclose((f(A2, b2, c + eps) - f0) / eps, dLdC, rtol=1e-5, atol=1e-7)
    for i in range(4):
        bp = b2.copy(); bp[i] += eps
        assert np.isclose((f(A2, bp, c) - f0) / eps, dLdB[i], rtol=1e-4, atol=1e-6)
        for l in range(i, 4):     # the forward validates symmetry: perturb (i,l) and (l,i) together
            Ap = A2.copy(); Ap[i, l] += eps
            if l != i: Ap[l, i] += eps
            want = dLdA[i, l] + (dLdA[l, i] if l != i else 0)
            assert np.isclose((f(Ap, b2, c) - f0) / eps, want, rtol=1e-4, atol=1e-6)


def test_leftover_jacobians_finite_differences(mm, gd):
    """grad_hermite_multidimensional_1leftoverMode: the reference implementation does not compile under this image's numba
    (see tests/golden/gen_golden_diagonal.py), so the Jacobians are pinned by finite differences of the (golden-pinned)
    forward path and by d(arr0)/dG0 = arr0 / G0."""
    A, b, c = gd["l3_A"], gd["l3_b"], complex(gd["l3_c"])
    A2, b2 = (np.ascontiguousarray(x) for x in mm.backend.reorder_AB_bargmann(A, b))
    cut = (4, 2, 3)
    fwd = lambda A_, b_, c_: np.ascontiguousarray(mm.strategies.hermite_multidimensional_1leftoverMode(A_, b_, c_, cut))
    G = fwd(A2, b2, c)
    dG0, dA, dB = mm.strategies.grad_hermite_multidimensional_1leftoverMode(A2, b2, c, G)
    assert dA.shape == G.shape + (6, 6) and dB.shape == G.shape + (6,)
    assert np.allclose(dG0, G / c, rtol=1e-12, atol=1e-15)
    eps = 1e-7
    for i in range(6):
        bp = b2.copy(); bp[i] += eps
        assert np.allclose((fwd(A2, bp, c) - G) / eps, dB[..., i], rtol=1e-4, atol=1e-6), f"dB[{i}]"
        for l in range(i, 6):
            Ap = A2.copy(); Ap[i, l] += eps
            if l != i: Ap[l, i] += eps
            want = dA[..., i, l] + (dA[..., l, i] if l != i else 0)
            assert np.allclose((fwd(Ap, b2, c) - G) / eps, want, rtol=1e-4, atol=1e-6), f"dA[{i},{l}]"


@pytest.fixture(scope="module")
def glg():
    return np.load(os.path.join(GOLDEN, "leftover_grad_golden.npz"))


def test_leftover_jacobians_golden(mm, glg):
    """grad_hermite_multidimensional_1leftoverMode vs the reference's own forward-mode Jacobians
    (singleLeftoverMode_grad.py:560-724, run as plain Python by tests/golden/gen_golden_leftover_grad.py) at the north_star
    gate 1e-10 rel / 1e-14 abs.  Inputs are stored already interleaved."""
    for name in glg["cases"]:
        A2, b2, c = glg[f"{name}_A"], glg[f"{name}_b"], complex(glg[f"{name}_c"])
        cut = tuple(int(x) for x in glg[f"{name}_cut"])
        G = mm.strategies.hermite_multidimensional_1leftoverMode(A2, b2, c, cut)
        assert_parity(np.ascontiguousarray(G), glg[f"{name}_G"], name + " amplitudes")
        dG0, dA, dB = mm.strategies.grad_hermite_multidimensional_1leftoverMode(A2, b2, c, G)
        assert_parity(dG0, glg[f"{name}_dG0"], name + " dG0")
        assert_parity(dA, glg[f"{name}_dA"], name + " dA")
        assert_parity(dB, glg[f"{name}_dB"], name + " dB")


def test_eight_mode_ket_diagonal_vs_vanilla_lattice(mm, golden):
    """cfg4 as written in BASELINE.json (8-mode Gaussian ket, diagonal strategy): the diagonal of the density matrix of a ket is
    |psi_n|^2, so the M = 8 diagonal sweep (A 16x16) must reproduce the squared moduli of the 8-mode vanilla lattice.  The reference
    layout of the auxiliary arrays (137 arrays of prod(cutoffs) entries) fits one B200 up to cutoff 9; tested at cutoffs 3-4."""
    A, b, c = golden["cfg4_A"], golden["cfg4_b"], complex(golden["cfg4_c"])
    Adm = np.zeros((16, 16), complex); Adm[:8, :8] = np.conj(A); Adm[8:, 8:] = A
    bdm = np.concatenate([np.conj(b), b]); cdm = abs(c) ** 2
    for cut in [(3, 2, 3, 2, 2, 3, 2, 3), (4,) * 8]:
        got = mm.hermite_renormalized_diagonal(Adm, bdm, cdm, cut)
        psi = mm.strategies.vanilla_numba(cut, A, b, c)
        want = np.abs(psi) ** 2
        assert got.shape == cut
        assert np.all(np.abs(got - want) <= 1e-14 + 1e-10 * np.abs(want)), cut


def test_rolling_level_buffers_same_results(mm, gd, monkeypatch):
    """The rolling-level-buffer sweep (mmh_diagonal_rolling.cu; default for large sweeps) against the reference goldens and against
    the full-layout sweep on every golden diagonal case, b-batched included."""
    for name in list(gd["diag_cases"]) + ["db"]:
        A, b, c = gd[f"{name}_A"], gd[f"{name}_b"], complex(gd[f"{name}_c"])
        cut = tuple(int(x) for x in gd[f"{name}_cut"])
        monkeypatch.setenv("MMH_DIAG_ROLLING", "0")
        full = mm.hermite_renormalized_diagonal(A, b, c, cut)
        monkeypatch.setenv("MMH_DIAG_ROLLING", "1")
        roll = mm.hermite_renormalized_diagonal(A, b, c, cut)
        assert roll.shape == full.shape
        assert_parity(np.ascontiguousarray(roll), np.ascontiguousarray(gd[f"{name}_G"]), name + " rolling vs reference")
        assert np.allclose(roll, full, rtol=1e-13, atol=1e-16), name


def test_rolling_eight_modes_small_cutoffs(mm, golden, monkeypatch):
    A, b, c = golden["cfg4_A"], golden["cfg4_b"], complex(golden["cfg4_c"])
    Adm = np.zeros((16, 16), complex); Adm[:8, :8] = np.conj(A); Adm[8:, 8:] = A
    bdm = np.concatenate([np.conj(b), b]); cdm = abs(c) ** 2
    monkeypatch.setenv("MMH_DIAG_ROLLING", "1")
    for cut in [(3, 2, 3, 2, 2, 3, 2, 3), (4,) * 8, (1, 5, 1, 4, 2, 1, 3, 2)]:
        got = mm.hermite_renormalized_diagonal(Adm, bdm, cdm, cut)
        want = np.abs(mm.strategies.vanilla_numba(cut, A, b, c)) ** 2
        assert np.all(np.abs(got - want) <= 1e-14 + 1e-10 * np.abs(want)), cut


def test_cfg4_as_written_eight_modes_cutoff_12(golden):
    """BASELINE config 4 as written: the 8-mode Gaussian ket through the diagonal strategy at cutoff 12 (A 16x16, 430 M diagonal
    amplitudes).  The reference cannot run it (its auxiliary arrays would take 0.94 TB); here it must equal the squared moduli of the
    (12,)^8 vanilla lattice of the same ket within 1e-10 relative / 1e-14 absolute.  Everything stays on the device (6.9 GB per array)."""
    import ctypes
    import torch
    from mrmustard_b200 import _lib as L
    free, _ = torch.cuda.mem_get_info()
    if free < (60 << 30):
        pytest.skip("needs ~45 GB of device memory")
    dev = torch.device("cuda:0")
    A, b, c = golden["cfg4_A"], golden["cfg4_b"], complex(golden["cfg4_c"])
    Adm = np.zeros((16, 16), complex); Adm[:8, :8] = np.conj(A); Adm[8:, 8:] = A
    bdm = np.concatenate([np.conj(b), b])
    import mrmustard_b200 as mm
    A2, b2 = (np.ascontiguousarray(x) for x in mm.backend.reorder_AB_bargmann(Adm, bdm))
    to = lambda x: torch.from_numpy(np.ascontiguousarray(np.asarray(x, dtype=np.complex128))).to(dev)
    dA2, db2, dG0 = to(A2), to(b2), to(np.array([abs(c) ** 2]))
    cut = (12,) * 8
    n = 12 ** 8
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    out = torch.empty(n, dtype=torch.complex128, device=dev)
    L.check(L.lib.mmh_diagonal(8, L.shape_array(cut), dA2.data_ptr(), db2.data_ptr(), 0, dG0.data_ptr(), out.data_ptr(), st))
    dA, db, dc = to(A), to(b), to(np.array([c]))
    psi = torch.empty(n, dtype=torch.complex128, device=dev)
    L.check(L.lib.mmh_forward(8, L.shape_array(cut), dA.data_ptr(), db.data_ptr(), dc.data_ptr(), psi.data_ptr(), 0, st))
    torch.cuda.synchronize()
    want = psi.real ** 2 + psi.imag ** 2
    del psi
    err = (out.real - want).abs()
    tol = 1e-14 + 1e-10 * want
    assert bool(torch.all(err <= tol)), f"max err {float(err.max())}, max |want| {float(want.max())}"
    assert float(out.imag.abs().max()) <= 1e-14
    assert abs(float(want.sum()) - float(out.real.sum())) <= 1e-9 * float(want.sum())


def test_fast_diagonal_deviation(mm, gd):
    """The one place where this implementation deliberately differs from the reference: for 2*output_cutoff - L < 1 the reference's
    fast_diagonal weight loop (fast_diagonal.py:68) stops before the top weight and leaves those conditional density matrices zero
    (golden `lq_Gfast`), while its own compactFock strategy holds the true values (golden `lq_Gcompact`).  Ours equals the
    compactFock values everywhere and the reference's fast_diagonal wherever that one is filled."""
    A, b, c = gd["lq_A"], gd["lq_b"], complex(gd["lq_c"])
    oc, pnr = int(gd["lq_oc"]), tuple(int(x) for x in gd["lq_pnr"])
    assert 2 * oc - (len(pnr) + 1) < 1
    F = np.ascontiguousarray(mm.strategies.fast_diagonal(A, b, c, oc, pnr))
    ref_fast, ref_compact = gd["lq_Gfast"], gd["lq_Gcompact"].transpose(2, 3, 0, 1)
    assert F.shape == ref_fast.shape
    assert np.allclose(F, ref_compact, rtol=1e-9, atol=1e-13)
    missing = np.abs(ref_fast).max(axis=(-2, -1)) == 0.0            # partitions the reference never filled
    assert missing.any() and missing[pnr] and np.abs(F[pnr]).max() > 1e-6
    assert np.allclose(F[~missing], ref_fast[~missing], rtol=1e-9, atol=1e-13)
