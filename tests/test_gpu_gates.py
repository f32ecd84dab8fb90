"""GPU parity tests of the gate-specific Fock strategies against the reference's golden vectors
(tests/golden/gen_golden_gates.py).  squeezer / squeezed / beamsplitter / stable_beamsplitter are recurrences in exactly the
reference's IEEE operations and must be BIT-IDENTICAL; the displacement (per-element log / exp) and all derivatives are held to
the north_star gate 1e-10 rel / 1e-14 abs."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, assert_parity

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gg():
    return np.load(os.path.join(GOLDEN, "gates_golden.npz"))


@pytest.fixture(scope="module")
def S():
    from mrmustard_b200 import strategies
    return strategies


def _pick(a, big):
    return a.ravel()[::7] if big else a


def _same_bits(x, y):
    return np.array_equal(np.asarray(x) + 0.0, np.asarray(y) + 0.0)     # fold -0.0 into +0.0


def test_squeezer_bit_exact_and_vjp(S, gg):
    for tag in gg["sq_cases"]:
        shape = tuple(int(x) for x in gg[f"{tag}_shape"]); r = float(gg[f"{tag}_r"]); th = float(gg[f"{tag}_theta"]); big = bool(gg[f"{tag}_big"])
        G = S.squeezer(shape, r, th)
        assert G.shape == shape and G.dtype == np.complex128
        assert _same_bits(_pick(G, big), gg[f"{tag}_G"]), tag
        k = int(gg[f"{tag}_gseed"])
        g = np.random.RandomState(k).standard_normal(shape) + 1j * np.random.RandomState(k + 1000).standard_normal(shape)
        dr, dphi = S.squeezer_vjp(G, g, r, th)
        assert_parity(np.float64(dr), np.float64(gg[f"{tag}_dr"]), tag + " dr")
        assert_parity(np.float64(dphi), np.float64(gg[f"{tag}_dphi"]), tag + " dphi")


def test_squeezed_bit_exact_and_vjp(S, gg):
    for tag in gg["sqz_cases"]:
        cut = int(gg[f"{tag}_cut"]); r = float(gg[f"{tag}_r"]); th = float(gg[f"{tag}_theta"])
        G = S.squeezed(cut, r, th)
        assert G.shape == (cut,)
        assert _same_bits(G, gg[f"{tag}_G"]), tag
        dr, dphi = S.squeezed_vjp(G, gg[f"{tag}_g"], r, th)
        assert_parity(np.float64(dr), np.float64(gg[f"{tag}_dr"]), tag + " dr")
        assert_parity(np.float64(dphi), np.float64(gg[f"{tag}_dphi"]), tag + " dphi")


def test_beamsplitter_bit_exact_and_vjp(S, gg):
    for tag in gg["bs_cases"]:
        shape = tuple(int(x) for x in gg[f"{tag}_shape"]); th = float(gg[f"{tag}_theta"]); ph = float(gg[f"{tag}_phi"])
        G = S.beamsplitter(shape, th, ph)
        Gs = S.stable_beamsplitter(shape, th, ph)
        assert G.shape == shape and Gs.shape == shape
        if f"{tag}_G" in gg.files:
            assert _same_bits(G, gg[f"{tag}_G"]), tag
            assert _same_bits(Gs, gg[f"{tag}_Gs"]), tag + " stable"
            g = gg[f"{tag}_g"]
        else:
            assert _same_bits(G.ravel()[::101], gg[f"{tag}_Gsample"]), tag
            assert _same_bits(Gs.ravel()[::101], gg[f"{tag}_Gssample"]), tag + " stable"
            assert np.isclose(np.sum(np.abs(G) ** 2), float(gg[f"{tag}_Gabs2"]), rtol=1e-12)
            assert np.isclose(np.sum(np.abs(Gs) ** 2), float(gg[f"{tag}_Gsabs2"]), rtol=1e-12)
            g = np.random.RandomState(int(gg[f"{tag}_gseed"])).standard_normal(shape) + 0j
        dth, dph = S.beamsplitter_vjp(G, g, th, ph)
        assert_parity(np.float64(dth), np.float64(gg[f"{tag}_dtheta"]), tag + " dtheta")
        assert_parity(np.float64(dph), np.float64(gg[f"{tag}_dphi"]), tag + " dphi")


def test_displacement_and_derivatives(S, gg):
    for tag in gg["disp_cases"]:
        cut = tuple(int(x) for x in gg[f"{tag}_cut"]); alpha = complex(gg[f"{tag}_alpha"]); big = bool(gg[f"{tag}_big"])
        D = S.displacement(cut, alpha)
        assert D.shape == cut and D.dtype == np.complex128
        assert_parity(_pick(D, big), gg[f"{tag}_D"], tag)
        if cut[0] == cut[1]:
            ja, jac = S.jacobian_displacement(D, alpha)
            gr, gphi = S.grad_displacement(D, abs(alpha), float(np.angle(alpha)))
            for got, nm in ((ja, "ja"), (jac, "jac"), (gr, "gr"), (gphi, "gphi")):
                assert_parity(_pick(got, big), gg[f"{tag}_{nm}"], f"{tag} {nm}")


def test_gates_vs_oracle_on_unseen_parameters(S):
    """Seeded parameters the goldens do not hold, against the CPU restatement (itself pinned to the goldens)."""
    from oracle import gates as og
    rng = np.random.RandomState(5)
    for _ in range(4):
        r, th = float(rng.uniform(0.05, 1.2)), float(rng.uniform(-3, 3))
        M, N = int(rng.randint(1, 40)), int(rng.randint(1, 40))
        assert _same_bits(S.squeezer((M, N), r, th), og.squeezer((M, N), r, th)), (M, N, r, th)
        assert _same_bits(S.squeezed(M + N, r, th), og.squeezed(M + N, r, th))
        shape = tuple(int(x) for x in rng.randint(1, 9, size=4))
        t, p = float(rng.uniform(0, 1.5)), float(rng.uniform(-3, 3))
        assert _same_bits(S.beamsplitter(shape, t, p), og.beamsplitter(shape, t, p)), (shape, t, p)
        assert _same_bits(S.stable_beamsplitter(shape, t, p), og.stable_beamsplitter(shape, t, p)), (shape, t, p)
        a = complex(rng.uniform(-1.5, 1.5), rng.uniform(-1.5, 1.5))
        cut = (int(rng.randint(1, 30)), int(rng.randint(1, 30)))
        assert_parity(S.displacement(cut, a), og.displacement(cut, a), f"displacement {cut} {a}")


def test_gate_argument_errors(S):
    with pytest.raises(ValueError):
        S.squeezer((3,), 0.1, 0.2)
    with pytest.raises(ValueError):
        S.beamsplitter((3, 3, 3), 0.1, 0.2)
    with pytest.raises(ValueError):
        S.squeezer((3, 0), 0.1, 0.2)
    with pytest.raises(ValueError):
        S.grad_displacement(np.zeros((3, 4), complex), 0.1, 0.2)
