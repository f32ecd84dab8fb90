"""The reference's OWN tests for the rows of SURVEY.md section 8f, replayed on the CUDA path (the live reference from baseline/_ref with
`dropin.install(fock=True)`): test_lattice_functions.py::test_displacement_grad (:86-104), ::test_bs_schwinger (:49-64),
::test_vanillaNumba_vs_binomial (:122-138), test_states/test_ket.py::test_auto_shape (:98-107), test_dm.py::test_auto_shape (:82-89),
test_compactFock.py::test_compactFock_diagonal / _1leftover (:19-71).  Same assertions, same tolerances, CUDA kernels underneath."""
import numpy as np
import pytest

from oracle import refimport

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not refimport.available(), reason="no reference install (baseline/_ref) in this tree")]


@pytest.fixture()
def live():
    refimport.install_shims(with_lab=True)
    import mrmustard
    from mrmustard_b200 import _lib, dropin
    dropin.install(fock=True)
    n0 = _lib.launch_count()
    yield mrmustard
    dropin.uninstall()
    assert _lib.launch_count() > n0, "the test never reached a CUDA kernel"


def test_displacement_grad(live):
    from mrmustard.math.lattice import strategies as S      # rebound to the CUDA path by the drop-in
    cutoff, r, theta = 4, 2.0, np.pi / 8
    T = S.displacement((cutoff, cutoff), r * np.exp(1j * theta))
    Dr, Dtheta = S.grad_displacement(T, r, theta)
    dr = dtheta = 0.001
    Drp = S.displacement((cutoff, cutoff), (r + dr) * np.exp(1j * theta))
    Drm = S.displacement((cutoff, cutoff), (r - dr) * np.exp(1j * theta))
    Dtp = S.displacement((cutoff, cutoff), r * np.exp(1j * (theta + dtheta)))
    Dtm = S.displacement((cutoff, cutoff), r * np.exp(1j * (theta - dtheta)))
    assert np.allclose(Dr, (Drp - Drm) / (2 * dr), atol=1e-5, rtol=0)
    assert np.allclose(Dtheta, (Dtp - Dtm) / (2 * dtheta), atol=1e-5, rtol=0)


def test_bs_schwinger(live):
    from mrmustard import math
    from mrmustard.lab import Ket, Unitary
    from mrmustard.math.lattice import strategies as S
    from mrmustard.math.lattice.strategies.beamsplitter import apply_BS_schwinger   # the reference's numpy implementation
    G = math.asnumpy(Ket.random((0, 1)).fock_array([20, 20]))
    BS = S.beamsplitter((20, 20, 20, 20), 1.0, 1.0)
    manual = np.einsum("ab, cdab", G, BS)
    assert np.allclose(manual, apply_BS_schwinger(1.0, 1.0, 0, 1, np.array(G)))
    Gg = math.asnumpy(Unitary.random((0, 1)).fock_array([20, 20, 20, 20]))
    BS = S.beamsplitter((20, 20, 20, 20), 2.0, -1.0)
    manual = np.einsum("cdab, abef", BS, Gg)
    assert np.allclose(manual, apply_BS_schwinger(2.0, -1.0, 0, 1, np.array(Gg)))


def test_vanilla_vs_binomial(live):
    from mrmustard import settings
    from mrmustard.lab import Ket
    from mrmustard.math.lattice import strategies as S
    with settings(SEED=42):   # as the reference's test: the same random ket every run (binomial stops at the l2 norm it reaches)
        A, b, c = (np.asarray(x) for x in Ket.random((0, 1)).bargmann_triple())
        ket_vanilla = S.vanilla_numba((10, 10), A, b, complex(c))[:5, :5]
        ket_binomial = S.binomial((5, 5), A, b, complex(c), max_l2=0.9999, global_cutoff=12)[0][:5, :5]
        assert np.allclose(ket_vanilla, ket_binomial)


def test_auto_shape_known_answers(live):
    from mrmustard.lab import Coherent, Number
    ket = Coherent(0, alpha=1)
    assert ket.auto_shape() == (8,)
    ket.manual_shape = (19,)
    assert ket.auto_shape() == (19,)
    assert ket.auto_shape(respect_manual_shape=False) == (8,)
    ket = Coherent(0, 1) >> Number(1, 10).dual
    assert ket.auto_shape() == (8, 11)
    dm = Coherent(0, 1).dm()
    assert dm.auto_shape() == (8, 8)
    dm = Coherent(0, 1).dm() >> Number(1, 10).dual
    assert dm.auto_shape() == (8, 11, 8, 11)


def test_compactfock_diagonal_and_leftover(live):
    from mrmustard import math
    from mrmustard.lab import DM
    cutoffs = (5, 5, 5)
    A, B, G0 = (np.asarray(x) for x in DM.random((0, 1, 2)).bargmann_triple())
    G_ref = math.asnumpy(math.hermite_renormalized(A, B, G0, shape=cutoffs * 2))
    ref_diag = np.array([G_ref[tuple(list(i) + list(i))] for i in np.ndindex(*cutoffs)]).reshape(cutoffs)
    assert np.allclose(ref_diag, math.asnumpy(math.hermite_renormalized_diagonal(A, B, G0, cutoffs=cutoffs)))
    G_left = math.asnumpy(math.hermite_renormalized_1leftoverMode(A, B, G0, output_cutoff=3, pnr_cutoffs=(1, 2)))
    G_ref = math.asnumpy(math.hermite_renormalized(A, B, G0, shape=(4, 2, 3, 4, 2, 3)))
    expected = np.diagonal(np.diagonal(G_ref, axis1=1, axis2=4), axis1=1, axis2=3)
    assert np.allclose(expected, G_left)
