"""GPU test of the drop-in: the LIVE unmodified reference (baseline/_ref on the GPU box,
/root/reference in the build container) is driven through `dropin.install()` with the real CUDA path underneath, and its own
callers -- CircuitComponent.fock_array, Ket.fock_array, math.hermite_renormalized* -- must return what the stock numba path
returns: bit-identical for the vanilla/stable/batched lattices, 1e-10 / 1e-14 for the VJP and the compactFock paths."""
import numpy as np
import pytest

from conftest import assert_parity
from oracle import refimport

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not refimport.available(), reason="no reference install (baseline/_ref) in this tree")]


@pytest.fixture(scope="module")
def ref():
    refimport.install_shims(with_lab=True)
    import mrmustard
    from mrmustard_b200 import _lib, dropin
    yield mrmustard, dropin, _lib
    dropin.uninstall()


def _both(dropin, _lib, fn):
    """fn() on the stock reference, then on the installed CUDA path; returns (want, got, kernel launches of the second)."""
    dropin.uninstall()
    want = fn()
    dropin.install()
    n0 = _lib.launch_count()
    try:
        got = fn()
    finally:
        n1 = _lib.launch_count()
        dropin.uninstall()
    return want, got, n1 - n0


def test_cfg2_fock_array_on_cuda_path(ref):
    """BASELINE config 2 as the user writes it: (BSgate >> Sgate >> Sgate).fock_array((50,)*4)."""
    mm, dropin, _lib = ref
    from mrmustard.lab import BSgate, Sgate
    u = BSgate((0, 1), theta=0.5, phi=0.2) >> Sgate(0, r=0.3) >> Sgate(1, r=0.2)
    want, got, launches = _both(dropin, _lib, lambda: np.asarray(u.fock_array((50, 50, 50, 50))))
    assert launches > 0, "the installed drop-in did not launch any CUDA kernel"
    assert got.shape == want.shape == (50, 50, 50, 50) and got.dtype == want.dtype
    assert np.array_equal(got, want)


def test_cfg1_and_ket_fock_array(ref):
    mm, dropin, _lib = ref
    from mrmustard.lab import DisplacedSqueezed, Ket
    st = DisplacedSqueezed(0, r=0.5, alpha=0.3)
    want, got, launches = _both(dropin, _lib, lambda: np.asarray(st.fock_array(200)))
    assert launches > 0 and np.array_equal(got, want)
    with mm.settings(SEED=5):
        k = Ket.random((0, 1, 2))
    want, got, launches = _both(dropin, _lib, lambda: np.asarray(k.fock_array((9, 8, 7))))
    assert launches > 0 and np.array_equal(got, want)


def test_manager_entry_points(ref):
    mm, dropin, _lib = ref
    math = mm.math
    rng = np.random.RandomState(4)
    A = rng.random((5, 3, 3)) + 1j * rng.random((5, 3, 3)); A = (A + A.transpose(0, 2, 1)) / 6
    b = rng.random((5, 3)) + 1j * rng.random((5, 3)); c = rng.random(5) + 0j
    for stable in (False, True):
        want, got, launches = _both(dropin, _lib, lambda: np.asarray(math.hermite_renormalized(A, b, c, (6, 5, 4), stable=stable)))
        assert launches > 0 and np.array_equal(got, want), f"batched, stable={stable}"
        want, got, launches = _both(dropin, _lib, lambda: np.asarray(math.hermite_renormalized(A[0], b, c[0], (6, 5, 4), stable=stable)))
        assert launches > 0 and np.array_equal(got, want), f"b-batched, stable={stable}"
    want, got, launches = _both(dropin, _lib, lambda: np.asarray(math.hermite_renormalized_binomial(A[1], b[1], c[1], (7, 6, 5), 0.9, None)))
    assert launches > 0 and np.array_equal(got, want)
    out = np.zeros((5, 6, 5, 4), dtype=np.complex128)
    dropin.install()
    try:
        res = math.hermite_renormalized(A, b, c, (6, 5, 4), out=out)
    finally:
        dropin.uninstall()
    assert np.shares_memory(res, out) and np.array_equal(out, np.asarray(math.hermite_renormalized(A, b, c, (6, 5, 4))))


def test_strategy_vjps_by_name(ref):
    """The jax bwd rules call strategies.vanilla_vjp_numba / vanilla_batch_vjp_numba by name (jax_vjps/hermite.py:86-102,146-175)."""
    mm, dropin, _lib = ref
    from mrmustard.math.lattice import strategies as S
    rng = np.random.RandomState(9)
    A = rng.random((4, 4)) + 1j * rng.random((4, 4)); A = (A + A.T) / 8
    b = rng.random(4) + 1j * rng.random(4); c = 0.7 + 0.1j
    shape = (7, 6, 5, 8)
    G = np.asarray(S.vanilla_numba(shape, A, b, c))
    g = rng.standard_normal(shape) + 1j * rng.standard_normal(shape)
    want, got, launches = _both(dropin, _lib, lambda: S.vanilla_vjp_numba(G, c, g))
    assert launches > 0
    for x, y, nm in zip(got, want, ("dLdA", "dLdb", "dLdc")):
        assert_parity(np.asarray(x, dtype=np.complex128), np.asarray(y, dtype=np.complex128), nm)
    Gb = np.stack([G, 0.5 * G]); gb = np.stack([g, g[::-1].copy()]); cb = np.array([c, 0.5 * c])
    want, got, launches = _both(dropin, _lib, lambda: S.vanilla_batch_vjp_numba(Gb, cb, gb))
    assert launches > 0
    for x, y, nm in zip(got, want, ("dLdA", "dLdb", "dLdc")):
        assert_parity(np.asarray(x), np.asarray(y), "batched " + nm)


def test_compactfock_entry_points(ref):
    mm, dropin, _lib = ref
    math = mm.math
    from mrmustard.lab import DM, Dgate
    with mm.settings(SEED=21):
        st = DM.random((0, 1, 2)) >> Dgate(0, 0.1) >> Dgate(1, 0.2) >> Dgate(2, 0.3)
    A, b, c = (np.asarray(x, dtype=np.complex128) for x in st.bargmann_triple())
    want, got, launches = _both(dropin, _lib, lambda: np.asarray(math.hermite_renormalized_diagonal(A, b, c, cutoffs=(5, 6, 4))))
    assert launches > 0
    assert_parity(np.ascontiguousarray(got), np.ascontiguousarray(want), "diagonal")
    want, got, launches = _both(dropin, _lib, lambda: np.asarray(
        math.hermite_renormalized_1leftoverMode(A, b, c, output_cutoff=3, pnr_cutoffs=(2, 3))))
    assert launches > 0 and got.shape == want.shape
    assert np.allclose(got, want, rtol=1e-9, atol=1e-12)


def test_gate_fock_arrays_on_cuda_path(ref):
    """Dgate / Sgate / BSgate / SqueezedVacuum.fock_array reach strategies.displacement / squeezer / beamsplitter / squeezed by name
    (lab/transformations/{dgate,sgate,bsgate}.py, lab/states/squeezed_vacuum.py -> backend_numpy.py:452-475)."""
    mm, dropin, _lib = ref
    from mrmustard.lab import BSgate, Dgate, Sgate, SqueezedVacuum
    want, got, launches = _both(dropin, _lib, lambda: np.asarray(Sgate(0, r=0.6, phi=0.9).fock_array((30, 30))))
    assert launches > 0 and np.array_equal(got + 0.0, want + 0.0)
    want, got, launches = _both(dropin, _lib, lambda: np.asarray(SqueezedVacuum(0, r=0.5, phi=0.3).fock_array((40,))))
    assert launches > 0 and np.array_equal(got + 0.0, want + 0.0)
    for method in ("vanilla", "stable"):
        want, got, launches = _both(dropin, _lib, lambda: np.asarray(BSgate((0, 1), theta=0.7, phi=1.3).fock_array((9, 8, 9, 8), method=method)))
        assert launches > 0 and np.array_equal(got + 0.0, want + 0.0), method
    want, got, launches = _both(dropin, _lib, lambda: np.asarray(Dgate(0, alpha=0.4 - 0.3j).fock_array((25, 25))))
    assert launches > 0
    assert_parity(np.ascontiguousarray(got), np.ascontiguousarray(want), "Dgate")
