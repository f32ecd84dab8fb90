"""GPU tests of the Fock-space consumers: label contraction (ArrayAnsatz.contract,
physics/ansatz/array_ansatz.py:159-225) and reduce (:227-267) against numpy's einsum on the same inputs, numpy and device paths,
and through the drop-in against the live reference's CircuitComponent contraction in the Fock representation."""
import numpy as np
import pytest

from conftest import assert_parity

pytestmark = pytest.mark.gpu


def _ref_contract(a1, idx1, a2, idx2, idx_out):
    """The reference's algorithm restated with numpy (array_ansatz.py:190-222)."""
    labels = sorted(set(idx1) | set(idx2), key=lambda x: (isinstance(x, int), x))
    ch = {lab: chr(97 + i) for i, lab in enumerate(labels)}
    s1, s2 = [slice(None)] * len(idx1), [slice(None)] * len(idx2)
    for lab in set(idx1) & set(idx2):
        p1, p2 = idx1.index(lab), idx2.index(lab)
        m = min(a1.shape[p1], a2.shape[p2])
        s1[p1] = slice(0, m); s2[p2] = slice(0, m)
    es = "".join(ch[i] for i in idx1) + "," + "".join(ch[i] for i in idx2) + "->" + "".join(ch[i] for i in idx_out)
    return np.einsum(es, a1[tuple(s1)], a2[tuple(s2)])


CASES = [
    ((7, 6), [0, 1], (6, 5), [1, 2], [0, 2]),                                        # matrix product
    ((9, 8, 9, 8), [0, 1, 2, 3], (7, 10), [2, 3], [0, 1]),                           # gate applied to a ket, dims truncated to the minimum
    ((5, 4, 6), [0, 1, 2], (6, 4, 3), [2, 1, 3], [3, 0]),                            # permuted output
    ((3, 5, 4), ["b", 0, 1], (3, 4, 6), ["b", 1, 2], ["b", 0, 2]),                   # shared batch label
    ((2, 3, 5), ["a", "b", 0], (4, 5, 6), ["c", 0, 1], ["a", "c", "b", 1]),          # outer product over batch labels
    ((6, 7), [0, 1], (5,), [2], [2, 0]),                                             # label summed inside one operand (1 not in the output)
    ((8,), [0], (8,), [0], []),                                                      # inner product -> scalar
    ((70, 33, 3), [0, 1, 2], (33, 3, 65), [1, 2, 3], [0, 3]),                        # several tiles, ragged edges
]


@pytest.mark.parametrize("sh1,idx1,sh2,idx2,idx_out", CASES)
def test_contract_vs_einsum(sh1, idx1, sh2, idx2, idx_out):
    import torch
    from mrmustard_b200 import fock
    rng = np.random.RandomState(len(sh1) * 10 + len(sh2))
    a1 = rng.standard_normal(sh1) + 1j * rng.standard_normal(sh1)
    a2 = rng.standard_normal(sh2) + 1j * rng.standard_normal(sh2)
    want = _ref_contract(a1, idx1, a2, idx2, idx_out)
    got = fock.contract(a1, idx1, a2, idx2, idx_out)
    assert got.shape == want.shape and got.dtype == np.complex128
    assert np.allclose(got, want, rtol=1e-12, atol=1e-12)
    gd = fock.contract(torch.from_numpy(a1).cuda(), idx1, torch.from_numpy(a2).cuda(), idx2, idx_out)
    assert gd.is_cuda and np.array_equal(gd.cpu().numpy(), np.asarray(got))         # same kernel, same bits


def test_contract_errors():
    from mrmustard_b200 import fock
    a = np.zeros((3, 3), complex)
    with pytest.raises(ValueError):
        fock.contract(a, [0], a, [0, 1], [1])
    with pytest.raises(ValueError):
        fock.contract(a, [0, 1], a, [1, 2], [5])
    with pytest.raises(NotImplementedError):
        fock.contract(a, [0, 0], a, [0, 1], [1])


def test_reduce_device_and_host():
    import torch
    from mrmustard_b200 import fock
    a = np.arange(2 * 4 * 5 * 3, dtype=float).reshape(2, 4, 5, 3) + 0j
    for shape in [(4, 5, 3), (2, 2, 2), (1, 5, 1), (6, 5, 4), (3, 7, 2)]:
        want = np.zeros((2, *shape), complex)
        sl = tuple(slice(0, min(s, t)) for s, t in zip(shape, a.shape[1:]))
        want[(slice(None), *sl)] = a[(slice(None), *sl)]
        got_h = fock.reduce(a, shape, batch_dims=1)
        got_d = fock.reduce(torch.from_numpy(a).cuda(), shape, batch_dims=1)
        assert np.array_equal(np.asarray(got_h), want), shape
        assert np.array_equal(got_d.cpu().numpy(), want), shape


def test_lattice_consumed_on_the_device(golden):
    """The point of the row: hermite_renormalized(device=True) -> contract, nothing crosses PCIe in between.  U|psi> for the cfg2
    unitary at cutoff 12 against the same contraction done by numpy on the host arrays."""
    import torch
    import mrmustard_b200 as mm
    from mrmustard_b200 import device, fock
    A, b, c = golden["cfg2_A"], golden["cfg2_b"], golden["cfg2_c"]
    to = lambda x: torch.from_numpy(np.ascontiguousarray(np.asarray(x, dtype=np.complex128))).cuda()
    U = device.hermite_renormalized(to(A), to(b), to(c), (12, 12, 12, 12))
    rng = np.random.RandomState(0)
    psi = rng.standard_normal((10, 12)) + 1j * rng.standard_normal((10, 12))
    out = fock.contract(U, [0, 1, 2, 3], to(psi), [2, 3], [0, 1])
    want = np.einsum("abcd,cd->ab", golden["cfg2_G12"][:, :, :10, :], psi)
    assert_parity(out.cpu().numpy(), want, "U|psi>")


def test_circuit_contraction_in_fock_through_the_dropin():
    from oracle import refimport
    if not refimport.available():
        pytest.skip("no reference install in this tree")
    refimport.install_shims(with_lab=True)
    import mrmustard as mmr
    from mrmustard.lab import BSgate, Ket, Sgate
    from mrmustard_b200 import _lib, dropin
    with mmr.settings(SEED=3):
        k = Ket.random((0, 1))
    u = BSgate((0, 1), theta=0.4, phi=0.3) >> Sgate(0, r=0.2)
    kf, uf = k.to_fock((9, 8)), u.to_fock((7, 6, 9, 8))
    want = np.asarray((kf >> uf).fock_array())
    dropin.install(fock=True)
    try:
        n0 = _lib.launch_count()
        got = np.asarray((kf >> uf).fock_array())
        assert _lib.launch_count() > n0
    finally:
        dropin.uninstall()
    assert got.shape == want.shape
    assert np.allclose(got, want, rtol=1e-10, atol=1e-13)
