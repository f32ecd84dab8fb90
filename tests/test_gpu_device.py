"""GPU tests of the device-resident API with autograd (mrmustard_b200.device): the torch stand-in for the reference's jax
custom_vjp boundary (math/jax_vjps/hermite.py:47-175)."""
import numpy as np
import pytest

from conftest import assert_parity, random_triple

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def T():
    import torch
    return torch


@pytest.fixture(scope="module")
def dv():
    from mrmustard_b200 import device
    return device


def _to(T, *xs):
    return tuple(T.from_numpy(np.ascontiguousarray(np.asarray(x, dtype=np.complex128))).cuda() for x in xs)


def test_forward_same_bits_as_numpy_api(T, dv):
    import mrmustard_b200 as mm
    for shape, batch in [((7, 6, 5), ()), ((9, 8), (5,)), ((4, 3, 2, 3), (2, 3))]:
        A, b, c = random_triple(len(shape), batch, seed=3)
        want = mm.hermite_renormalized(A, b, c, shape)
        got = dv.hermite_renormalized(*_to(T, A, b, c), shape)
        assert got.is_cuda and tuple(got.shape) == want.shape
        assert np.array_equal(got.cpu().numpy(), want)
        got2 = mm.hermite_renormalized(*_to(T, A, b, c), shape, device=True)      # the same through backend(..., device=True)
        assert np.array_equal(got2.cpu().numpy(), want)
    A, b, c = random_triple(3, (4,), seed=5)     # b-batched: A[D,D], c scalar
    want = mm.hermite_renormalized(A[0], b, c[0], (5, 4, 3))
    got = dv.hermite_renormalized(*_to(T, A[0], b), complex(c[0]), (5, 4, 3))
    assert np.array_equal(got.cpu().numpy(), want)
    want = mm.hermite_renormalized(A[1], b[1], c[1], (5, 4, 3), stable=True)
    got = dv.hermite_renormalized(*_to(T, A[1], b[1], c[1]), (5, 4, 3), stable=True)
    assert np.array_equal(got.cpu().numpy(), want)


def test_out_is_written_in_place(T, dv):
    A, b, c = random_triple(2, (6,), seed=8)
    dA, db, dc = _to(T, A, b, c)
    out = T.zeros((6, 9, 7), dtype=T.complex128, device="cuda")
    res = dv.hermite_renormalized(dA, db, dc, (9, 7), out=out)
    assert res.data_ptr() == out.data_ptr()
    assert np.array_equal(out.cpu().numpy(), dv.hermite_renormalized(dA, db, dc, (9, 7)).cpu().numpy())
    with pytest.raises(ValueError):
        dv.hermite_renormalized(dA.requires_grad_(), db, dc, (9, 7), out=out)
    with pytest.raises(TypeError):
        dv.hermite_renormalized(A, b, c, (9, 7))          # numpy arrays are not accepted by the device API


def test_vjp_reference_convention_golden(T, dv, golden):
    """device.vanilla_vjp == strategies.vanilla_vjp_numba of the reference (goldens of the 4-mode ket of cfg5 at cutoff 8)."""
    G, g, c = golden["cfg5_G8"], golden["cfg5_g8"], golden["cfg5_c"]
    dA, db, dc = dv.vanilla_vjp(*_to(T, G, c, g))
    assert_parity(dA.cpu().numpy(), golden["cfg5_dA8"], "dLdA")
    assert_parity(db.cpu().numpy(), golden["cfg5_db8"], "dLdb")
    assert_parity(dc.cpu().numpy(), golden["cfg5_dc8"], "dLdc")
    for tag in golden["batch_cases"]:
        G, g, c = golden[f"{tag}_G"], golden[f"{tag}_g"], golden[f"{tag}_c"]
        dA, db, dc = dv.vanilla_batch_vjp(*_to(T, G, c, g))
        assert_parity(dA.cpu().numpy(), golden[f"{tag}_dA"], tag + " dLdA")
        assert_parity(db.cpu().numpy(), golden[f"{tag}_db"], tag + " dLdb")
        assert_parity(dc.cpu().numpy(), golden[f"{tag}_dc"], tag + " dLdc")


def test_autograd_gradcheck(T, dv):
    """torch.autograd.gradcheck of the autograd nodes (torch's conjugate-Wirtinger convention), unbatched and batched.  A enters
    through its symmetric part because the lattice is a function of a symmetric matrix (the VJP is symmetrised, gradients.py:79)."""
    A, b, c = random_triple(3, (), seed=11)
    X, bb, cc = (t.requires_grad_() for t in _to(T, A * 0.7, b * 0.5, c))
    f = lambda X_, b_, c_: dv.hermite_renormalized((X_ + X_.T) / 2, b_, c_, (4, 3, 3))
    assert T.autograd.gradcheck(f, (X, bb, cc), eps=1e-6, atol=1e-6, rtol=1e-5)
    A, b, c = random_triple(2, (3,), seed=12)
    X, bb, cc = (t.requires_grad_() for t in _to(T, A * 0.7, b * 0.5, c))
    fb = lambda X_, b_, c_: dv.hermite_renormalized((X_ + X_.transpose(-1, -2)) / 2, b_, c_, (4, 5))
    assert T.autograd.gradcheck(fb, (X, bb, cc), eps=1e-6, atol=1e-6, rtol=1e-5)


def test_fidelity_gradient_step_matches_reference_convention(T, dv, golden):
    """cfg5 in miniature: loss = 1 - |<t|G(A,b,c)>|^2 for a fixed target t; loss.backward() on the device must give
    conj(vanilla_vjp(G, c, dL/dG)) with the reference's holomorphic cotangent dL/dG = -conj(<t|G>) conj(t) -- the quantity the
    reference's jax bwd returns and its optimizer then conjugates (training/optimizer.py:104)."""
    A, b, c = golden["cfg5_A"], golden["cfg5_b"], golden["cfg5_c"]
    shape = (8, 8, 8, 8)
    rng = np.random.RandomState(2)
    t = rng.standard_normal(shape) + 1j * rng.standard_normal(shape)
    t /= np.linalg.norm(t)
    dA, db, dc, dt = _to(T, A, b, c, t)
    dA.requires_grad_(); db.requires_grad_(); dc.requires_grad_()
    G = dv.hermite_renormalized(dA, db, dc, shape)
    ov = T.sum(dt.conj() * G)
    loss = 1.0 - (ov.real ** 2 + ov.imag ** 2)
    loss.backward()
    Gh = G.detach().cpu().numpy()
    s = np.sum(np.conj(t) * Gh)
    g_holo = -np.conj(s) * np.conj(t)                       # dL/dG_k  (L = 1 - s conj(s), s = sum conj(t_k) G_k)
    import mrmustard_b200 as mm
    rA, rb, rc = mm.strategies.vanilla_vjp_numba(Gh, complex(c), g_holo)
    # torch: grad = dL/dRe + i dL/dIm = 2 dL/d(conj theta) = 2 conj(dL/dtheta) for a real loss of a holomorphic map
    assert np.allclose(dA.grad.cpu().numpy(), 2 * np.conj(rA), rtol=1e-9, atol=1e-13)
    assert np.allclose(db.grad.cpu().numpy(), 2 * np.conj(rb), rtol=1e-9, atol=1e-13)
    assert np.allclose(dc.grad.cpu().numpy(), 2 * np.conj(rc), rtol=1e-9, atol=1e-13)


def test_fused_fidelity_step_matches_autograd_and_reference(T, dv, golden):
    """device.FidelityStep (forward + overlap + VJP with the constant cotangent, 22 numbers read back) against (i) the autograd
    path with the loss written in torch ops and (ii) the numpy-facing reference-convention VJP."""
    import mrmustard_b200 as mm
    A, b, c = golden["cfg5_A"], golden["cfg5_b"], golden["cfg5_c"]
    shape = (8, 8, 8, 8)
    rng = np.random.RandomState(4)
    t = rng.standard_normal(shape) + 1j * rng.standard_normal(shape)
    t /= np.linalg.norm(t)
    step = dv.FidelityStep(shape, t)
    loss, dA, db, dc = step(A, b, c)
    eager = dv.FidelityStep(shape, t, use_graph=False)(A, b, c)           # graph replay and eager launches give the same bits
    assert eager[0] == loss and np.array_equal(eager[1], dA) and np.array_equal(eager[2], db) and eager[3] == dc
    l3, dA3, _, _ = step(A * 0.5, b, c)                                     # the replayed graph reads the NEW inputs
    assert l3 != loss and not np.array_equal(dA3, dA)
    assert step(A, b, c)[0] == loss
    G = mm.strategies.vanilla_numba(shape, A, b, complex(c))
    s = np.sum(np.conj(t) * G)
    assert np.isclose(loss, 1.0 - abs(s) ** 2, rtol=1e-12, atol=1e-14)
    rA, rb, rc = mm.strategies.vanilla_vjp_numba(G, complex(c), -np.conj(s) * np.conj(t))
    assert_parity(dA, rA, "dLdA"); assert_parity(db, rb, "dLdb"); assert_parity(np.complex128(dc), np.complex128(rc), "dLdc")
    # device-resident inputs give the same numbers; repeated calls reuse the buffers
    loss2, dA2, db2, dc2 = step(*_to(T, A, b, c))
    assert loss2 == loss and np.array_equal(dA2, dA) and np.array_equal(db2, db) and dc2 == dc
    # autograd path (torch convention = 2 conj of the reference convention for a real loss)
    pa, pb, pc = (x.requires_grad_() for x in _to(T, A, b, c))
    Gd = dv.hermite_renormalized(pa, pb, pc, shape)
    ov = T.sum(T.from_numpy(t).cuda().conj() * Gd)
    (1.0 - (ov.real ** 2 + ov.imag ** 2)).backward()
    assert np.allclose(pa.grad.cpu().numpy(), 2 * np.conj(dA), rtol=1e-9, atol=1e-13)
    assert np.allclose(pb.grad.cpu().numpy(), 2 * np.conj(db), rtol=1e-9, atol=1e-13)
