"""Host-side mirror of the reference's `mrmustard.math.lattice.strategies` for the Gaussian-to-Fock path.

Same function names, argument meaning, return types and `out` contract as the reference's numba
strategies; the work is done by hand-written sm_100a kernels behind the C ABI (include/mmhermite.h).
numpy in, numpy out: inputs are copied to the device, the lattice is computed there and copied back into
page-locked host memory (or into `out`).  There is no CPU fallback.

Reference: mrmustard/math/lattice/strategies/vanilla/{core,batch,gradients}.py, strategies/binomial.py.
"""
from __future__ import annotations

import ctypes

import numpy as np

from . import _lib
from ._lib import check, lib, shape_array

__all__ = [
    "vanilla_numba", "stable_numba", "vanilla_batch_numba", "vanilla_vjp_numba",
    "vanilla_batch_vjp_numba", "binomial", "vanilla", "stable",
    "hermite_multidimensional_diagonal", "hermite_multidimensional_1leftoverMode", "fast_diagonal",
    "grad_hermite_multidimensional_diagonal", "hermite_renormalized_diagonal_vjp",
    "grad_hermite_multidimensional_1leftoverMode", "vanilla_contract_numba",
    "squeezer", "squeezed", "beamsplitter", "stable_beamsplitter", "displacement", "jacobian_displacement", "grad_displacement",
    "beamsplitter_vjp", "squeezer_vjp", "squeezed_vjp", "autoshape_numba",
]


def _c128(x, shape=None) -> np.ndarray:
    a = np.ascontiguousarray(np.asarray(x, dtype=np.complex128))
    if shape is not None:
        a = a.reshape(shape)
    return a


def _p(a: np.ndarray) -> ctypes.c_void_p:
    return ctypes.c_void_p(a.ctypes.data)


def _check_out(out, nelem: int) -> None:
    # core.py:73 `out.ravel()` + `G.reshape(shape)`: out must be C-contiguous complex128 of exactly prod(shape) entries
    if not isinstance(out, np.ndarray) or out.dtype != np.complex128 or not out.flags.c_contiguous:
        raise TypeError("out must be a C-contiguous complex128 numpy array")
    if out.size != nelem:
        raise ValueError(f"cannot reshape array of size {out.size} into the requested lattice of {nelem} entries")


def _check_shape(shape) -> tuple[int, ...]:
    shape = tuple(int(s) for s in shape)
    if any(s < 1 for s in shape):
        raise ValueError(f"shape {shape} must have all entries >= 1")
    return shape


def _forward(shape, A, b, c, out, stable: bool) -> np.ndarray:
    shape = _check_shape(shape)
    A = _c128(A)
    b = _c128(b)
    D = b.shape[-1]                      # core.py:66 `D = b.shape[-1]`
    if len(shape) != D:
        raise ValueError(f"len(shape)={len(shape)} must equal b.shape[-1]={D}")
    A = A.reshape(D, D)
    b = b.reshape(D)
    c = _c128(c, (1,))
    n = int(np.prod(shape, dtype=np.int64))
    if out is None:
        G = _lib.pinned_empty(shape)
    else:
        _check_out(out, n)
        G = out
    check(lib.mmh_forward_host(D, shape_array(shape), _p(A), _p(b), _p(c), _p(G), int(bool(stable))))
    if out is not None:
        return out if out.shape == shape else out.reshape(shape)
    return G


def vanilla_numba(shape, A, b, c, out=None) -> np.ndarray:
    """Fock lattice of the Bargmann triple, first-non-zero-index pivot (vanilla/core.py:25-124)."""
    return _forward(shape, A, b, c, out, False)


def stable_numba(shape, A, b, c, out=None) -> np.ndarray:
    """Same, averaged over all valid pivots (vanilla/core.py:127-213)."""
    return _forward(shape, A, b, c, out, True)


vanilla = vanilla_numba
stable = stable_numba


def vanilla_batch_numba(shape, A, b, c, stable: bool = False, out=None) -> np.ndarray:
    """Batch of independent lattices, batch on the first axis (vanilla/batch.py:27-61)."""
    shape = _check_shape(shape)
    b = _c128(b)
    if b.ndim != 2:
        raise ValueError("b must have shape (batch, D)")
    B, D = b.shape                       # batch.py:52 `batch_size = b.shape[0]`
    if len(shape) != D:
        raise ValueError(f"len(shape)={len(shape)} must equal b.shape[-1]={D}")
    A = _c128(A)
    if A.shape != (B, D, D):
        A = np.ascontiguousarray(np.broadcast_to(A, (B, D, D)))
    c = _c128(c)
    if c.shape != (B,):
        c = np.ascontiguousarray(np.broadcast_to(c, (B,)))
    n = int(np.prod(shape, dtype=np.int64))
    if out is None:
        G = _lib.pinned_empty((B, *shape))
    else:
        _check_out(out, B * n)
        G = out
    check(lib.mmh_forward_batched_host(B, D, shape_array(shape), _p(A), _p(b), _p(c), _p(G), int(bool(stable))))
    return G


def vanilla_vjp_numba(G, c, dLdG):
    """(dL/dA, dL/db, dL/dc) from the forward lattice and its cotangent (vanilla/gradients.py:25-82)."""
    G = _c128(G)
    dLdG = _c128(dLdG)
    if dLdG.shape != G.shape:
        raise ValueError(f"dLdG.shape={dLdG.shape} must equal G.shape={G.shape}")
    D = G.ndim
    c = _c128(c, (1,))
    dA = np.empty((D, D), np.complex128)
    db = np.empty((D,), np.complex128)
    dc = np.empty((1,), np.complex128)
    check(lib.mmh_vjp_host(D, shape_array(G.shape), _p(G), _p(c), _p(dLdG), _p(dA), _p(db), _p(dc)))
    return dA, db, complex(dc[0])


def vanilla_batch_vjp_numba(G, c, dLdG):
    """Per-triple VJP, batch on the first axis (vanilla/gradients.py:85-116)."""
    G = _c128(G)
    dLdG = _c128(dLdG)
    if dLdG.shape != G.shape:
        raise ValueError(f"dLdG.shape={dLdG.shape} must equal G.shape={G.shape}")
    B = G.shape[0]
    D = G.ndim - 1
    c = _c128(c, (B,))
    dA = np.empty((B, D, D), np.complex128)
    db = np.empty((B, D), np.complex128)
    dc = np.empty((B,), np.complex128)
    if B:
        check(lib.mmh_vjp_batched_host(B, D, shape_array(G.shape[1:]), _p(G), _p(c), _p(dLdG), _p(dA), _p(db), _p(dc)))
    return dA, db, dc


def binomial(local_cutoffs, A, b, c, max_l2, global_cutoff):
    """Fill by total photon number with early stop; returns (G, norm) (strategies/binomial.py:30-72)."""
    shape = _check_shape(local_cutoffs)
    D = len(shape)
    A = _c128(A, (D, D))
    b = _c128(b, (D,))
    c = _c128(c, (1,))
    G = _lib.pinned_empty(shape)
    norm = ctypes.c_double(0.0)
    max_l2 = float("inf") if max_l2 is None else float(max_l2)   # binomial.py:68-69 (TypeError -> never stop)
    check(lib.mmh_binomial_host(D, shape_array(shape), _p(A), _p(b), _p(c), max_l2, int(global_cutoff), _p(G),
                                ctypes.byref(norm)))
    return G, norm.value


# ---- compactFock: diagonal / one leftover mode -------------------------------------------------------------
def _input_validation(A, rtol=1e-05, atol=1e-08):
    """compactFock/inputValidation.py:29-56 (same exception types and messages)."""
    if not isinstance(A, np.ndarray):
        raise TypeError("Input matrix must be a NumPy array.")
    n = A.shape
    if n[0] != n[1]:
        raise ValueError("Input matrix must be square.")
    if np.isnan(A).any():
        raise ValueError("Input matrix must not contain NaNs.")
    if not np.allclose(A, A.T, rtol=rtol, atol=atol):
        raise ValueError("Input matrix must be symmetric.")
    return True


def hermite_multidimensional_diagonal(A, B, G0, cutoffs, rtol=1e-05, atol=1e-08):
    """PNR-diagonal amplitudes G[a,a,b,b,...]; A, B in interleaved order (compactFock/inputValidation.py:61-79).

    Returns arr0 only (the reference returns the tuple (arr0, arr2, arr1010, arr1001, arr1) and every caller keeps
    [0]); B may be (2M,) or (2M, batch) with the batch on the LAST axis of B and of the result."""
    _input_validation(A, atol=atol, rtol=rtol)
    B = np.asarray(B)
    if B.ndim > 2:
        raise ValueError("B should be either unbactched or two dimensional (vector and batch dimension)")
    if A.shape[0] != B.shape[0]:
        raise ValueError("The matrix A and vector B have incompatible dimensions")
    try:
        cutoffs = tuple(int(c) for c in cutoffs)
    except TypeError:
        raise ValueError("cutoffs should be array like of length M") from None
    M = len(cutoffs)
    if A.shape[0] // 2 != M:
        raise ValueError("The matrix A and cutoffs have incompatible dimensions")
    cutoffs = _check_shape(cutoffs)
    A = _c128(A)
    B = _c128(B)
    G0 = _c128(G0, (1,))
    nb = 0 if B.ndim == 1 else B.shape[1]
    out = _lib.pinned_empty(cutoffs + ((nb,) if B.ndim == 2 else ()))
    if out.size:
        check(lib.mmh_diagonal_host(M, shape_array(cutoffs), _p(A), _p(B), nb, _p(G0), _p(out)))
    return out


def hermite_multidimensional_1leftoverMode(A, B, G0, cutoffs, rtol=1e-05, atol=1e-08):
    """Density matrix of the first (undetected) mode for every PNR pattern of the others; A, B interleaved
    (compactFock/inputValidation.py:103-122).  Returns arr0[c0, c0, *cutoffs[1:]]."""
    _input_validation(A, atol=atol, rtol=rtol)
    B = np.asarray(B)
    if A.shape[0] != B.shape[0]:
        raise ValueError("The matrix A and vector B have incompatible dimensions")
    try:
        cutoffs = tuple(int(c) for c in cutoffs)
    except TypeError:
        raise ValueError("cutoffs should be array like of length M") from None
    M = len(cutoffs)
    if A.shape[0] // 2 != M:
        raise ValueError("The matrix A and cutoffs have incompatible dimensions")
    if M <= 1:
        raise ValueError("The number of modes should be greater than 1.")
    cutoffs = _check_shape(cutoffs)
    A = _c128(A)
    B = _c128(B, (2 * M,))
    G0 = _c128(G0, (1,))
    out = _lib.pinned_empty((cutoffs[0], cutoffs[0]) + cutoffs[1:])
    check(lib.mmh_1leftover_host(M, shape_array(cutoffs), _p(A), _p(B), _p(G0), _p(out)))
    return out


def fast_diagonal(A, b, c, output_cutoff, pnr_cutoffs, stable=False):
    """Conditional density matrices, output [*(pnr+1), out+1, out+1]; A, b in bargmann order [m0.. | m0..]
    (strategies/fast_diagonal.py:32-77).  `stable` only changed the rounding of the reference's seed block and is
    accepted for call compatibility.
    Deviation from the reference (stated by tests/test_gpu_diagonal.py::test_fast_diagonal_deviation): its weight loop
    `range(1, 2*output_cutoff + 2*sum(pnr_cutoffs) - L)` (fast_diagonal.py:68, L = number of modes) ends before the top weight
    2*sum(pnr_cutoffs) whenever 2*output_cutoff - L < 1 and leaves the highest-weight conditional density matrices ZERO; this
    implementation returns the true amplitudes there (equal to the compactFock path and to the diagonal of the vanilla lattice)."""
    pnr_cutoffs = tuple(int(p) for p in pnr_cutoffs)
    L = len(pnr_cutoffs) + 1
    perm = [i for m in range(L) for i in (m, m + L)]
    A = np.asarray(A)[perm, :][:, perm]
    b = np.asarray(b)[perm]
    cut = (int(output_cutoff) + 1,) + tuple(p + 1 for p in pnr_cutoffs)
    out = hermite_multidimensional_1leftoverMode(np.ascontiguousarray(A), b, c, cut)
    return out.transpose(tuple(range(2, 2 + L - 1)) + (0, 1))


def grad_hermite_multidimensional_diagonal(A, B, G0, arr0, arr2=None, arr1010=None, arr1001=None, arr1=None):
    """Jacobians (arr0_dG0, arr0_dA, arr0_dB) of the diagonal amplitudes (compactFock/inputValidation.py:82-100,
    diagonal_grad.py:261-354).  The reference takes the five forward arrays; only arr0.shape (= cutoffs) is used here,
    the forward sweep is recomputed on the device.  Entries of A are treated as independent (no symmetrisation)."""
    A = _c128(A)
    B = _c128(B)
    if A.shape[0] != B.shape[0]:
        raise ValueError("The matrix A and vector B have incompatible dimensions")
    if B.ndim != 1:
        raise ValueError("B batched")          # the reference's jax bwd rule raises the same (jax_vjps/hermite.py:297-298)
    cutoffs = tuple(int(s) for s in np.shape(arr0))
    M = A.shape[0] // 2
    if len(cutoffs) != M:
        raise ValueError("The matrix A and cutoffs have incompatible dimensions")
    G0 = _c128(G0, (1,))
    dG0 = np.empty(cutoffs, np.complex128)
    dA = np.empty(cutoffs + (2 * M, 2 * M), np.complex128)
    dB = np.empty(cutoffs + (2 * M,), np.complex128)
    check(lib.mmh_diagonal_grad_host(M, shape_array(cutoffs), _p(A), _p(B), _p(G0), _p(dG0), _p(dA), _p(dB)))
    return dG0, dA, dB


def hermite_renormalized_diagonal_vjp(A, B, G0, cutoffs, dLdpoly):
    """The contraction the reference's jax bwd rule performs (jax_vjps/hermite.py:324-329): (dLdA, dLdB, dLdC) from the
    cotangent of arr0.  A, B interleaved."""
    cutoffs = tuple(int(c) for c in cutoffs)
    dG0, dA, dB = grad_hermite_multidimensional_diagonal(A, B, G0, np.empty(cutoffs, np.complex128))
    g = np.asarray(dLdpoly)
    ax = tuple(range(g.ndim))
    return np.sum(g[..., None, None] * dA, axis=ax), np.sum(g[..., None] * dB, axis=ax), np.sum(g * dG0, axis=ax)


def grad_hermite_multidimensional_1leftoverMode(A, B, G0, arr0, arr2=None, arr1010=None, arr1001=None, arr1=None):
    """Jacobians (arr0_dG0, arr0_dA, arr0_dB) of the one-leftover-mode amplitudes (compactFock/inputValidation.py:125-142,
    singleLeftoverMode_grad.py:560-724).  Only arr0.shape = (c0, c0, *cutoffs_tail) is used; the forward sweep is recomputed."""
    A = _c128(A)
    B = _c128(B)
    if A.shape[0] != B.shape[0]:
        raise ValueError("The matrix A and vector B have incompatible dimensions")
    M = A.shape[0] // 2
    if M <= 1:
        raise ValueError("The number of modes should be greater than 1.")
    shp = tuple(int(s) for s in np.shape(arr0))
    if len(shp) != M + 1 or shp[0] != shp[1]:
        raise ValueError("arr0 must have shape (c0, c0, *cutoffs_tail)")
    cutoffs = shp[1:]
    G0 = _c128(G0, (1,))
    dG0 = np.empty(shp, np.complex128)
    dA = np.empty(shp + (2 * M, 2 * M), np.complex128)
    dB = np.empty(shp + (2 * M,), np.complex128)
    check(lib.mmh_1leftover_grad_host(M, shape_array(cutoffs), _p(A), _p(B), _p(G0), _p(dG0), _p(dA), _p(dB)))
    return dG0, dA, dB


def vanilla_contract_numba(shape, shape_derived, A, b, c_poly, stable=False) -> np.ndarray:
    """Lattice with vacuum amplitude 1 over `shape + shape_derived`, contracted over the derived axes with the polynomial
    coefficients `c_poly`: the fused form of
        G = hermite_renormalized(A, b, ones, shape + shape_derived); einsum("...k,...k->...", G.reshape(.., -1), c.reshape(.., -1))
    in CircuitComponent.fock_array (lab/circuit_components.py:516-530) and PolyExpAnsatz.decompose_ansatz
    (physics/ansatz/polyexp_ansatz.py:447-462).  Unbatched (A[D,D], b[D], c_poly[*shape_derived]) or batched on the first
    axis (A[B,D,D], b[B,D], c_poly[B,*shape_derived]).  The lattice never leaves the device."""
    shape = _check_shape(shape)
    shape_derived = _check_shape(shape_derived)
    full = shape + shape_derived
    D = len(full)
    A = _c128(A)
    b = _c128(b)
    if b.shape[-1] != D:
        raise ValueError(f"len(shape + shape_derived)={D} must equal b.shape[-1]={b.shape[-1]}")
    batched = b.ndim == 2
    B = b.shape[0] if batched else 1
    nd = int(np.prod(shape_derived, dtype=np.int64))
    A = A.reshape(B, D, D)
    b = b.reshape(B, D)
    c_poly = _c128(c_poly)
    if c_poly.size != B * nd:
        raise ValueError(f"c_poly has {c_poly.size} entries, expected batch x prod(shape_derived) = {B * nd}")
    c_poly = c_poly.reshape(B, nd)
    out = _lib.pinned_empty((B, *shape))
    check(lib.mmh_forward_contract_host(B, D, shape_array(full), len(shape), _p(A), _p(b), _p(c_poly), _p(out),
                                        int(bool(stable))))
    return out if batched else out.reshape(shape)


# ---- gate-specific strategies -------------------------------------------------------------
def _gate(what: int, shape, a0: float, a1: float) -> np.ndarray:
    shape = _check_shape(shape)
    out = _lib.pinned_empty(shape)
    check(lib.mmh_gate_host(what, shape_array(shape), float(a0), float(a1), _p(out)))
    return out


def squeezer(shape, r, theta, dtype=np.complex128) -> np.ndarray:
    """Fock matrix of the squeezing gate, S[out, in] (strategies/squeezer.py:29-66).  Bit-identical to the numba strategy."""
    if len(tuple(shape)) != 2:
        raise ValueError("squeezer expects shape = (M, N)")
    return _gate(0, shape, r, theta)


def squeezed(cutoff, r, theta, dtype=np.complex128) -> np.ndarray:
    """Fock amplitudes of the single-mode squeezed vacuum (strategies/squeezer.py:127-147)."""
    return _gate(1, (int(cutoff),), r, theta)


def beamsplitter(shape, theta, phi, dtype=np.complex128) -> np.ndarray:
    """Fock tensor G[out0, out1, in0, in1] of the beamsplitter (strategies/beamsplitter.py:37-91), including the reference's
    `for n in range(N - m)` bound of the q = 0 face (:67), which leaves the face entries with out0 + out1 >= N zero."""
    if len(tuple(shape)) != 4:
        raise ValueError("beamsplitter expects shape = (M, N, P, Q)")
    return _gate(2, shape, theta, phi)


def stable_beamsplitter(shape, theta, phi) -> np.ndarray:
    """All-pivot average of the beamsplitter recurrence (strategies/beamsplitter.py:94-172)."""
    if len(tuple(shape)) != 4:
        raise ValueError("stable_beamsplitter expects shape = (M, N, P, Q)")
    return _gate(3, shape, theta, phi)


def displacement(cutoffs, alpha, dtype=np.complex128) -> np.ndarray:
    """Fock matrix D[out, in] of the displacement gate from the log-domain Laguerre form (strategies/displacement.py:24-65)."""
    cutoffs = tuple(int(c) for c in cutoffs)
    if len(cutoffs) != 2:
        raise ValueError("displacement expects cutoffs = (N, M)")
    alpha = complex(alpha)
    return _gate(4, cutoffs, alpha.real, alpha.imag)


def _disp_derivs(what: int, D, a0: float, a1: float):
    D = _c128(D)
    if D.ndim != 2:
        raise ValueError("expected a 2-dimensional gate array")
    o1 = np.empty(D.shape, np.complex128)
    o2 = np.empty(D.shape, np.complex128)
    check(lib.mmh_displacement_derivs_host(what, D.shape[0], D.shape[1], _p(D), float(a0), float(a1), _p(o1), _p(o2)))
    return o1, o2


def jacobian_displacement(D, alpha):
    """(dD/dalpha, dD/dconj(alpha)) of the displacement gate (strategies/displacement.py:117-139)."""
    alpha = complex(alpha)
    return _disp_derivs(0, D, alpha.real, alpha.imag)


def grad_displacement(T, r, phi):
    """(dT/dr, dT/dphi) of the displacement gate (strategies/displacement.py:85-114); T square."""
    T = _c128(T)
    if T.ndim != 2 or T.shape[0] != T.shape[1]:
        raise ValueError("grad_displacement expects a square gate array")
    return _disp_derivs(1, T, r, phi)


def _gate_vjp(kind: int, G, dLdG):
    """Sums over the gate's support of dLdG[k] * vanilla_step_grad(G, k) (lattice/steps.py:145-171):
    (dLdA upper-triangular [D, D], dLdb [D], sum(G * dLdG))."""
    G = _c128(G)
    dLdG = _c128(dLdG)
    if dLdG.shape != G.shape:
        raise ValueError(f"dLdG.shape={dLdG.shape} must equal G.shape={G.shape}")
    D = G.ndim
    out = np.empty(D * D + D + 1, np.complex128)
    check(lib.mmh_gate_vjp_host(kind, D, shape_array(G.shape), _p(G), _p(dLdG), _p(out)))
    return out[: D * D].reshape(D, D), out[D * D: D * D + D], complex(out[-1])


def beamsplitter_vjp(G, dLdG, theta, phi):
    """(dL/dtheta, dL/dphi) (strategies/beamsplitter.py:175-243): lattice reduction on the GPU, then the chain rule through
    dV/dtheta, dV/dphi of the 2x2 beamsplitter unitary (:232-241)."""
    dLdA, _, _ = _gate_vjp(0, G, dLdG)
    st, ct = np.sin(theta), np.cos(theta)
    e, em = np.exp(1j * phi), np.exp(-1j * phi)
    dLdtheta = 2 * np.real(-st * dLdA[0, 2] - ct * em * dLdA[0, 3] + ct * e * dLdA[1, 2] - st * dLdA[1, 3])
    dLdphi = 2 * np.real(1j * st * em * dLdA[0, 3] + 1j * st * e * dLdA[1, 2])
    return dLdtheta, dLdphi


def squeezer_vjp(G, dLdG, r, phi):
    """(dL/dr, dL/dphi) of the squeezing gate (strategies/squeezer.py:69-124)."""
    dLdA, _, dLdC = _gate_vjp(1, G, dLdG)
    d_sech = -np.tanh(r) / np.cosh(r)
    d_tanh = 1.0 / np.cosh(r) ** 2
    tanh = np.tanh(r)
    e, ec = np.exp(1j * phi), np.exp(-1j * phi)
    dLdr = 2 * np.real(-dLdA[0, 0] * e * d_tanh + dLdA[0, 1] * d_sech + dLdA[1, 1] * ec * d_tanh - np.conj(dLdC) * 0.5 * tanh)
    dLdphi = 2 * np.real(-dLdA[0, 0] * 1j * e * tanh - dLdA[1, 1] * 1j * ec * tanh)
    return dLdr, dLdphi


def squeezed_vjp(G, dLdG, r, phi):
    """(dL/dr, dL/dphi) of the squeezed vacuum ket (strategies/squeezer.py:150-191)."""
    dLdA, _, dLdC = _gate_vjp(2, G, dLdG)
    tanh = np.tanh(r)
    d_tanh = 1.0 / np.cosh(r) ** 2
    e = np.exp(1j * phi)
    dLdr = 2 * np.real(-dLdA[0, 0] * e * d_tanh - np.conj(dLdC) * 0.5 * tanh)
    dLdphi = 2 * np.real(-dLdA[0, 0] * 1j * e * tanh)
    return dLdr, dLdphi


# ---- autoshape -------------------------------------------------------------------------------
def autoshape_numba(A, b, c, max_prob, max_shape, min_shape) -> np.ndarray:
    """Fock shape of a Gaussian density matrix such that every single-mode marginal keeps max_prob of its trace
    (math/lattice/autoshape.py:24-154): A[2M,2M], b[2M] in bargmann order, c scalar -> int64[M] clipped to [min_shape, max_shape]."""
    b = _c128(b)
    if b.ndim != 1 or b.shape[0] % 2:
        raise ValueError("b must be a vector of even length 2M")
    M = b.shape[0] // 2
    A = _c128(A, (2 * M, 2 * M))
    c = _c128(c, (1,))
    out = np.empty(M, np.int64)
    check(lib.mmh_autoshape_host(M, _p(A), _p(b), _p(c), float(max_prob), int(max_shape), int(min_shape),
                                 out.ctypes.data_as(ctypes.POINTER(ctypes.c_int64))))
    return out
