"""CPU restatement of the reference's gate-specific Fock strategies — TEST INFRASTRUCTURE, not product code.

numpy restatements (level-synchronous, vectorised over the independent points of a level) of
  displacement / laguerre / jacobian_displacement / grad_displacement   (strategies/displacement.py:24-139)
  squeezer / squeezer_vjp / squeezed / squeezed_vjp                       (strategies/squeezer.py:29-191)
  beamsplitter / stable_beamsplitter / beamsplitter_vjp                   (strategies/beamsplitter.py:37-243)
Parity status: pinned against golden vectors generated from the unmodified reference
(tests/golden/gen_golden_gates.py -> tests/golden/gates_golden.npz) by tests/test_oracle_gates.py: squeezer / squeezed /
beamsplitter / stable_beamsplitter BIT-IDENTICAL to the numba strategies (same IEEE operations in the same order, scalar
transcendentals through libm like numba's lowering, complex / real as a component-wise division), displacement and all
derivatives at 1e-10 rel / 1e-14 abs (per-element log / exp).
"""
from __future__ import annotations

import cmath
import math

import numpy as np

SQRT = np.sqrt(np.arange(100000))


# ---- displacement (displacement.py:24-65) -------------------------------------------------------------------
def laguerre(x, N, alpha):
    """First N generalised Laguerre polynomials L_m^{(alpha)}(x), three-term recurrence (displacement.py:68-82)."""
    L = np.zeros(N, dtype=np.complex128)
    L[0] = 1.0
    for m in range(N - 1):
        L[m + 1] = ((2 * m + 1 + alpha - x) * L[m] - (m + alpha) * (L[m - 1] if m > 0 else 0.0)) / (m + 1)
    return L


def displacement(cutoffs, alpha):
    """D[n, m] on the diagonal n - m = d from the log-domain closed form with L_m^{(d)}(|alpha|^2) (displacement.py:36-65)."""
    r, phi = abs(alpha), np.angle(alpha)
    N, M = cutoffs
    flipped = N < M
    if flipped:
        N, M = M, N
    D = np.zeros((N, M), dtype=np.complex128)
    rng = np.arange(max(*cutoffs)); rng[0] = 1
    logfac = np.cumsum(np.log(rng))
    with np.errstate(divide="ignore"):
        for d in range(N):
            m_max = min(M, N - d)
            logL = np.log(laguerre(r ** 2.0, m_max, d))
            m = np.arange(m_max)
            n = m + d
            sign = np.where(flipped & (n > m) & (d % 2 == 1), -1.0, 1.0)
            cj = np.where(flipped & (n > m), -1.0, 1.0)
            val = sign * np.exp(0.5 * (logfac[m] - logfac[n]) + d * np.log(r) - r ** 2.0 / 2.0 + cj * 1j * phi * d + logL)
            D[n, m] = val
            lo = n < M
            D[m[lo], n[lo]] = (-1.0) ** d * np.conj(val[lo])
    return D.T.copy() if flipped else D


def jacobian_displacement(D, alpha):
    """dD/dalpha, dD/dconj(alpha) (displacement.py:117-139); the reference's wrapped reads at index -1 meet sqrt(0)."""
    M, N = D.shape
    sq = np.sqrt(np.arange(M + N))
    up = np.zeros_like(D); up[1:, :] = D[:-1, :]
    left = np.zeros_like(D); left[:, 1:] = D[:, :-1]
    ja = -0.5 * np.conj(alpha) * D + sq[:M, None] * up
    jac = -0.5 * alpha * D - sq[None, :N] * left
    return ja, jac


def grad_displacement(T, r, phi):
    """dT/dr, dT/dphi (displacement.py:85-114)."""
    c = T.shape[0]
    sq = np.sqrt(np.arange(c))
    ei, eic = np.exp(1j * phi), np.exp(-1j * phi)
    up = np.zeros_like(T); up[1:, :] = T[:-1, :]
    left = np.zeros_like(T); left[:, 1:] = T[:, :-1]
    gr = -r * T + sq[:, None] * ei * up - sq[None, :] * eic * left
    gphi = sq[:, None] * 1j * (r * ei) * up + sq[None, :] * 1j * (r * eic) * left
    return gr, gphi


# ---- squeezer / squeezed (squeezer.py:29-66, :127-147) --------------------------------------------------------
def squeezer(shape, r, theta):
    """S[m, n] = sqrt(n-1)/sqrt(n) e^{-i theta} tanh r S[m, n-2] + sqrt(m)/sqrt(n) sech r S[m-1, n-1] on (m+n) even; every point
    of a level m + n = u depends on level u - 2 only."""
    M, N = shape
    S = np.zeros(shape, dtype=np.complex128)
    # libm scalars (math / cmath), which is what numba lowers np.tanh / np.cosh / np.exp of scalars to: np.tanh differs from libm's
    # tanh in the last bit for some r, and the recurrence amplifies that to 1e-7 at cutoff 60
    et = cmath.exp(1j * theta) * math.tanh(r)
    etc = np.conj(et)
    sech = 1.0 / math.cosh(r)
    S[0, 0] = math.sqrt(sech)
    for m in range(2, M, 2):
        S[m, 0] = -SQRT[m - 1] / SQRT[m] * et * S[m - 2, 0]
    for u in range(2, M + N - 1, 2):
        for m in range(max(0, u - N + 1), min(M, u)):     # n = u - m >= 1
            n = u - m
            v = SQRT[n - 1] / SQRT[n] * etc * (S[m, n - 2] if n >= 2 else 0.0)
            if m >= 1:
                v = v + SQRT[m] / SQRT[n] * sech * S[m - 1, n - 1]
            S[m, n] = v
    return S


def squeezed(cutoff, r, theta):
    S = np.zeros(cutoff, dtype=np.complex128)
    et = cmath.exp(1j * theta) * -math.tanh(r)
    S[0] = math.sqrt(1.0 / math.cosh(r))
    for m in range(2, cutoff, 2):
        S[m] = SQRT[m - 1] / SQRT[m] * et * S[m - 2]
    return S


def _masked_step_grads(G, mask, dLdG):
    """sum over the index set `mask` of dLdG[k] * (dA(k), db(k)), with (dA, db) of steps.vanilla_step_grad (lattice/steps.py:145-171):
    db_i = sqrt(k_i) G[k - e_i]; dA_ii = 0.5 sqrt(k_i (k_i - 1)) G[k - 2 e_i]; dA_ij = sqrt(k_i k_j) G[k - e_i - e_j] (j > i)."""
    D = G.ndim
    dLdA = np.zeros((D, D), dtype=np.complex128)
    dLdb = np.zeros(D, dtype=np.complex128)
    for k in np.argwhere(mask):
        k = tuple(int(x) for x in k)
        g = dLdG[k]
        for i in range(D):
            if k[i] == 0:
                continue
            p = k[:i] + (k[i] - 1,) + k[i + 1:]
            dLdb[i] += np.sqrt(float(k[i])) * G[p] * g
            if k[i] > 1:
                dLdA[i, i] += 0.5 * np.sqrt(float(k[i] * (k[i] - 1))) * G[p[:i] + (p[i] - 1,) + p[i + 1:]] * g
            for j in range(i + 1, D):
                if k[j] > 0:
                    dLdA[i, j] += np.sqrt(float(k[i] * k[j])) * G[p[:j] + (p[j] - 1,) + p[j + 1:]] * g
    return dLdA, dLdb


def squeezer_vjp(G, dLdG, r, phi):
    """(dL/dr, dL/dphi) (squeezer.py:69-124): step gradients over the support (m+n) even, then the chain rule."""
    M, N = G.shape
    mm, nn = np.meshgrid(np.arange(M), np.arange(N), indexing="ij")
    mask = ((mm + nn) % 2 == 0) & ((nn >= 1) | ((mm >= 2) & (nn == 0)))
    dLdA, _ = _masked_step_grads(G, mask, dLdG)
    dLdC = np.sum(G * dLdG)
    d_sech, d_tanh, tanh = -np.tanh(r) / np.cosh(r), 1.0 / np.cosh(r) ** 2, np.tanh(r)
    e, ec = np.exp(1j * phi), np.exp(-1j * phi)
    dLdr = 2 * np.real(-dLdA[0, 0] * e * d_tanh + dLdA[0, 1] * d_sech + dLdA[1, 1] * ec * d_tanh - np.conj(dLdC) * 0.5 * tanh)
    dLdphi = 2 * np.real(-dLdA[0, 0] * 1j * e * tanh - dLdA[1, 1] * 1j * ec * tanh)
    return dLdr, dLdphi


def squeezed_vjp(G, dLdG, r, phi):
    """(dL/dr, dL/dphi) of the squeezed vacuum ket (squeezer.py:150-191)."""
    M = G.shape[0]
    m = np.arange(M)
    mask = (m % 2 == 0) & (m >= 2)
    dLdA, _ = _masked_step_grads(G, mask, dLdG)
    tanh, d_tanh, e = np.tanh(r), 1.0 / np.cosh(r) ** 2, np.exp(1j * phi)
    dLdC = np.sum(G * dLdG)
    dLdr = 2 * np.real(-dLdA[0, 0] * e * d_tanh - np.conj(dLdC) * 0.5 * tanh)
    dLdphi = 2 * np.real(-dLdA[0, 0] * 1j * e * tanh)
    return dLdr, dLdphi


# ---- beamsplitter (beamsplitter.py:37-91, :94-172, :175-243) ---------------------------------------------------
def _cdiv(z, s):
    """complex / real as numba lowers it (component-wise true division); numpy's scalar complex division multiplies by the
    reciprocal of the denominator instead, which differs in the last bit."""
    z = complex(z)
    return complex(z.real / s, z.imag / s)


def beamsplitter(shape, theta, phi):
    """Photon-number conserving fill: first the q = 0 face G[m, n, m+n, 0] level by level in m + n, then for every p the
    entries q = m + n - p >= 1 level by level in m + n (each level reads the previous one only)."""
    ct = math.cos(theta)
    st = math.sin(theta) * cmath.exp(1j * phi)
    stc = np.conj(st)
    M, N, P, Q = shape
    G = np.zeros(shape, dtype=np.complex128)
    G[0, 0, 0, 0] = 1.0
    for p in range(1, min(P, N)):       # `for n in range(N - m)` (beamsplitter.py:67): the face stops at p = m + n < N
        for m in range(0, min(M, p + 1)):
            n = p - m
            v = 0.0
            if m > 0:
                v = v + ct * SQRT[m] / SQRT[p] * G[m - 1, n, p - 1, 0]
            if n > 0:
                v = v + _cdiv(st * SQRT[n], SQRT[p]) * G[m, n - 1, p - 1, 0]
            G[m, n, p, 0] = v
    for L in range(1, M + N - 1):
        for m in range(max(0, L - N + 1), min(M, L + 1)):
            n = L - m
            for p in range(max(0, L - Q + 1), min(P, L)):
                q = L - p
                v = 0.0
                if m > 0:
                    v = v + _cdiv(-stc * SQRT[m], SQRT[q]) * G[m - 1, n, p, q - 1]
                if n > 0:
                    v = v + ct * SQRT[n] / SQRT[q] * G[m, n - 1, p, q - 1]
                G[m, n, p, q] = v
    return G


def stable_beamsplitter(shape, theta, phi):
    """Average over every available pivot (beamsplitter.py:94-172); level m + n reads level m + n - 1 only."""
    ct = np.cos(theta)
    st = np.sin(theta) * np.exp(1j * phi)
    stc = np.conj(st)
    M, N, P, Q = shape
    G = np.zeros(shape, dtype=np.complex128)
    G[0, 0, 0, 0] = 1.0

    def g(m, n, p, q):
        return G[m, n, p, q] if (m >= 0 and n >= 0 and p >= 0 and q >= 0) else 0.0

    for L in range(1, M + N - 1):
        for m in range(max(0, L - N + 1), min(M, L + 1)):
            n = L - m
            for p in range(max(0, L - Q + 1), min(P, L + 1)):
                q = L - p
                val, piv = 0.0, 0
                if q == 0:        # the q = 0 face uses the three pivots m, n, p with the (.., p-1, 0) neighbours only
                    if m > 0:
                        val += ct * SQRT[p] / SQRT[m] * g(m - 1, n, p - 1, 0); piv += 1
                    if n > 0:
                        val += _cdiv(st * SQRT[p], SQRT[n]) * g(m, n - 1, p - 1, 0); piv += 1
                    if p > 0:
                        val += ct * SQRT[m] / SQRT[p] * g(m - 1, n, p - 1, 0) + _cdiv(st * SQRT[n], SQRT[p]) * g(m, n - 1, p - 1, 0); piv += 1
                else:
                    if m > 0:
                        val += ct * SQRT[p] / SQRT[m] * g(m - 1, n, p - 1, q) - _cdiv(stc * SQRT[q], SQRT[m]) * g(m - 1, n, p, q - 1); piv += 1
                    if n > 0:
                        val += _cdiv(st * SQRT[p], SQRT[n]) * g(m, n - 1, p - 1, q) + ct * SQRT[q] / SQRT[n] * g(m, n - 1, p, q - 1); piv += 1
                    if p > 0:
                        val += ct * SQRT[m] / SQRT[p] * g(m - 1, n, p - 1, q) + _cdiv(st * SQRT[n], SQRT[p]) * g(m, n - 1, p - 1, q); piv += 1
                    val += _cdiv(-stc * SQRT[m], SQRT[q]) * g(m - 1, n, p, q - 1) + ct * SQRT[n] / SQRT[q] * g(m, n - 1, p, q - 1); piv += 1
                G[m, n, p, q] = _cdiv(val, piv)
    return G


def beamsplitter_vjp(G, dLdG, theta, phi):
    """(dL/dtheta, dL/dphi) (beamsplitter.py:175-243): step gradients over the conserving support, then the chain rule."""
    M, N, P, Q = G.shape
    mm, nn, pp, qq = np.meshgrid(np.arange(M), np.arange(N), np.arange(P), np.arange(Q), indexing="ij")
    mask = (mm + nn == pp + qq) & (mm + nn > 0)
    dLdA, _ = _masked_step_grads(G, mask, dLdG)
    st, ct = np.sin(theta), np.cos(theta)
    e, em = np.exp(1j * phi), np.exp(-1j * phi)
    dLdtheta = 2 * np.real(-st * dLdA[0, 2] - ct * em * dLdA[0, 3] + ct * e * dLdA[1, 2] - st * dLdA[1, 3])
    dLdphi = 2 * np.real(1j * st * em * dLdA[0, 3] + 1j * st * e * dLdA[1, 2])
    return dLdtheta, dLdphi
