"""CPU oracle for the compactFock "diagonal" and "1 leftover mode" paths — TEST INFRASTRUCTURE, not product code.

A plain numpy/Python restatement (per-`params` loops, small cases only) of
  mrmustard/math/lattice/strategies/compactFock/diagonal_amps.py:19-248          (`diagonal`)
  mrmustard/math/lattice/strategies/compactFock/singleLeftoverMode_amps.py:21-475 (`leftover`)
  mrmustard/math/lattice/strategies/fast_diagonal.py:32-77 output convention      (`fast_diagonal`)
Parity status: pinned against golden vectors generated from the unmodified reference
(tests/golden/gen_golden_diagonal.py -> tests/golden/diagonal_golden.npz); tolerance 1e-10 rel / 1e-14 abs
because the reference evaluates `A[i] @ G_in` through BLAS, whose summation order is unspecified.
Only tests/ may import this module.
"""
from __future__ import annotations

import itertools

import numpy as np


def _levels(cutoffs):
    """helperFunctions.construct_dict_params (helperFunctions.py:31-45): params grouped by their sum, ndindex order."""
    lv = {w: [] for w in range(sum(cutoffs))}
    for params in np.ndindex(*cutoffs):
        lv[sum(params)].append(params)
    return lv


def _sub(params, i):
    p = list(params); p[i] -= 1
    return tuple(p)


def _add(params, i):
    p = list(params); p[i] += 1
    return tuple(p)


def diagonal(A, B, G0, cutoffs):
    """fock_representation_diagonal_amps (diagonal_amps.py:200-248) -> arr0, shape cutoffs (+ batch when B.ndim == 2).

    A, B are in the interleaved order [m0, m0, m1, m1, ...] (after reorder_AB_bargmann, backend_numpy.py:367-377)."""
    A = np.asarray(A, dtype=np.complex128)
    B = np.asarray(B, dtype=np.complex128)
    cutoffs = tuple(int(c) for c in cutoffs)
    M = len(cutoffs)
    tail = () if B.ndim == 1 else (B.shape[1],)
    arr0 = np.zeros(cutoffs + tail, np.complex128)
    arr2 = np.zeros((M, *cutoffs) + tail, np.complex128)
    arr1 = np.zeros((2 * M, *cutoffs) + tail, np.complex128)
    arr1010 = np.zeros((M, max(M - 1, 1), *cutoffs) + tail, np.complex128)
    arr1001 = np.zeros((M, max(M - 1, 1), *cutoffs) + tail, np.complex128)
    arr0[(0,) * M] = G0
    sqrt = np.sqrt
    ex = (lambda k: k) if B.ndim == 1 else (lambda k: np.expand_dims(k, 1))
    for w, plist in _levels(cutoffs).items():
        for params in plist:
            rep = np.repeat(np.array(params), 2)
            # ---- diagonal pivot (diagonal_amps.py:98-141, visited per :179)
            if cutoffs[0] == 1 or params[0] < cutoffs[0] - 1:
                K_l, K_i = sqrt(rep), sqrt(rep + 1)
                G_in = np.zeros((2 * M,) + tail, np.complex128)
                GB = arr0[params] * B if B.ndim == 1 else arr0[params][None, :] * B
                for i in range(2 * M):
                    if params[i // 2] > 0:
                        G_in[i] = arr1[(i + 1 - 2 * (i % 2), *_sub(params, i // 2))]
                G_in = ex(K_l) * G_in
                for i in range(2 * M):
                    if params[i // 2] + 1 < cutoffs[i // 2] and (i != 1 or params[0] + 2 < cutoffs[0]):
                        arr1[(i, *params)] = (GB[i] + A[i] @ G_in) / K_i[i]
            # ---- off-diagonal pivots (diagonal_amps.py:19-94, visited per :183)
            for d in range(M):
                if all(p == 0 for p in params[:d]) and params[d] < cutoffs[d] - 1:
                    pivot = rep.copy(); pivot[2 * d] += 1
                    K_l, K_i = sqrt(pivot), sqrt(pivot + 1)
                    G_in = np.zeros((2 * M,) + tail, np.complex128)
                    a1 = arr1[(2 * d, *params)]
                    GB = a1 * B if B.ndim == 1 else a1[None, :] * B
                    G_in[2 * d] = arr0[params]
                    if params[d] > 0:
                        G_in[2 * d + 1] = arr2[(d, *_sub(params, d))]
                    for i in range(d + 1, M):
                        if params[i] > 0:
                            G_in[2 * i] = arr1001[(d, i - d - 1, *_sub(params, i))]
                            G_in[2 * i + 1] = arr1010[(d, i - d - 1, *_sub(params, i))]
                    G_in = ex(K_l) * G_in
                    arr0[_add(params, d)] = (GB[2 * d + 1] + A[2 * d + 1] @ G_in) / K_i[2 * d + 1]
                    if params[d] + 2 < cutoffs[d]:
                        arr2[(d, *params)] = (GB[2 * d] + A[2 * d] @ G_in) / K_i[2 * d]
                    for i in range(d + 1, M):
                        if params[i] + 1 < cutoffs[i]:
                            arr1010[(d, i - d - 1, *params)] = (GB[2 * i] + A[2 * i] @ G_in) / K_i[2 * i]
                            arr1001[(d, i - d - 1, *params)] = (GB[2 * i + 1] + A[2 * i + 1] @ G_in) / K_i[2 * i + 1]
    return arr0


def leftover(A, B, G0, cutoffs):
    """fock_representation_1leftoverMode_amps (singleLeftoverMode_amps.py:424-475) -> arr0[c0, c0, *cutoffs_tail].

    A, B interleaved; indices 0, 1 belong to the undetected mode."""
    A = np.asarray(A, dtype=np.complex128)
    B = np.asarray(B, dtype=np.complex128)
    cutoffs = tuple(int(c) for c in cutoffs)
    M = len(cutoffs)
    c0, ct = cutoffs[0], cutoffs[1:]
    Md = M - 1
    z = (0,) * Md
    arr0 = np.zeros((c0, c0) + ct, np.complex128)
    arr2 = np.zeros((c0, c0, Md) + ct, np.complex128)
    arr1 = np.zeros((c0, c0, 2 * Md) + ct, np.complex128)
    arr1010 = np.zeros((c0, c0, Md, max(Md - 1, 1)) + ct, np.complex128)
    arr1001 = np.zeros((c0, c0, Md, max(Md - 1, 1)) + ct, np.complex128)
    arr0[(0, 0) + z] = G0
    sqrt = np.sqrt
    # seed block (singleLeftoverMode_amps.py:324-334)
    for m in range(c0 - 1):
        prev = arr0[(m - 1, 0) + z] if m > 0 else 0.0
        arr0[(m + 1, 0) + z] = (arr0[(m, 0) + z] * B[0] + sqrt(m) * A[0, 0] * prev) / sqrt(m + 1)
    for m in range(c0):
        for n in range(c0 - 1):
            pm = arr0[(m - 1, n) + z] if m > 0 else 0.0
            pn = arr0[(m, n - 1) + z] if n > 0 else 0.0
            arr0[(m, n + 1) + z] = (arr0[(m, n) + z] * B[1] + sqrt(m) * A[1, 0] * pm + sqrt(n) * A[1, 1] * pn) / sqrt(n + 1)

    def write_block(i, arr_write, write, arr_pivot, read_GB, G_in, K_i):
        # singleLeftoverMode_amps.py:21-74: i indexes the full (2M) A; K_i is indexed by i - 2
        for m in range(c0):
            for n in range(c0):
                v = arr_pivot[(m, n) + read_GB] * B[i]
                if m > 0:
                    v = v + A[i, 0] * (arr_pivot[(m - 1, n) + read_GB] * sqrt(m))
                if n > 0:
                    v = v + A[i, 1] * (arr_pivot[(m, n - 1) + read_GB] * sqrt(n))
                v = v + A[i, 2:] @ G_in[m, n]
                arr_write[(m, n) + write] = v / K_i[i - 2]

    for w, plist in _levels(ct).items():
        for params in plist:
            rep = np.repeat(np.array(params), 2)
            if ct[0] == 1 or params[0] < ct[0] - 1:                      # diag pivot (:225-287)
                K_l, K_i = sqrt(rep), sqrt(rep + 1)
                G_in = np.zeros((c0, c0, 2 * Md), np.complex128)
                for i in range(2 * Md):
                    if params[i // 2] > 0:
                        G_in[:, :, i] = arr1[(slice(None), slice(None), i + 1 - 2 * (i % 2)) + _sub(params, i // 2)]
                G_in = G_in * K_l
                for i in range(2 * Md):
                    if params[i // 2] + 1 < ct[i // 2] and (i != 1 or params[0] + 2 < ct[0]):
                        write_block(i + 2, arr1, (i,) + params, arr0, params, G_in, K_i)
            for d in range(Md):                                            # off-diag pivots (:101-222)
                if all(p == 0 for p in params[:d]) and params[d] < ct[d] - 1:
                    pivot = rep.copy(); pivot[2 * d] += 1
                    K_l, K_i = sqrt(pivot), sqrt(pivot + 1)
                    G_in = np.zeros((c0, c0, 2 * Md), np.complex128)
                    G_in[:, :, 2 * d] = arr0[(slice(None), slice(None)) + params]
                    if params[d] > 0:
                        G_in[:, :, 2 * d + 1] = arr2[(slice(None), slice(None), d) + _sub(params, d)]
                    for i in range(d + 1, Md):
                        if params[i] > 0:
                            G_in[:, :, 2 * i] = arr1001[(slice(None), slice(None), d, i - d - 1) + _sub(params, i)]
                            G_in[:, :, 2 * i + 1] = arr1010[(slice(None), slice(None), d, i - d - 1) + _sub(params, i)]
                    G_in = G_in * K_l
                    read_GB = (2 * d,) + params
                    write_block(2 * d + 3, arr0, _add(params, d), arr1, read_GB, G_in, K_i)
                    if params[d] + 2 < ct[d]:
                        write_block(2 * d + 2, arr2, (d,) + params, arr1, read_GB, G_in, K_i)
                    for i in range(d + 1, Md):
                        if params[i] + 1 < ct[i]:
                            write_block(2 * i + 2, arr1010, (d, i - d - 1) + params, arr1, read_GB, G_in, K_i)
                            write_block(2 * i + 3, arr1001, (d, i - d - 1) + params, arr1, read_GB, G_in, K_i)
    return arr0


def reorder_AB_bargmann(A, B):
    """backend_numpy.py:367-377: [m0..,m0..] -> [m0,m0,m1,m1,..]."""
    A = np.asarray(A); B = np.asarray(B)
    ordering = np.arange(2 * A.shape[0] // 2).reshape(2, -1).T.flatten()
    return A[ordering][:, ordering], B[ordering]


def fast_diagonal(A, b, c, output_cutoff, pnr_cutoffs):
    """Output convention of strategies.fast_diagonal (fast_diagonal.py:32-77): [*(pnr+1), out+1, out+1]; A, b in
    bargmann order [m0..mL-1 | m0..mL-1]."""
    A2, b2 = reorder_AB_bargmann(A, b)
    cut = (output_cutoff + 1,) + tuple(p + 1 for p in pnr_cutoffs)
    out = leftover(A2, b2, c, cut)
    L1 = len(pnr_cutoffs)
    return out.transpose(tuple(range(2, 2 + L1)) + (0, 1))
