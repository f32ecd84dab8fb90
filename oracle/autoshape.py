"""CPU restatement of the reference's autoshape_numba — TEST INFRASTRUCTURE, not product code.

numpy restatement of mrmustard/math/lattice/autoshape.py:24-154 in the form the CUDA kernel uses (one linear solve per mode for
the Schur complement instead of an explicit inverse; the same two-buffer diagonal recurrence with early stop).  Parity status:
pinned against golden vectors from the unmodified reference (tests/golden/gen_golden_autoshape.py) by
tests/test_oracle_autoshape.py -- the output is integer-valued and must be identical.
"""
from __future__ import annotations

import numpy as np

SQRT = np.sqrt(np.arange(100000))


def autoshape(A, b, c, max_prob, max_shape, min_shape):
    A = np.asarray(A, dtype=np.complex128)
    b = np.asarray(b, dtype=np.complex128)
    c = complex(np.asarray(c).reshape(()))
    M = b.shape[0] // 2
    shape = np.ones(M, dtype=np.int64)
    n = 2 * M - 2
    X = np.zeros((n, n), dtype=np.complex128)
    for i in range(M - 1):
        X[i, i + M - 1] = X[i + M - 1, i] = 1.0
    for m in range(M):                                           # autoshape.py:101
        rest = [i for i in range(M) if i != m]
        full_m = [m, M + m]
        full_n = rest + [M + i for i in rest]                    # variable s (M - 1) + i'  <->  s M + rest[i']
        A_mm = A[np.ix_(full_m, full_m)]
        A_nn = A[np.ix_(full_n, full_n)]
        A_mn = A[np.ix_(full_m, full_n)]
        b_m, b_n = b[full_m], b[full_n]
        if n:
            Z = np.linalg.solve(A_nn - X, np.concatenate([A_mn.T, b_n[:, None]], axis=1))
            A_ = A_mm - A_mn @ Z[:, :2]                          # :116
            b_ = b_m - A_mn @ Z[:, 2]                            # :117
            c_ = c * np.exp(-0.5 * b_n @ Z[:, 2]) / np.sqrt(np.linalg.det(A_nn - X))   # :118-122
        else:
            A_, b_, c_ = A_mm, b_m, c
        buf2 = np.zeros((2, 2), dtype=np.complex128)
        buf3 = np.zeros((2, 3), dtype=np.complex128)
        buf3[0, 1] = c_
        norm = abs(c_)
        k = 0
        while norm < max_prob and k < max_shape:                 # :129-151
            p, q = k % 2, (k + 1) % 2
            buf2[q] = (b_ * buf3[p, 1] + A_ @ buf2[p] * SQRT[k]) / SQRT[k + 1]
            buf3[q, 0] = (b_[0] * buf2[q, 0] + A_[0, 0] * buf3[p, 1] * SQRT[k + 1] + A_[0, 1] * buf3[p, 0] * SQRT[k]) / SQRT[k + 2]
            buf3[q, 1] = (b_[1] * buf2[q, 0] + A_[1, 0] * buf3[p, 1] * SQRT[k + 1] + A_[1, 1] * buf3[p, 0] * SQRT[k]) / SQRT[k + 1]
            buf3[q, 2] = (b_[1] * buf2[q, 1] + A_[1, 0] * buf3[p, 2] * SQRT[k] + A_[1, 1] * buf3[p, 1] * SQRT[k + 1]) / SQRT[k + 2]
            norm += abs(buf3[q, 1])
            k += 1
        shape[m] = k
    return np.clip(shape, min_shape, max_shape)
