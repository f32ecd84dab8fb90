"""CPU oracle for the Gaussian-to-Fock hot path — TEST INFRASTRUCTURE, not product code.

A plain-C restatement (hermite_oracle.c) of the reference's numba strategies, exposed through ctypes
with the same call signatures as the reference functions it restates (file:line cited per function).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this package.  Parity status: pinned against golden vectors generated from the unmodified reference
(tests/golden/gen_golden.py) — see tests/test_oracle_golden.py.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libmmoracle.so")
_lib = None


def build(force: bool = False) -> str:
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith(".c")]
    stale = force or not os.path.exists(_SO) or any(
        os.path.getmtime(s) > os.path.getmtime(_SO) for s in srcs)
    if stale:
        subprocess.check_call(["make", "-s", "-C", _HERE, "libmmoracle.so"] + (["-B"] if force else []))
    return _SO


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_SO)
    return _lib


def _c128(x, shape=None):
    a = np.ascontiguousarray(np.asarray(x, dtype=np.complex128))
    if shape is not None:
        a = a.reshape(shape)
    return a


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _shape_arr(shape):
    return np.ascontiguousarray(np.asarray(tuple(int(s) for s in shape), dtype=np.int64))


def vanilla(shape, A, b, c, out=None, stable=False):
    """vanilla_numba / stable_numba (vanilla/core.py:25-124, :127-213)."""
    shape = tuple(int(s) for s in shape)
    D = len(shape)
    A = _c128(A, (D, D)); b = _c128(b, (D,)); c = _c128(c, (1,))
    sh = _shape_arr(shape)
    if out is None:
        G = np.empty(shape, dtype=np.complex128)
        zero = 1
    else:
        G = out
        assert G.dtype == np.complex128 and G.flags.c_contiguous
        zero = 0
    fn = lib().mmo_stable if stable else lib().mmo_vanilla
    fn(ctypes.c_int(D), _p(sh), _p(A), _p(b), _p(c), _p(G), ctypes.c_int(zero))
    return G if out is not None else G.reshape(shape)


def stable(shape, A, b, c, out=None):
    return vanilla(shape, A, b, c, out=out, stable=True)


def vanilla_batch(shape, A, b, c, stable=False, out=None, nthreads=None):
    """vanilla_batch_numba (vanilla/batch.py:27-61)."""
    shape = tuple(int(s) for s in shape)
    D = len(shape)
    B = int(np.asarray(b).shape[0])
    A = _c128(A, (B, D, D)); b = _c128(b, (B, D)); c = _c128(c, (B,))
    sh = _shape_arr(shape)
    G = out if out is not None else np.empty((B, *shape), dtype=np.complex128)
    nthreads = nthreads or os.cpu_count() or 1
    lib().mmo_vanilla_batch(ctypes.c_int64(B), ctypes.c_int(D), _p(sh), _p(A), _p(b), _p(c),
                            ctypes.c_int(int(bool(stable))), _p(G), ctypes.c_int(nthreads))
    return G


def vanilla_vjp(G, c, dLdG):
    """vanilla_vjp_numba (vanilla/gradients.py:25-82)."""
    G = _c128(G); dLdG = _c128(dLdG, G.shape); c = _c128(c, (1,))
    D = G.ndim
    sh = _shape_arr(G.shape)
    dA = np.empty((D, D), np.complex128); db = np.empty((D,), np.complex128); dc = np.empty((1,), np.complex128)
    lib().mmo_vanilla_vjp(ctypes.c_int(D), _p(sh), _p(G), _p(c), _p(dLdG), _p(dA), _p(db), _p(dc))
    return dA, db, complex(dc[0])


def vanilla_batch_vjp(G, c, dLdG, nthreads=None):
    """vanilla_batch_vjp_numba (vanilla/gradients.py:85-116)."""
    G = _c128(G); dLdG = _c128(dLdG, G.shape)
    B = G.shape[0]; D = G.ndim - 1
    c = _c128(c, (B,))
    sh = _shape_arr(G.shape[1:])
    dA = np.empty((B, D, D), np.complex128); db = np.empty((B, D), np.complex128); dc = np.empty((B,), np.complex128)
    nthreads = nthreads or os.cpu_count() or 1
    lib().mmo_vanilla_batch_vjp(ctypes.c_int64(B), ctypes.c_int(D), _p(sh), _p(G), _p(c), _p(dLdG),
                                _p(dA), _p(db), _p(dc), ctypes.c_int(nthreads))
    return dA, db, dc


def binomial(local_cutoffs, A, b, c, max_l2, global_cutoff):
    """strategies.binomial (binomial.py:30-72) -> (G, norm)."""
    shape = tuple(int(s) for s in local_cutoffs)
    D = len(shape)
    A = _c128(A, (D, D)); b = _c128(b, (D,)); c = _c128(c, (1,))
    sh = _shape_arr(shape)
    G = np.empty(shape, np.complex128)
    norm = ctypes.c_double(0.0)
    lib().mmo_binomial(ctypes.c_int(D), _p(sh), _p(A), _p(b), _p(c), ctypes.c_double(float(max_l2)),
                       ctypes.c_int64(int(global_cutoff)), _p(G), ctypes.byref(norm))
    return G, norm.value


def vanilla_contract(shape, shape_derived, A, b, c_poly, stable=False):
    """CircuitComponent.fock_array's derived-variable branch (lab/circuit_components.py:516-530), restated: the lattice over
    shape + shape_derived with vacuum amplitude 1 (C oracle), then the reference's einsum over the flattened derived axes."""
    shape = tuple(int(s) for s in shape); shape_derived = tuple(int(s) for s in shape_derived)
    G = vanilla(shape + shape_derived, A, b, 1.0 + 0.0j, stable=stable)
    G = G.reshape(shape + (-1,))
    cs = np.asarray(c_poly, dtype=np.complex128).reshape(-1)
    core = "".join(chr(97 + i) for i in range(G.ndim))
    return np.einsum(f"{core},{core[-1]}->{core[:-1]}", G, cs)
