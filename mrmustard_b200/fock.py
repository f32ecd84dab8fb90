"""Fock-space consumers of the lattice on the GPU: label contraction and reduce.

`contract` mirrors ArrayAnsatz.contract (mrmustard/physics/ansatz/array_ansatz.py:159-225): two Fock arrays, one label per axis
(str = batch label, int = core index), einsum over the labels with every SHARED label truncated to the common minimum of its two
dims; `reduce` mirrors ArrayAnsatz.reduce (:227-267): slice or zero-pad the trailing core dims.  numpy arrays go through the
host-pointer entry points; CUDA tensors stay on the device (the lattice produced by mrmustard_b200.device never crosses PCIe
before it is consumed).  There is no CPU fallback.
"""
from __future__ import annotations

import ctypes

import numpy as np

from . import _lib
from ._lib import check, lib, shape_array

__all__ = ["contract", "reduce"]


def _labels(idx1, idx2, idx_out):
    all_in = set(idx1) | set(idx2)
    if not set(idx_out).issubset(all_in):
        raise ValueError("Output labels must be present in input labels.")
    order = sorted(all_in, key=lambda x: (isinstance(x, int), x))        # array_ansatz.py:197
    code = {lab: i for i, lab in enumerate(order)}
    if len(code) > 127:
        raise ValueError("too many distinct labels")
    to = lambda idx: (ctypes.c_int * len(idx))(*[code[i] for i in idx])
    return to(idx1), to(idx2), to(idx_out)


def _out_shape(shape1, idx1, shape2, idx2, idx_out):
    dims = {}
    for lab, d in zip(idx1, shape1):
        dims[lab] = int(d)
    for lab, d in zip(idx2, shape2):
        dims[lab] = min(dims[lab], int(d)) if lab in dims else int(d)
    return tuple(dims[lab] for lab in idx_out)


def _is_cuda_tensor(x) -> bool:
    return type(x).__module__.startswith("torch") and getattr(x, "is_cuda", False)


def contract(array1, idx1, array2, idx2, idx_out):
    """einsum of two Fock arrays by labels (ArrayAnsatz.contract).  Returns the contracted array (numpy in -> numpy out, CUDA
    tensors in -> CUDA tensor out) with axes in the order of idx_out."""
    idx1, idx2, idx_out = list(idx1), list(idx2), list(idx_out)
    if len(idx1) != array1.ndim:
        raise ValueError(f"expected len(idx1)={array1.ndim} got {len(idx1)}")
    if len(idx2) != array2.ndim:
        raise ValueError(f"expected len(idx2)={array2.ndim} got {len(idx2)}")
    l1, l2, lo = _labels(idx1, idx2, idx_out)
    if len(set(idx1)) != len(idx1) or len(set(idx2)) != len(idx2) or len(set(idx_out)) != len(idx_out):
        raise NotImplementedError("a label repeated inside one operand (a trace) is not supported by the CUDA contraction")
    oshape = _out_shape(array1.shape, idx1, array2.shape, idx2, idx_out)
    if _is_cuda_tensor(array1) or _is_cuda_tensor(array2):
        import torch
        a = array1.to(torch.complex128).contiguous()
        b = array2.to(device=a.device, dtype=torch.complex128).contiguous()
        with torch.cuda.device(a.device):
            out = torch.empty(oshape, dtype=torch.complex128, device=a.device)
            if out.numel():
                check(lib.mmh_fock_contract(a.ndim, shape_array(a.shape), l1, b.ndim, shape_array(b.shape), l2, len(idx_out), lo,
                                            a.data_ptr(), b.data_ptr(), out.data_ptr(), None,
                                            ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))
        return out
    a = np.ascontiguousarray(np.asarray(array1, dtype=np.complex128))
    b = np.ascontiguousarray(np.asarray(array2, dtype=np.complex128))
    out = _lib.pinned_empty(oshape)
    if out.size:
        check(lib.mmh_fock_contract_host(a.ndim, shape_array(a.shape), l1, b.ndim, shape_array(b.shape), l2, len(idx_out), lo,
                                         ctypes.c_void_p(a.ctypes.data), ctypes.c_void_p(b.ctypes.data), ctypes.c_void_p(out.ctypes.data)))
    return out


def reduce(array, shape, batch_dims: int = 0):
    """Slice / zero-pad the core dims of a Fock array to `shape` (ArrayAnsatz.reduce).  CUDA tensors only stay on the device;
    numpy arrays are views / np.pad on the host exactly as in the reference (no arithmetic is involved)."""
    shape = tuple(int(s) for s in shape)
    core = tuple(array.shape[batch_dims:])
    if len(shape) != len(core):
        raise ValueError(f"Expected shape of length {len(core)}, got {len(shape)}.")
    if shape == core:
        return array
    if not _is_cuda_tensor(array):
        a = np.asarray(array)
        if any(s > t for s, t in zip(shape, core)):
            return np.pad(a, [(0, 0)] * batch_dims + [(0, max(s - t, 0)) for s, t in zip(shape, core)])[
                (..., *tuple(slice(0, s) for s in shape))]
        return a[(..., *tuple(slice(0, s) for s in shape))]
    import torch
    a = array.to(torch.complex128).contiguous()
    full_out = tuple(a.shape[:batch_dims]) + shape
    with torch.cuda.device(a.device):
        out = torch.empty(full_out, dtype=torch.complex128, device=a.device)
        check(lib.mmh_fock_reduce(a.ndim, shape_array(a.shape), shape_array(full_out), a.data_ptr(), out.data_ptr(),
                                  ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))
    return out
