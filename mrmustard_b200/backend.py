"""Operator-level mirror of the reference's `math.hermite_renormalized*` entry points.

`hermite_renormalized` is restated from BackendManager.hermite_renormalized's batching semantics
(mrmustard/math/backend_manager.py:643-727: fully batched / b-batched / unbatched, batch flatten and
reshape, `stable or settings.STABLE_FOCK_CONVERSION`, the out-shape check and its ValueError);
`hermite_renormalized_batched` and `hermite_renormalized_binomial` reproduce the BackendNumpy methods
(mrmustard/math/backend_numpy.py:394-421).  Everything computes on the GPU through `strategies`.
"""
from __future__ import annotations

import numpy as np

from . import strategies


class _Settings:
    """The two reference settings that reach this path (mrmustard/utils/settings.py:67-74,101)."""
    STABLE_FOCK_CONVERSION = False
    AUTOSHAPE_PROBABILITY = 0.99999


settings = _Settings()


_installed = False   # set by dropin.install() / uninstall()


def _reference_settings():
    """The live mrmustard settings object while the drop-in is installed into a reference process, else ours."""
    import sys
    mm = sys.modules.get("mrmustard") if _installed else None
    return getattr(mm, "settings", settings) if mm is not None else settings


def hermite_renormalized_unbatched(A, b, c, shape, stable=False, out=None):
    """BackendNumpy.hermite_renormalized (backend_numpy.py:381-392)."""
    if stable:
        return strategies.stable_numba(tuple(shape), A, b, c, out)
    return strategies.vanilla_numba(tuple(shape), A, b, c, out)


def hermite_renormalized_batched(A, b, c, shape, stable=False, out=None):
    """BackendNumpy.hermite_renormalized_batched (backend_numpy.py:394-403)."""
    return strategies.vanilla_batch_numba(tuple(shape), A, b, c, stable, out)


def hermite_renormalized_binomial(A, B, C, shape, max_l2=None, global_cutoff=None):
    """BackendNumpy.hermite_renormalized_binomial (backend_numpy.py:405-421)."""
    shape = tuple(shape)
    s = _reference_settings()
    return strategies.binomial(
        shape, A, B, C,
        max_l2=max_l2 or s.AUTOSHAPE_PROBABILITY,
        global_cutoff=global_cutoff or sum(shape) - len(shape) + 1,
    )[0]


def reorder_AB_bargmann(A, B):
    """BackendNumpy.reorder_AB_bargmann (backend_numpy.py:367-377): [m0..,m0..] -> [m0,m0,m1,m1,..]."""
    A = np.asarray(A)
    B = np.asarray(B)
    ordering = np.arange(2 * A.shape[0] // 2).reshape(2, -1).T.flatten()
    A = np.take(np.take(A, ordering, axis=1), ordering, axis=0)
    B = np.take(B, ordering, axis=0)
    return A, B


def hermite_renormalized_diagonal(A, B, C, cutoffs, reorderedAB=True):
    """BackendNumpy.hermite_renormalized_diagonal (backend_numpy.py:423-432)."""
    A, B = reorder_AB_bargmann(A, B) if reorderedAB else (np.asarray(A), np.asarray(B))
    return strategies.hermite_multidimensional_diagonal(np.ascontiguousarray(A), B, C, cutoffs)


def hermite_renormalized_1leftoverMode(A, b, c, output_cutoff, pnr_cutoffs, stable=False, reorderedAB=True):
    """BackendNumpy.hermite_renormalized_1leftoverMode (backend_numpy.py:434-446): fast_diagonal + transpose to
    (out+1, out+1, *(pnr+1)).  As on the numpy backend the inputs are always taken in bargmann order and permuted
    (fast_diagonal.py:56-58); `stable`/`reorderedAB` are accepted for call compatibility (see SURVEY.md §3.4 for the
    reference's positional-argument slip)."""
    pnr_cutoffs = tuple(pnr_cutoffs)
    return strategies.fast_diagonal(A, b, c, output_cutoff, pnr_cutoffs, stable).transpose(
        (-2, -1, *tuple(range(len(pnr_cutoffs)))))


def hermite_renormalized(A, b, c, shape, stable=False, out=None, device=False):
    """BackendManager.hermite_renormalized, restated from backend_manager.py:643-727 (same three batching branches, same
    out-shape check and error strings).  device=True is this package's extension: A, b, c are CUDA tensors, the result is a CUDA
    tensor on the same device with autograd attached and nothing crosses PCIe (mrmustard_b200.device)."""
    if device:
        from . import device as _device
        return _device.hermite_renormalized(A, b, c, shape, stable or _reference_settings().STABLE_FOCK_CONVERSION, out)
    A = np.asarray(A)
    b = np.asarray(b)
    c = np.asarray(c)
    shape = tuple(shape)

    def check_out_shape(batch_shape):
        if out is not None and any(d_out < d for d_out, d in zip(out.shape, batch_shape + shape)):
            raise ValueError(f"batch+shape {batch_shape + shape} is too large for out.shape={out.shape}")

    stable = stable or _reference_settings().STABLE_FOCK_CONVERSION
    if A.ndim > 2 and b.ndim > 1 and c.ndim > 0:
        batch_shape = A.shape[:-2]
        check_out_shape(batch_shape)
        if b.shape[:-1] != batch_shape:
            raise ValueError(f"b.shape={b.shape} must match batch_shape={batch_shape}")
        if c.shape[: len(batch_shape)] != batch_shape:
            raise ValueError(f"c.shape={c.shape} must match batch_shape={batch_shape}")
        B = int(np.prod(batch_shape))
        result = hermite_renormalized_batched(
            A.reshape(B, *A.shape[-2:]), b.reshape(B, b.shape[-1]), c.reshape(B), shape, stable,
            out.reshape(B, *shape) if out is not None else None)
        return result.reshape(batch_shape + shape)
    if A.ndim == 2 and b.ndim > 1:  # b-batched
        batch_shape = b.shape[:-1]
        check_out_shape(batch_shape)
        B = int(np.prod(batch_shape))
        result = hermite_renormalized_batched(
            np.broadcast_to(A, (B, *A.shape)), b.reshape(B, b.shape[-1]), np.broadcast_to(c, (B,)), shape, stable,
            out.reshape(B, *shape) if out is not None else None)
        return result.reshape(batch_shape + shape)
    check_out_shape(())
    return hermite_renormalized_unbatched(A, b, c, shape, stable, out)


def hermite_renormalized_contracted(A, b, c, shape, stable=False):
    """Fock array of a PolyExpAnsatz with derived variables: `c` carries the polynomial coefficients,
    c.shape = batch_shape + shape_derived_vars, and the result is what CircuitComponent.fock_array computes in its
    `num_derived_vars > 0` branch (lab/circuit_components.py:516-530) -- hermite_renormalized over
    `shape + shape_derived_vars` with unit vacuum amplitude followed by the einsum over the derived axes -- without
    ever materialising the big lattice on the host.  Same batching as hermite_renormalized (A[..., D, D], b[..., D])."""
    A = np.asarray(A)
    b = np.asarray(b)
    c = np.asarray(c)
    shape = tuple(shape)
    stable = stable or _reference_settings().STABLE_FOCK_CONVERSION
    batch_shape = b.shape[:-1]
    D = b.shape[-1]
    n_derived = D - len(shape)
    if n_derived < 0:
        raise ValueError(f"len(shape)={len(shape)} exceeds the number of variables {D}")
    if c.shape[: len(batch_shape)] != batch_shape or c.ndim != len(batch_shape) + n_derived:
        raise ValueError(f"c.shape={c.shape} must be batch_shape={batch_shape} + {n_derived} derived axes")
    shape_derived = c.shape[len(batch_shape):]
    if A.shape[:-2] not in ((), batch_shape):
        raise ValueError(f"A.shape={A.shape} must match batch_shape={batch_shape}")
    if not batch_shape:
        return strategies.vanilla_contract_numba(shape, shape_derived, A, b, c, stable)
    B = int(np.prod(batch_shape))
    Ab = np.broadcast_to(A, (*batch_shape, D, D)).reshape(B, D, D)
    out = strategies.vanilla_contract_numba(shape, shape_derived, Ab, b.reshape(B, D), c.reshape(B, -1), stable)
    return out.reshape(batch_shape + shape)
