"""Device-resident entry points with autograd: CUDA tensors in, CUDA tensors out, nothing crosses PCIe.

This is the stand-in for the reference's jax boundary: there the lattice is a
`jax.custom_vjp` whose forward calls `strategies.vanilla_numba` through `pure_callback` and whose backward calls
`strategies.vanilla_vjp_numba` on the residuals (G, c) (mrmustard/math/jax_vjps/hermite.py:47-102 unbatched,
:107-175 batched).  jax is not installable in this image, so the same two-function contract is expressed as a
`torch.autograd.Function` over the device-pointer C ABI (`mmh_forward[_batched]`, `mmh_vjp[_batched]`): forward saves
(G, c), backward runs the reverse recurrence on the device.  torch is plumbing (device memory, streams, the autograd
tape); all arithmetic is in libmmhermite.so.

Cotangent convention.  The reference's bwd returns the plain, un-conjugated sum  dL/dtheta = sum_k g_k dG_k/dtheta
and its optimizer conjugates afterwards (jax_vjps/hermite.py:86-102, training/optimizer.py:104).  torch hands a
backward the conjugate Wirtinger cotangent  gbar_k = dL/d(conj G_k)  and expects  dL/d(conj theta)  back; for a
holomorphic G(theta) that is  conj( sum_k conj(gbar_k) dG_k/dtheta ), i.e. the reference VJP applied to conj(gbar),
conjugated.  `vanilla_vjp(G, c, dLdG)` below exposes the reference's own convention on device tensors.

As in the reference, dL/dA is returned symmetrised ((U + U^T) / 2, gradients.py:79) -- the gradient with respect to a
symmetric A whose two triangles move together.
"""
from __future__ import annotations

import ctypes
import math

import torch

from . import _lib
from ._lib import check, lib, shape_array

__all__ = ["hermite_renormalized", "hermite_renormalized_batched", "vanilla_vjp", "vanilla_batch_vjp",
           "HermiteRenormalized", "HermiteRenormalizedBatched", "FidelityStep"]

_C128 = torch.complex128


def _stream() -> ctypes.c_void_p:
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _dev(x, what: str) -> torch.Tensor:
    if not isinstance(x, torch.Tensor) or not x.is_cuda:
        raise TypeError(f"{what} must be a CUDA tensor (this is the device-resident API; numpy callers use mrmustard_b200.backend)")
    if x.dtype != _C128:
        x = x.to(_C128)
    return x.contiguous()


def _shape(shape) -> tuple[int, ...]:
    shape = tuple(int(s) for s in shape)
    if any(s < 1 for s in shape):
        raise ValueError(f"shape {shape} must have all entries >= 1")
    return shape


def _forward_raw(A, b, c, shape, stable, out=None):
    """A[D,D], b[D], c[] (or [1]) -> G[*shape] on the current stream of A's device."""
    D = len(shape)
    if b.shape[-1] != D:
        raise ValueError(f"len(shape)={D} must equal b.shape[-1]={b.shape[-1]}")
    with torch.cuda.device(A.device):
        G = out if out is not None else torch.empty(shape, dtype=_C128, device=A.device)
        check(lib.mmh_forward(D, shape_array(shape), A.data_ptr(), b.data_ptr(), c.data_ptr(), G.data_ptr(),
                              int(bool(stable)), _stream()))
    return G


def _forward_batched_raw(A, b, c, shape, stable, out=None):
    D = len(shape)
    B = b.shape[0]
    if b.shape[-1] != D:
        raise ValueError(f"len(shape)={D} must equal b.shape[-1]={b.shape[-1]}")
    with torch.cuda.device(A.device):
        G = out if out is not None else torch.empty((B, *shape), dtype=_C128, device=A.device)
        if B:
            check(lib.mmh_forward_batched(B, D, shape_array(shape), A.data_ptr(), b.data_ptr(), c.data_ptr(), G.data_ptr(),
                                          int(bool(stable)), _stream()))
    return G


def vanilla_vjp(G, c, dLdG):
    """strategies.vanilla_vjp_numba on device tensors (vanilla/gradients.py:25-82): reference convention, no conjugation.
    Returns (dLdA[D,D] symmetrised, dLdb[D], dLdc[])."""
    G, dLdG = _dev(G, "G"), _dev(dLdG, "dLdG")
    c = _dev(c, "c").reshape(1)
    if dLdG.shape != G.shape:
        raise ValueError(f"dLdG.shape={tuple(dLdG.shape)} must equal G.shape={tuple(G.shape)}")
    D = G.ndim
    with torch.cuda.device(G.device):
        out = torch.empty(D * D + D + 1, dtype=_C128, device=G.device)
        dA, db, dc = out[: D * D], out[D * D: D * D + D], out[D * D + D:]
        check(lib.mmh_vjp(D, shape_array(G.shape), G.data_ptr(), c.data_ptr(), dLdG.data_ptr(), dA.data_ptr(), db.data_ptr(),
                          dc.data_ptr(), _stream()))
    return dA.view(D, D), db, dc.view(())


def vanilla_batch_vjp(G, c, dLdG):
    """strategies.vanilla_batch_vjp_numba on device tensors (vanilla/gradients.py:85-116): per-triple gradients."""
    G, dLdG = _dev(G, "G"), _dev(dLdG, "dLdG")
    if dLdG.shape != G.shape:
        raise ValueError(f"dLdG.shape={tuple(dLdG.shape)} must equal G.shape={tuple(G.shape)}")
    B, D = G.shape[0], G.ndim - 1
    c = _dev(c, "c").reshape(B)
    with torch.cuda.device(G.device):
        dA = torch.empty((B, D, D), dtype=_C128, device=G.device)
        db = torch.empty((B, D), dtype=_C128, device=G.device)
        dc = torch.empty((B,), dtype=_C128, device=G.device)
        if B:
            check(lib.mmh_vjp_batched(B, D, shape_array(G.shape[1:]), G.data_ptr(), c.data_ptr(), dLdG.data_ptr(), dA.data_ptr(),
                                      db.data_ptr(), dc.data_ptr(), _stream()))
    return dA, db, dc


class HermiteRenormalized(torch.autograd.Function):
    """hermite_renormalized_unbatched_jax with its custom_vjp (jax_vjps/hermite.py:47-102) as a torch autograd node."""

    @staticmethod
    def forward(ctx, A, b, c, shape, stable):
        G = _forward_raw(A, b, c.reshape(1), shape, stable)
        ctx.save_for_backward(G, c)
        return G

    @staticmethod
    def backward(ctx, gbar):
        G, c = ctx.saved_tensors
        dA, db, dc = vanilla_vjp(G, c, gbar.conj().resolve_conj())
        return dA.conj().resolve_conj(), db.conj().resolve_conj(), dc.conj().resolve_conj().reshape(c.shape), None, None


class HermiteRenormalizedBatched(torch.autograd.Function):
    """hermite_renormalized_batched_jax with its custom_vjp (jax_vjps/hermite.py:107-175)."""

    @staticmethod
    def forward(ctx, A, b, c, shape, stable):
        G = _forward_batched_raw(A, b, c, shape, stable)
        ctx.save_for_backward(G, c)
        return G

    @staticmethod
    def backward(ctx, gbar):
        G, c = ctx.saved_tensors
        dA, db, dc = vanilla_batch_vjp(G, c, gbar.conj().resolve_conj())
        return dA.conj().resolve_conj(), db.conj().resolve_conj(), dc.conj().resolve_conj(), None, None


def hermite_renormalized_batched(A, b, c, shape, stable=False, out=None):
    """BackendNumpy.hermite_renormalized_batched (backend_numpy.py:394-403) on CUDA tensors: A[B,D,D], b[B,D], c[B] ->
    G[B,*shape] on the same device, differentiable with respect to A, b, c (vanilla rule; as in the reference the VJP is
    that of the vanilla recurrence).  `out` (a C-contiguous complex128 CUDA tensor of B*prod(shape) entries) is written in
    place and disables autograd for the call, like the jax backend refuses `out` (backend_jax.py:485-486)."""
    shape = _shape(shape)
    A, b, c = _dev(A, "A"), _dev(b, "b"), _dev(c, "c")
    B, D = b.shape
    if A.shape != (B, D, D):
        A = A.expand(B, D, D).contiguous()
    if c.shape != (B,):
        c = c.expand(B).contiguous()
    if out is not None:
        if any(t.requires_grad for t in (A, b, c)):
            raise ValueError("'out' keyword is not supported together with autograd")
        if out.dtype != _C128 or not out.is_contiguous() or out.numel() != B * math.prod(shape):
            raise ValueError("out must be a contiguous complex128 CUDA tensor of batch x prod(shape) entries")
        _forward_batched_raw(A, b, c, shape, stable, out)
        return out.view(B, *shape)
    return HermiteRenormalizedBatched.apply(A, b, c, shape, bool(stable))


def hermite_renormalized(A, b, c, shape, stable=False, out=None):
    """BackendManager.hermite_renormalized (backend_manager.py:643-727) on CUDA tensors, same three batching branches:
    fully batched (A[...,D,D], b[...,D], c[...]), b-batched (A[D,D], b[...,D], c scalar) and unbatched."""
    shape = _shape(shape)
    A, b = _dev(A, "A"), _dev(b, "b")
    if not isinstance(c, torch.Tensor):
        c = torch.tensor(complex(c), dtype=_C128, device=A.device)
    c = _dev(c, "c")
    if A.ndim > 2 and b.ndim > 1 and c.ndim > 0:
        batch_shape = tuple(A.shape[:-2])
        if tuple(b.shape[:-1]) != batch_shape:
            raise ValueError(f"b.shape={tuple(b.shape)} must match batch_shape={batch_shape}")
        if tuple(c.shape[: len(batch_shape)]) != batch_shape:
            raise ValueError(f"c.shape={tuple(c.shape)} must match batch_shape={batch_shape}")
        D = b.shape[-1]
        G = hermite_renormalized_batched(A.reshape(-1, D, D), b.reshape(-1, D), c.reshape(-1), shape, stable,
                                         out.reshape(-1, *shape) if out is not None else None)
        return G.reshape(batch_shape + shape)
    if A.ndim == 2 and b.ndim > 1:
        batch_shape = tuple(b.shape[:-1])
        D = b.shape[-1]
        Bn = 1
        for s in batch_shape:
            Bn *= s
        G = hermite_renormalized_batched(A.expand(Bn, D, D), b.reshape(Bn, D), c.reshape(()).expand(Bn), shape, stable,
                                         out.reshape(Bn, *shape) if out is not None else None)
        return G.reshape(batch_shape + shape)
    if out is not None:
        if any(t.requires_grad for t in (A, b, c)):
            raise ValueError("'out' keyword is not supported together with autograd")
        _forward_raw(A, b, c.reshape(1), shape, stable, out)
        return out.view(shape)
    return HermiteRenormalized.apply(A, b, c, shape, bool(stable))


class FidelityStep:
    """One optimisation step's worth of device work for a fidelity cost, BASELINE config 5: L(A, b, c) = 1 - |<target|G(A, b, c)>|^2
    and its gradient, the quantity `Optimizer.minimize` needs per iteration (mrmustard/training/optimizer.py:82-105 with the
    custom_vjp of math/jax_vjps/hermite.py:47-102).

        step = FidelityStep(shape, target)            # target: lattice-shaped array (numpy or CUDA tensor), kept on the device
        loss, dLdA, dLdb, dLdc = step(A, b, c)        # host (numpy) or device triples

    Per call: ONE 21-number H2D copy of the packed triple (host inputs), the forward lattice (mmh_forward), the overlap
    s = <target|G> (mmh_overlap), the reverse recurrence with the constant cotangent conj(target) (mmh_vjp; the VJP is linear in
    its cotangent and the fidelity's cotangent is the rank-one -conj(s) conj(target), so the scalar factor is applied to the 21
    results instead of to the lattice), and ONE D2H copy of 22 numbers.  The lattice never leaves the device and no lattice-sized
    temporary is formed.  Gradients are returned in the reference's convention (the plain sums dL/dtheta = sum_k (dL/dG_k)
    dG_k/dtheta that the jax bwd returns and the optimizer then conjugates, optimizer.py:104); dLdA is symmetrised like
    vanilla_vjp_numba's (gradients.py:79)."""

    def __init__(self, shape, target, device_index=None, use_graph=True):
        self.use_graph = bool(use_graph)
        self.shape = _shape(shape)
        self.D = len(self.shape)
        dev = torch.device("cuda", torch.cuda.current_device() if device_index is None else device_index)
        self.dev = dev
        t = target if isinstance(target, torch.Tensor) else torch.from_numpy(__import__("numpy").ascontiguousarray(target))
        t = t.to(device=dev, dtype=_C128).reshape(self.shape)
        self.tconj = t.conj().resolve_conj().contiguous()
        D = self.D
        self.n_in = D * D + D + 1
        self.h_in = torch.empty(self.n_in, dtype=_C128).pin_memory()
        self.d_in = torch.empty(self.n_in, dtype=_C128, device=dev)
        self.d_out = torch.empty(self.n_in + 1, dtype=_C128, device=dev)          # [s | dA (D*D) | db (D) | dc]
        self.h_out = torch.empty(self.n_in + 1, dtype=_C128).pin_memory()
        self.G = torch.empty(self.shape, dtype=_C128, device=dev)
        self._sh = shape_array(self.shape)
        self._n = int(self.G.numel())

    def _enqueue(self, host_inputs: bool):
        """H2D of the packed triple (host inputs), forward, overlap, VJP, D2H of the 22 results -- all on the current stream."""
        D = self.D
        st = _stream()
        if host_inputs:
            self.d_in.copy_(self.h_in, non_blocking=True)
        pA = self.d_in.data_ptr()
        pb, pc = pA + 16 * D * D, pA + 16 * (D * D + D)
        po = self.d_out.data_ptr()
        check(lib.mmh_forward(D, self._sh, pA, pb, pc, self.G.data_ptr(), 0, st))
        check(lib.mmh_overlap(self._n, self.tconj.data_ptr(), self.G.data_ptr(), po, st))
        check(lib.mmh_vjp(D, self._sh, self.G.data_ptr(), pc, self.tconj.data_ptr(), po + 16, po + 16 * (1 + D * D),
                          po + 16 * (1 + D * D + D), st))
        self.h_out.copy_(self.d_out, non_blocking=True)

    def _capture(self):
        """The step as a CUDA graph (two copies + seven kernels with fixed addresses): replaying it removes the per-launch host
        work of an optimisation loop.  Falls back to eager launches if the capture is refused."""
        side = torch.cuda.Stream(device=self.dev)
        side.wait_stream(torch.cuda.current_stream(self.dev))
        try:
            with torch.cuda.stream(side):
                self._enqueue(True)          # warm-up on the capture stream (scratch, tables, plans)
                self._enqueue(True)
            side.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=side):
                self._enqueue(True)
            self._graph = g
        except Exception as e:               # pragma: no cover
            self._graph = False
            self._graph_error = repr(e)
            torch.cuda.synchronize(self.dev)

    def __call__(self, A, b, c):
        import numpy as np
        D, dev = self.D, self.dev
        with torch.cuda.device(dev):
            host = not (isinstance(A, torch.Tensor) and A.is_cuda)
            if host:
                h = self.h_in.numpy()
                h[: D * D] = np.asarray(A, dtype=np.complex128).reshape(-1)
                h[D * D: D * D + D] = np.asarray(b, dtype=np.complex128).reshape(-1)
                h[D * D + D] = complex(np.asarray(c).reshape(()))
                if getattr(self, "_graph", None) is None and self.use_graph:
                    self._capture()
                if getattr(self, "_graph", None):
                    self._graph.replay()
                    torch.cuda.synchronize(dev)
                else:
                    self._enqueue(True)
                    torch.cuda.current_stream().synchronize()
            else:
                self.d_in[: D * D] = A.reshape(-1)
                self.d_in[D * D: D * D + D] = b.reshape(-1)
                self.d_in[D * D + D:] = c.reshape(-1)
                self._enqueue(False)
                torch.cuda.current_stream().synchronize()
        o = self.h_out.numpy()
        s = complex(o[0])
        k = -np.conj(s)                                   # dL/dG_k = -conj(s) conj(t_k)
        loss = 1.0 - (s.real * s.real + s.imag * s.imag)
        return loss, (k * o[1: 1 + D * D]).reshape(D, D), k * o[1 + D * D: 1 + D * D + D], complex(k * o[1 + D * D + D])
