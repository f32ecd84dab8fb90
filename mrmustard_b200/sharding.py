"""Multi-GPU partitioning of the batched path: one process per GPU, triples sharded contiguously, no collective on
the data path (the lattices of different triples are independent); an optional final gather.

Reference semantics: vanilla_batch_numba / vanilla_batch_vjp_numba are `prange` loops over independent triples
(vanilla/batch.py:56-61, vanilla/gradients.py:110-116) — there is nothing to reduce across the batch.
`torch.distributed` is plumbing only: NCCL for device tensors on the GPU box, gloo for the CPU tests of this logic.
"""
from __future__ import annotations

import numpy as np


def shard_range(batch: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous, balanced partition of range(batch): the first batch % world ranks get one extra triple."""
    base, extra = divmod(int(batch), int(world))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def _dist():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        return dist, dist.get_rank(), dist.get_world_size()
    return None, 0, 1


def _all_gather_rows(local: np.ndarray, batch: int, dist, world: int) -> np.ndarray:
    """Gather variable-length shards (first axis) from every rank into the full array, on every rank."""
    import torch
    row_shape = local.shape[1:]
    maxrows = (batch + world - 1) // world
    use_cuda = dist.get_backend() == "nccl"
    dev = torch.device("cuda", torch.cuda.current_device()) if use_cuda else torch.device("cpu")
    pad = torch.zeros((maxrows, *row_shape), dtype=torch.complex128, device=dev)
    if local.shape[0]:
        pad[: local.shape[0]] = torch.from_numpy(np.ascontiguousarray(local)).to(dev)
    # complex collectives are not supported by every backend: ship the (re, im) view
    send = torch.view_as_real(pad).contiguous()
    recv = [torch.empty_like(send) for _ in range(world)]
    dist.all_gather(recv, send)
    parts = []
    for r in range(world):
        lo, hi = shard_range(batch, r, world)
        parts.append(torch.view_as_complex(recv[r])[: hi - lo].cpu().numpy())
    return np.concatenate(parts, axis=0) if parts else local


def forward_batched_sharded(shape, A, b, c, stable=False, gather=True, compute=None):
    """hermite_renormalized_batched over the ranks of the default process group.

    Every rank passes the FULL (A[B,D,D], b[B,D], c[B]); rank r computes triples shard_range(B, r, world) on its own
    GPU.  gather=False returns (local_shard, (lo, hi)) — the throughput configuration, results stay sharded;
    gather=True returns the full (B, *shape) array on every rank (one all_gather, reported separately in bench.py).
    `compute` defaults to the CUDA path (strategies.vanilla_batch_numba); the CPU tests inject the oracle."""
    if compute is None:
        from . import strategies
        compute = strategies.vanilla_batch_numba
    A = np.asarray(A); b = np.asarray(b); c = np.asarray(c)
    B = b.shape[0]
    dist, rank, world = _dist()
    lo, hi = shard_range(B, rank, world)
    if hi > lo:
        local = compute(tuple(shape), A[lo:hi], b[lo:hi], c[lo:hi], stable)
    else:
        local = np.empty((0, *tuple(shape)), np.complex128)
    if not gather:
        return local, (lo, hi)
    if world == 1:
        return local
    return _all_gather_rows(local, B, dist, world)


def vjp_batched_sharded(G_local, c, dLdG_local, rows, gather=True, compute=None):
    """vanilla_batch_vjp_numba on the local shard (rows = (lo, hi) of the global batch); per-triple gradients, so the
    only collective is the optional gather of dLdA[B,D,D], dLdb[B,D], dLdc[B]."""
    if compute is None:
        from . import strategies
        compute = strategies.vanilla_batch_vjp_numba
    lo, hi = rows
    c = np.asarray(c)
    dA, db, dc = compute(G_local, c[lo:hi], dLdG_local)
    dist, rank, world = _dist()
    if not gather or world == 1:
        return dA, db, dc
    B = c.shape[0]
    return (_all_gather_rows(dA, B, dist, world), _all_gather_rows(db, B, dist, world),
            _all_gather_rows(dc.reshape(-1, 1), B, dist, world).reshape(-1))
