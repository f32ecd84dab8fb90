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


# ---------------------------------------------------------------------------------------------------------------
# ONE large lattice over several GPUs (second bullet)
# ---------------------------------------------------------------------------------------------------------------
# With the first-non-zero pivot every point with k_0 >= 1 reads only panels k_0 - 1 and k_0 - 2 (mmh_forward.cu), and
# inside panel k_0 - 1 only offsets f - strides[j], j >= 1, i.e. at most strides[1] positions back.  So the stage-0
# panels are cut into contiguous ranges of the panel offset f, rank r owns range r of EVERY panel, all ranks compute
# panel s concurrently, and between steps rank r sends the last strides[1] amplitudes of its range to the rank(s)
# above it (NCCL send/recv over NVLink; no reduction, no barrier beyond the p2p dependencies).  The sub-lattice
# k_0 = 0 (1/shape[0] of the work) is computed redundantly by every rank.
class CudaPanelOps:
    """Device-side operations of forward_single_sharded through the C ABI (torch tensors are plumbing only)."""

    def __init__(self):
        import torch
        from . import _lib
        self.torch, self._lib = torch, _lib
        self.device = torch.device("cuda", torch.cuda.current_device())

    def alloc(self, n):
        return self.torch.empty(n, dtype=self.torch.complex128, device=self.device)

    def to_dev(self, x):
        return self.torch.from_numpy(np.ascontiguousarray(np.asarray(x, dtype=np.complex128))).to(self.device)

    def _stream(self):
        import ctypes
        return ctypes.c_void_p(self.torch.cuda.current_stream().cuda_stream)

    def sublattice(self, G, shape, A, b, c):
        L = self._lib
        D = len(shape)
        if getattr(self, "_keep", None) is None:   # device copies of the sub-triple, made once per plan
            self._keep = (self.to_dev(np.asarray(A)[1:, 1:]), self.to_dev(np.asarray(b)[1:]), self.to_dev(np.asarray(c).reshape(1)))
        dA, db, dc = self._keep
        L.check(L.lib.mmh_forward(D - 1, L.shape_array(shape[1:]), dA.data_ptr(), db.data_ptr(), dc.data_ptr(),
                                  G.data_ptr(), 0, self._stream()))

    def prepare(self, shape, A, b):
        self._A, self._b = self.to_dev(A), self.to_dev(b)

    def panel_range(self, G, shape, step, f_lo, f_hi):
        L = self._lib
        L.check(L.lib.mmh_forward_panel_range(len(shape), L.shape_array(shape), self._A.data_ptr(), self._b.data_ptr(),
                                              G.data_ptr(), 0, int(step), int(f_lo), int(f_hi), self._stream()))


class SingleLatticePlan:
    """ONE lattice, stage 0 sharded over the ranks of the default process group by panel ranges; `run()` fills this rank's
    full-size buffer `G` (the complete sub-lattice k_0 = 0 plus this rank's range of every panel and its halos) and can be
    called repeatedly (bench.py --workload cfg4).  Device buffers and the send/recv schedule are built once."""

    def __init__(self, shape, A, b, c, ops=None):
        self.shape = tuple(int(s) for s in shape)
        if len(self.shape) < 2:
            raise ValueError("forward_single_sharded needs at least two modes")
        self.dist, self.rank, self.world = _dist()
        self.ops = ops if ops is not None else CudaPanelOps()
        self.A, self.b, self.c = A, b, c
        self.N = int(np.prod(self.shape, dtype=np.int64))
        self.S0 = self.shape[0]
        self.P = self.N // self.S0
        self.H = int(np.prod(self.shape[2:], dtype=np.int64))    # strides[1]: the deepest look-back inside panel s - 1
        self.G = self.ops.alloc(self.N)
        self.ranges = [shard_range(self.P, r, self.world) for r in range(self.world)]
        self.f_lo, self.f_hi = self.ranges[self.rank]
        # halo schedule of one step: (peer, lo, hi, is_send); rank q needs [f_lo_q - H, f_lo_q) of panel s from the ranks below it
        self.sched = []
        for q in range(self.world):
            need_lo, need_hi = max(self.ranges[q][0] - self.H, 0), self.ranges[q][0]
            for r in range(q):
                lo, hi = max(self.ranges[r][0], need_lo), min(self.ranges[r][1], need_hi)
                if hi <= lo:
                    continue
                if self.rank == r:
                    self.sched.append((q, lo, hi, True))
                elif self.rank == q:
                    self.sched.append((r, lo, hi, False))
        self.ops.prepare(self.shape, A, b)

    def run_graphed(self):
        """run() replayed from a CUDA graph: the per-step host work of the eager loop (two kernel launches, a NCCL group call,
        four event operations -- ~0.4 ms per step against 0.15 ms of device work at 4 GPUs) is paid once at capture.  NCCL
        send / recv and the library's launches are all stream-ordered, so the whole march is capturable.  Falls back to the eager
        loop if capture is not possible (CPU ops, an old NCCL)."""
        import torch
        if not self.G.is_cuda:
            return self.run()
        if getattr(self, "_graph", None) is None:
            side = torch.cuda.Stream(device=self.G.device)
            side.wait_stream(torch.cuda.current_stream(self.G.device))
            with torch.cuda.stream(side):            # warm-up on the capture stream: scratch, tables, NCCL channels
                self.run()
                self.run()
            torch.cuda.current_stream(self.G.device).wait_stream(side)
            torch.cuda.synchronize(self.G.device)
            try:
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, stream=side):
                    self.run()
                self._graph = g
            except Exception as e:   # pragma: no cover - capture support depends on the NCCL / driver combination
                self._graph = False
                self._graph_error = repr(e)
                torch.cuda.synchronize(self.G.device)
        if self._graph is False:
            return self.run()
        self._graph.replay()
        return self.G

    def run(self):
        """Fill this rank's part of the lattice.  Per panel step the rank first computes the LAST H offsets of its range -- the
        amplitudes the ranks above it wait for -- hands them to the communication stream (NCCL send / recv over NVLink) and computes
        the rest of its range while they travel; the next step waits for the received halo only (CUDA events, no host block)."""
        import torch
        dist, ops, G, P = self.dist, self.ops, self.G, self.P
        ops.sublattice(G, self.shape, self.A, self.b, self.c)                 # panel 0, redundantly on every rank
        Gr = torch.view_as_real(G)                                            # NCCL/gloo p2p on the (re, im) view
        cuda = G.is_cuda
        split = self.f_hi - min(self.H, self.f_hi - self.f_lo)                # [split, f_hi) is what the ranks above need
        if cuda and self.sched:
            if not hasattr(self, "_comm"):
                self._comm = torch.cuda.Stream(device=G.device)
            comm, main = self._comm, torch.cuda.current_stream(G.device)
        for s in range(1, self.S0):
            exchange = bool(self.sched) and s < self.S0 - 1
            if not exchange:
                ops.panel_range(G, self.shape, s, self.f_lo, self.f_hi)
                continue
            ops.panel_range(G, self.shape, s, split, self.f_hi)                # boundary first
            reqs = [dist.P2POp(dist.isend if snd else dist.irecv, Gr[s * P + lo: s * P + hi], peer)
                    for peer, lo, hi, snd in self.sched]
            if cuda:
                ready = torch.cuda.Event(); ready.record(main)
                with torch.cuda.stream(comm):
                    comm.wait_event(ready)
                    for w in dist.batch_isend_irecv(reqs):
                        w.wait()
                    done = torch.cuda.Event(); done.record(comm)
                ops.panel_range(G, self.shape, s, self.f_lo, split)            # the rest overlaps the transfer
                main.wait_event(done)
            else:
                ops.panel_range(G, self.shape, s, self.f_lo, split)
                for w in dist.batch_isend_irecv(reqs):
                    w.wait()
        return G

    def gather(self):
        """The full lattice on every rank: rank r owns columns [f_lo_r, f_hi_r) of the (S0 - 1, P) matrix of panels 1..S0-1."""
        import torch
        if self.world == 1:
            return self.G
        dist, G, P, S0 = self.dist, self.G, self.P, self.S0
        maxw = max(hi - lo for lo, hi in self.ranges)
        body = G[P:].view(S0 - 1, P)
        send = torch.zeros((S0 - 1, maxw), dtype=torch.complex128, device=G.device)
        send[:, : self.f_hi - self.f_lo] = body[:, self.f_lo:self.f_hi]
        recv = [torch.empty_like(torch.view_as_real(send)) for _ in range(self.world)]
        dist.all_gather(recv, torch.view_as_real(send).contiguous())
        for r, (lo, hi) in enumerate(self.ranges):
            if r != self.rank:
                body[:, lo:hi] = torch.view_as_complex(recv[r])[:, : hi - lo]
        return G


def forward_single_sharded(shape, A, b, c, gather=True, ops=None):
    """hermite_renormalized (vanilla rule) of ONE lattice, stage 0 sharded over the ranks of the default process group.

    Returns the full lattice on every rank (gather=True) or (G_local, (f_lo, f_hi)) where G_local is this rank's full-size
    buffer holding the complete sub-lattice k_0 = 0 and, for k_0 >= 1, this rank's range of every panel (plus halos)."""
    plan = SingleLatticePlan(shape, A, b, c, ops)
    plan.run()
    if not gather:
        return plan.G, (plan.f_lo, plan.f_hi)
    return plan.gather()
