"""world_size-2 gloo tests (CPU) of the batch-sharding logic; the per-rank compute is the oracle, injected by the test."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import random_triple


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, B, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle
    from mrmustard_b200 import sharding
    A, b, c = random_triple(2, (B,), seed=5)
    shape = (5, 4)
    fwd = lambda shape, A, b, c, stable: oracle.vanilla_batch(shape, A, b, c, stable, nthreads=1)
    full = sharding.forward_batched_sharded(shape, A, b, c, gather=True, compute=fwd)
    local, rows = sharding.forward_batched_sharded(shape, A, b, c, gather=False, compute=fwd)
    want = oracle.vanilla_batch(shape, A, b, c, nthreads=1)
    ok = np.array_equal(full, want) and np.array_equal(local, want[rows[0]:rows[1]])
    g = np.random.RandomState(3).standard_normal(want.shape) + 0j
    vjp = lambda G, c, g: oracle.vanilla_batch_vjp(G, c, g, nthreads=1)
    dA, db, dc = sharding.vjp_batched_sharded(local, c, g[rows[0]:rows[1]], rows, gather=True, compute=vjp)
    wA, wb, wc = oracle.vanilla_batch_vjp(want, c, g, nthreads=1)
    ok = ok and np.array_equal(dA, wA) and np.array_equal(db, wb) and np.array_equal(dc, wc)
    q.put((rank, bool(ok), rows))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("B", [7, 8, 1])
def test_two_rank_sharding(B):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, B, q)) for r in range(world)]
    for p in procs: p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs: p.join(timeout=60)
    assert all(ok for _, ok, _ in res), res
    rows = sorted(r for _, _, r in res)
    assert rows[0][0] == 0 and rows[-1][1] == B and rows[0][1] == rows[1][0]


def test_shard_range_is_a_partition():
    from mrmustard_b200.sharding import shard_range
    for B in (0, 1, 5, 8, 65536, 65537):
        for world in (1, 2, 3, 4, 8):
            edges = [shard_range(B, r, world) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == B
            assert all(edges[i][1] == edges[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in edges]
            assert max(sizes) - min(sizes) <= 1


# ---- one lattice sharded over two ranks (panel ranges + p2p halo exchange); CPU ops injected by the test ----------
class _CpuPanelOps:
    """numpy restatement of the two device operations, driven by the oracle's arithmetic (test-only)."""

    def alloc(self, n):
        return torch.zeros(n, dtype=torch.complex128)

    def sublattice(self, G, shape, A, b, c):
        import oracle
        sub = oracle.vanilla(shape[1:], np.asarray(A)[1:, 1:], np.asarray(b)[1:], complex(np.asarray(c).reshape(-1)[0]))
        G[: sub.size] = torch.from_numpy(sub.ravel().copy())

    def prepare(self, shape, A, b):
        self.A, self.b = np.asarray(A, complex), np.asarray(b, complex)

    def panel_range(self, G, shape, step, f_lo, f_hi):
        # vanilla update with pivot 0 (core.py:108-122), same operation order as the oracle
        g = G.numpy()
        D = len(shape)
        strides = [int(np.prod(shape[i + 1:])) for i in range(D)]
        P = strides[0]
        for f in range(f_lo, f_hi):
            flat = step * P + f
            pivot = flat - P
            k = np.unravel_index(f, shape[1:]) if D > 1 else ()
            v = self.b[0] * g[pivot]
            if step >= 2:
                v = v + (self.A[0, 0] * np.sqrt(step - 1)) * g[pivot - P]
            for j in range(1, D):
                if k[j - 1] > 0:
                    v = v + (self.A[0, j] * np.sqrt(k[j - 1])) * g[pivot - strides[j]]
            g[flat] = v / np.sqrt(step)


def _single_worker(rank, world, port, shape, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle
    from mrmustard_b200 import sharding
    A, b, c = random_triple(len(shape), (), seed=9)
    G = sharding.forward_single_sharded(shape, A, b, complex(c), gather=True, ops=_CpuPanelOps())
    want = oracle.vanilla(shape, A, b, complex(c))
    q.put((rank, bool(np.allclose(G.numpy().reshape(shape), want, rtol=1e-12, atol=1e-15))))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("shape", [(5, 4, 3), (4, 7), (3, 2, 3, 2)])
def test_single_lattice_two_ranks(shape):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_single_worker, args=(r, world, port, shape, q)) for r in range(world)]
    for p in procs: p.start()
    res = [q.get(timeout=180) for _ in range(world)]
    for p in procs: p.join(timeout=60)
    assert all(ok for _, ok in res), res
