"""Multi-GPU (NCCL) parity tests; skipped unless at least two CUDA devices are visible (run with gpurun --gpus 2)."""
import os
import socket

import numpy as np
import pytest
import torch

from conftest import random_triple

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    import oracle
    from mrmustard_b200 import sharding
    ok = True
    # batched path, sharded triples
    A, b, c = random_triple(2, (101,), seed=5)
    full = sharding.forward_batched_sharded((12, 11), A, b, c, gather=True)
    ok &= bool(np.array_equal(full, oracle.vanilla_batch((12, 11), A, b, c)))
    # one lattice, stage 0 sharded by panel ranges with NCCL send/recv halos (bit-exact: same per-point arithmetic)
    for shape, seed in [((9, 8, 7, 6), 3), ((6, 5, 4, 3, 4, 5, 3), 4), ((30, 300), 6)]:
        A, b, c = random_triple(len(shape), (), seed=seed)
        G = sharding.forward_single_sharded(shape, A, b, complex(c), gather=True)
        ok &= bool(np.array_equal(G.cpu().numpy().reshape(shape), oracle.vanilla(shape, A, b, complex(c))))
    # the same march replayed from a CUDA graph (NCCL send/recv + the library's launches captured once), twice
    shape, seed = (9, 8, 7, 6), 3
    A, b, c = random_triple(len(shape), (), seed=seed)
    plan = sharding.SingleLatticePlan(shape, A, b, complex(c))
    want = oracle.vanilla(shape, A, b, complex(c))
    for _ in range(2):
        plan.G.zero_()
        plan.run_graphed()
        ok &= bool(np.array_equal(plan.gather().cpu().numpy().reshape(shape), want))
    graphed = getattr(plan, "_graph", None) not in (None, False)
    q.put((rank, ok, graphed, getattr(plan, "_graph_error", None)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_two_gpu_sharding():
    import torch.multiprocessing as mp
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs: p.start()
    res = [q.get(timeout=300) for _ in range(world)]
    for p in procs: p.join(timeout=60)
    assert all(r[1] for r in res), res
    print("graph replay used:", [r[2] for r in res], [r[3] for r in res])
