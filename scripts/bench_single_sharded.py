#!/usr/bin/env python
"""(12,)*8 single lattice (430 M amplitudes, 6.9 GB): 1 GPU (mmh_forward) vs N GPUs (forward_single_sharded).
Run with torchrun; device time via CUDA events, max over ranks."""
import os, sys, ctypes
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
if world > 1: dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
from mrmustard_b200 import _lib, sharding
gold = np.load("tests/golden/vanilla_golden.npz")
A, b, c = gold["cfg4_A"], gold["cfg4_b"], complex(gold["cfg4_c"])
cut = int(sys.argv[1]) if len(sys.argv) > 1 else 12
shape = (cut,) * 8
n = cut ** 8
def run():
    return sharding.forward_single_sharded(shape, A, b, c, gather=False)
run(); torch.cuda.synchronize()
ms = []
for _ in range(3):
    if world > 1: dist.barrier()
    torch.cuda.synchronize()
    a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); G, rows = run(); e.record(); torch.cuda.synchronize(); ms.append(a.elapsed_time(e))
t = torch.tensor([min(ms)], device="cuda")
if world > 1: dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    print(f"shape {shape}: {world} GPU(s) {float(t):.2f} ms -> {n / float(t) / 1e6:.2f} G amp/s (sharded stage 0, results left sharded)")
    if world == 1:
        dev = torch.device("cuda")
        dA, db, dc = (torch.from_numpy(np.ascontiguousarray(x)).to(dev) for x in (A, b, np.array([c])))
        Gf = torch.empty(n, dtype=torch.complex128, device=dev)
        f = lambda: _lib.check(_lib.lib.mmh_forward(8, _lib.shape_array(shape), dA.data_ptr(), db.data_ptr(), dc.data_ptr(), Gf.data_ptr(), 0, None))
        f(); torch.cuda.synchronize()
        a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); f(); e.record(); torch.cuda.synchronize()
        print(f"   mmh_forward (single call): {a.elapsed_time(e):.2f} ms -> {n / a.elapsed_time(e) / 1e6:.2f} G amp/s; equal to sharded: {bool(torch.equal(Gf, G))}")
if world > 1: dist.destroy_process_group()
