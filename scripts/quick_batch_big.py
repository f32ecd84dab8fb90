#!/usr/bin/env python
"""Throughput of a batch of large lattices (pipelined launches): B x (50,)^4 through mmh_forward_batched -- debug aid."""
import hashlib, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mrmustard_b200 import _lib
gold = np.load("tests/golden/vanilla_golden.npz")
dev = torch.device("cuda:0")
shape = (50,) * 4; sh = _lib.shape_array(shape); n = 50 ** 4
for B in (1, 2, 4, 8, 16):
    A = np.repeat(gold["cfg2_A"][None], B, 0); b = np.repeat(gold["cfg2_b"][None], B, 0); c = np.repeat(gold["cfg2_c"].reshape(1), B, 0)
    dA, db, dc = (torch.from_numpy(np.ascontiguousarray(x)).to(dev) for x in (A, b, c))
    dG = torch.empty((B, n), dtype=torch.complex128, device=dev)
    def run(): _lib.check(_lib.lib.mmh_forward_batched(B, 4, sh, dA.data_ptr(), db.data_ptr(), dc.data_ptr(), dG.data_ptr(), 0, None))
    for _ in range(2): run()
    torch.cuda.synchronize()
    ms = []
    for _ in range(8):
        a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); run(); e.record(); torch.cuda.synchronize(); ms.append(a.elapsed_time(e))
    host = dG.cpu().numpy()
    ok = all(hashlib.sha256((host[l].reshape(shape) + 0.0).tobytes()).hexdigest() == str(gold["cfg2_G50_sha"]) for l in range(B))
    t = np.median(ms) * 1e3
    print(f"B={B:2d}: {t:8.1f} us total, {t / B:6.1f} us per lattice, {B * n / t / 1e3:6.1f} G amp/s, parity {'OK' if ok else 'MISMATCH'}")
