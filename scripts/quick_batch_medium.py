#!/usr/bin/env python
"""Batches of medium lattices: one-CTA-per-lattice kernel vs the pipelined single-lattice path (debug aid)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from conftest import random_triple
from mrmustard_b200 import _lib
dev = torch.device("cuda:0")
for shape in [(20,) * 4, (30,) * 4, (64, 64, 64)]:
    D = len(shape); n = int(np.prod(shape)); sh = _lib.shape_array(shape)
    for B in [int(x) for x in os.environ.get('QB', '8,32,148,512').split(',')]:
        A, b, c = random_triple(D, (B,), seed=3)
        dA, db, dc = (torch.from_numpy(np.ascontiguousarray(x)).to(dev) for x in (A, b, c))
        dG = torch.empty((B, n), dtype=torch.complex128, device=dev)
        res = []
        for thr, nobox in (("1000000", None), ("1", "1"), ("1", None)):
            os.environ["MMH_PER_CTA_BATCH"] = thr
            if nobox: os.environ["MMH_NO_BOX"] = nobox
            else: os.environ.pop("MMH_NO_BOX", None)
            def run(): _lib.check(_lib.lib.mmh_forward_batched(B, D, sh, dA.data_ptr(), db.data_ptr(), dc.data_ptr(), dG.data_ptr(), 0, None))
            run(); torch.cuda.synchronize()
            ms = []
            for _ in range(3):
                a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(); run(); e.record(); torch.cuda.synchronize(); ms.append(a.elapsed_time(e))
            res.append(np.median(ms))
        print(f"{shape} B={B:4d}: pipelined {res[0]:8.3f} ms ({B*n/res[0]/1e6:6.1f} G amp/s)   k_fwd_cta {res[1]:8.3f} ms ({B*n/res[1]/1e6:6.1f} G amp/s)   box march {res[2]:8.3f} ms ({B*n/res[2]/1e6:6.1f} G amp/s)", flush=True)
        del dG
