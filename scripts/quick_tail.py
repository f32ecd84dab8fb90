#!/usr/bin/env python
"""Device timing of tiny single lattices (the trailing-stage kernels) -- debug aid."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mrmustard_b200 import _lib
rng = np.random.RandomState(0)
dev = torch.device("cuda:0")
for shape in [(2, 50), (50, 50), (450, 50), (50,), (450,), (50, 50, 50)]:
    D = len(shape)
    A = rng.random((D, D)) + 1j * rng.random((D, D)); A = (A + A.T) / 4
    b = rng.random(D) + 1j * rng.random(D); c = np.array([0.5 + 0.1j])
    dA, db, dc = (torch.from_numpy(np.ascontiguousarray(x)).to(dev) for x in (A, b, c))
    sh = _lib.shape_array(shape)
    dG = torch.empty(shape, dtype=torch.complex128, device=dev)
    def run(): _lib.check(_lib.lib.mmh_forward(D, sh, dA.data_ptr(), db.data_ptr(), dc.data_ptr(), dG.data_ptr(), 0, None))
    for _ in range(3): run()
    torch.cuda.synchronize()
    ms = []
    for _ in range(20):
        a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); run(); e.record(); torch.cuda.synchronize(); ms.append(a.elapsed_time(e))
    print(f"{shape}: median {np.median(ms)*1e3:.1f} us  min {min(ms)*1e3:.1f} us")
