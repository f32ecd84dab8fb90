#!/usr/bin/env python
"""Device timing + parity of mid-size single lattices whose stage-0 panel fits one CTA (one CTA vs tiles) -- debug aid."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from conftest import random_triple
from mrmustard_b200 import _lib
import oracle
dev = torch.device("cuda:0")
for shape in [(1000, 1000), (300, 300), (100, 500), (2000, 200), (60, 30, 30), (200, 16, 16), (40, 1000)]:
    D = len(shape)
    A, b, c = random_triple(D, (), seed=3)
    A = A * 0.5
    dA, db, dc = (torch.from_numpy(np.ascontiguousarray(x)).to(dev) for x in (A, b, np.array([c])))
    sh = _lib.shape_array(shape)
    dG = torch.empty(shape, dtype=torch.complex128, device=dev)
    def run(): _lib.check(_lib.lib.mmh_forward(D, sh, dA.data_ptr(), db.data_ptr(), dc.data_ptr(), dG.data_ptr(), 0, None))
    for _ in range(3): run()
    torch.cuda.synchronize()
    ms = []
    for _ in range(10):
        a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); run(); e.record(); torch.cuda.synchronize(); ms.append(a.elapsed_time(e))
    ok = np.array_equal(dG.cpu().numpy(), oracle.vanilla(shape, A, b, complex(c)))
    print(f"{shape}: median {np.median(ms)*1e3:.1f} us  parity {'OK' if ok else 'MISMATCH'}")
