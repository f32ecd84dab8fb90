#!/usr/bin/env python
"""A batch of (20,)^4 lattices through the box march, for ncu."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from conftest import random_triple
from mrmustard_b200 import _lib
dev = torch.device("cuda:0")
shape = (20,) * 4; B = int(sys.argv[1]) if len(sys.argv) > 1 else 296
D = 4; n = int(np.prod(shape)); sh = _lib.shape_array(shape)
A, b, c = random_triple(D, (B,), seed=3)
dA, db, dc = (torch.from_numpy(np.ascontiguousarray(x)).to(dev) for x in (A, b, c))
dG = torch.empty((B, n), dtype=torch.complex128, device=dev)
for _ in range(2):
    _lib.check(_lib.lib.mmh_forward_batched(B, D, sh, dA.data_ptr(), db.data_ptr(), dc.data_ptr(), dG.data_ptr(), 0, None))
torch.cuda.synchronize()
