#!/usr/bin/env python
"""A few cfg3 forward calls for ncu (launch list / full capture of the batched march kernels)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from conftest import random_triple
from mrmustard_b200 import _lib
dev = torch.device("cuda:0")
shape, B = (40, 40), 65536
sh = _lib.shape_array(shape)
A, b, c = random_triple(2, (B,), seed=673)
dA, db, dc = (torch.from_numpy(np.ascontiguousarray(x)).to(dev) for x in (A, b, c))
dG = torch.empty((B, 1600), dtype=torch.complex128, device=dev)
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 3):
    _lib.check(_lib.lib.mmh_forward_batched(B, 2, sh, dA.data_ptr(), db.data_ptr(), dc.data_ptr(), dG.data_ptr(), 0, None))
torch.cuda.synchronize()
