#!/usr/bin/env python
"""A few cfg3 forward + VJP calls for ncu (launch list / full capture of the batched lane kernels)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from conftest import random_triple
from mrmustard_b200 import _lib
lib, check = _lib.lib, _lib.check
dev = torch.device("cuda:0")
shape, B = (40, 40), 65536
sh = _lib.shape_array(shape)
A, b, c = random_triple(2, (B,), seed=673)
dA, db, dc = (torch.from_numpy(np.ascontiguousarray(x)).to(dev) for x in (A, b, c))
G = torch.empty((B, 1600), dtype=torch.complex128, device=dev)
g = torch.randn((B, 1600), dtype=torch.float64, device=dev).to(torch.complex128)
oA = torch.empty((B, 2, 2), dtype=torch.complex128, device=dev); ob = torch.empty((B, 2), dtype=torch.complex128, device=dev); oc = torch.empty(B, dtype=torch.complex128, device=dev)
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 3):
    check(lib.mmh_forward_batched(B, 2, sh, dA.data_ptr(), db.data_ptr(), dc.data_ptr(), G.data_ptr(), 0, None))
    check(lib.mmh_vjp_batched(B, 2, sh, G.data_ptr(), dc.data_ptr(), g.data_ptr(), oA.data_ptr(), ob.data_ptr(), oc.data_ptr(), None))
torch.cuda.synchronize()
