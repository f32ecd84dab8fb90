#!/usr/bin/env python
"""A few forwards of one lattice (default (50,)^4, the cfg2 golden triple) -- the target of ncu captures of the single-lattice kernels."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mrmustard_b200 import _lib
shape = tuple(int(x) for x in sys.argv[1].split(",")) if len(sys.argv) > 1 else (50,) * 4
D = len(shape)
gold = np.load("tests/golden/vanilla_golden.npz")
if shape == (50,) * 4:
    A, b, c = gold["cfg2_A"], gold["cfg2_b"], np.asarray(gold["cfg2_c"]).reshape(1)
else:
    rng = np.random.default_rng(3)
    A = rng.uniform(-1, 1, (D, D)) + 1j * rng.uniform(-1, 1, (D, D)); A = (A + A.T) / 2; A /= np.abs(np.linalg.eigvals(A)).max() * 1.5
    b = rng.uniform(-1, 1, D) + 1j * rng.uniform(-1, 1, D); c = np.array([0.4 + 0.3j])
dev = torch.device("cuda:0")
dA, db, dc = (torch.from_numpy(np.ascontiguousarray(x)).to(dev) for x in (A, b, c))
G = torch.empty(shape, dtype=torch.complex128, device=dev)
sh = _lib.shape_array(shape)
for _ in range(int(os.environ.get("N_RUNS", "4"))):
    _lib.check(_lib.lib.mmh_forward(D, sh, dA.data_ptr(), db.data_ptr(), dc.data_ptr(), G.data_ptr(), 0, None))
torch.cuda.synchronize()
print("ok", complex(G.flatten()[-1]))
