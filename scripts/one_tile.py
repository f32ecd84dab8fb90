#!/usr/bin/env python
"""One-tile march (no halo): (50,10,10,10) through the tiled kernel with a 1x1x1 grid -- isolates the compute step (debug aid)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["MMH_FORCE_TILED"] = "1"; os.environ["MMH_TILE_G"] = "1,1,1"; os.environ["MMH_TILE_STAGE"] = "0"
from mrmustard_b200 import _lib
gold = np.load("tests/golden/vanilla_golden.npz")
dev = torch.device("cuda:0")
dA, db, dc = (torch.from_numpy(np.ascontiguousarray(gold[k])).to(dev) for k in ("cfg2_A", "cfg2_b", "cfg2_c"))
dc = dc.reshape(1)
shape = (int(sys.argv[1]) if len(sys.argv) > 1 else 50, 10, 10, 10)
sh = _lib.shape_array(shape)
dG = torch.empty(shape, dtype=torch.complex128, device=dev)
def run(): _lib.check(_lib.lib.mmh_forward(4, sh, dA.data_ptr(), db.data_ptr(), dc.data_ptr(), dG.data_ptr(), 0, None))
for _ in range(3): run()
torch.cuda.synchronize()
ms = []
for _ in range(10):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); run(); b.record(); torch.cuda.synchronize(); ms.append(a.elapsed_time(b))
print(f"one tile {shape}: median {np.median(ms)*1e3:.1f} us")
