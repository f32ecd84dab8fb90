#!/usr/bin/env python
"""(1000,1000) single lattice: stage-0 tile count sweep (debug aid)."""
import os, sys, subprocess
code = r'''
import os, sys
import numpy as np, torch
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
from conftest import random_triple
from mrmustard_b200 import _lib
shape = (1000, 1000)
A, b, c = random_triple(2, (), seed=3); A = A * 0.5
dev = torch.device("cuda:0")
dA, db, dc = (torch.from_numpy(np.ascontiguousarray(x)).to(dev) for x in (A, b, np.array([c])))
sh = _lib.shape_array(shape)
dG = torch.empty(shape, dtype=torch.complex128, device=dev)
def run(): _lib.check(_lib.lib.mmh_forward(2, sh, dA.data_ptr(), db.data_ptr(), dc.data_ptr(), dG.data_ptr(), 0, None))
for _ in range(3): run()
torch.cuda.synchronize()
ms = []
for _ in range(8):
    a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); run(); e.record(); torch.cuda.synchronize(); ms.append(a.elapsed_time(e))
print("%.1f us" % (np.median(ms) * 1e3))
'''
for g in ["2", "4", "8", "12", "16", "24", "32", "48", "64", "100"]:
    env = dict(os.environ, MMH_TILE_G=g, MMH_TILE_STAGE="0")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env)
    print("tiles", g, ":", r.stdout.strip() or r.stderr.strip()[-200:], flush=True)
