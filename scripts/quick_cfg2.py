#!/usr/bin/env python
"""Quick device-resident timing of one cfg2 lattice (debug aid): prints ms per lattice and a parity flag."""
import ctypes, hashlib, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mrmustard_b200 import _lib
gold = np.load("tests/golden/vanilla_golden.npz")
dev = torch.device("cuda:0")
dA, db, dc = (torch.from_numpy(np.ascontiguousarray(gold[k])).to(dev) for k in ("cfg2_A", "cfg2_b", "cfg2_c"))
dc = dc.reshape(1)
shape = (50,) * 4
sh = _lib.shape_array(shape)
dG = torch.empty(shape, dtype=torch.complex128, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def run(): _lib.check(_lib.lib.mmh_forward(4, sh, dA.data_ptr(), db.data_ptr(), dc.data_ptr(), dG.data_ptr(), 0, None))
for _ in range(3): run()
torch.cuda.synchronize()
ms = []
for _ in range(20):
    flush.fill_(1)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); run(); b.record(); torch.cuda.synchronize(); ms.append(a.elapsed_time(b))
ok = hashlib.sha256((dG.cpu().numpy() + 0.0).tobytes()).hexdigest() == str(gold["cfg2_G50_sha"])
print(f"cfg2 lattice: median {np.median(ms)*1e3:.1f} us, min {min(ms)*1e3:.1f} us, parity {'OK' if ok else 'MISMATCH'}")
