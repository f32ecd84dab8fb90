#!/usr/bin/env python
"""Kernel-level timeline of a batch of 4 cfg2 lattices (pipelined launches) from the device stamps -- debug aid."""
import ctypes, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mrmustard_b200 import _lib
lib = ctypes.CDLL(_lib.SO_PATH)
gold = np.load("tests/golden/vanilla_golden.npz")
dev = torch.device("cuda:0")
B = 4; shape = (50,) * 4; sh = _lib.shape_array(shape); n = 50 ** 4
A = np.repeat(gold["cfg2_A"][None], B, 0); b = np.repeat(gold["cfg2_b"][None], B, 0); c = np.repeat(gold["cfg2_c"].reshape(1), B, 0)
dA, db, dc = (torch.from_numpy(np.ascontiguousarray(x)).to(dev) for x in (A, b, c))
dG = torch.empty((B, n), dtype=torch.complex128, device=dev)
def run(): _lib.check(_lib.lib.mmh_forward_batched(B, 4, sh, dA.data_ptr(), db.data_ptr(), dc.data_ptr(), dG.data_ptr(), 0, None))
for _ in range(3): run()
torch.cuda.synchronize()
out = (ctypes.c_ulonglong * 256)()
run(); torch.cuda.synchronize()
assert lib.mmh_debug_timeline(out) == 0
t = np.array(list(out), dtype=np.int64).reshape(4, 16, 4)
t0 = t[0, 8, 0]
for l in range(B):
    for slot, name in ((8, "tail"), (1, "stage 1"), (0, "stage 0")):
        e = (t[l, slot] - t0) / 1e3
        print(f"lattice {l} {name:8s} entry {e[0]:8.2f}  wait passed {e[1]:8.2f}  first step {e[2]:8.2f}  exit(CTA 0) {e[3]:8.2f}")
