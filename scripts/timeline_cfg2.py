#!/usr/bin/env python
"""Kernel-level timeline of one cfg2 forward from the device stamps (mmh_debug_timeline): entry / dependency wait passed /
first step / exit of CTA 0 for the trailing-stage kernel (slot 8) and the tiled stages (slots = stage index)."""
import ctypes, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mrmustard_b200 import _lib
lib = ctypes.CDLL(_lib.SO_PATH)
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
out = (ctypes.c_ulonglong * 256)()
for rep in range(3):
    flush.fill_(1); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); run(); b.record(); torch.cuda.synchronize()
    assert lib.mmh_debug_timeline(out) == 0
    t = np.array(list(out), dtype=np.int64).reshape(4, 16, 4)[0]
    t0 = t[8, 0]
    print(f"rep {rep}: event time {a.elapsed_time(b)*1e3:.1f} us")
    for slot, name in ((8, "tail (chain + stage 2)"), (1, "stage 1 tiled"), (0, "stage 0 tiled")):
        e = (t[slot] - t0) / 1e3
        print(f"   {name:24s} entry {e[0]:7.2f}  wait passed {e[1]:7.2f}  first step {e[2]:7.2f}  exit(CTA 0) {e[3]:7.2f}")
