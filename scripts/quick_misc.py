#!/usr/bin/env python
"""Timings of the secondary entry points (stable, binomial, diagonal, leftover, Jacobians) — debug aid."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mrmustard_b200 as mm
from mrmustard_b200 import _lib
lib, check = _lib.lib, _lib.check
dev = torch.device("cuda:0")
gold = np.load("tests/golden/vanilla_golden.npz"); gd = np.load("tests/golden/diagonal_golden.npz")
def dev_time(fn, reps=5):
    fn(); torch.cuda.synchronize(); ms = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ms.append(a.elapsed_time(b))
    return float(np.median(ms))
def wall(fn, reps=3):
    fn(); t = []
    for _ in range(reps):
        t0 = time.perf_counter(); fn(); t.append(time.perf_counter() - t0)
    return 1e3 * min(t)
A, b, c = (torch.from_numpy(np.ascontiguousarray(gold[k])).to(dev) for k in ("cfg2_A", "cfg2_b", "cfg2_c")); c = c.reshape(1)
for shape in [(50,) * 4, (1000, 1000)]:
    if len(shape) == 2:
        A, b, c = (torch.from_numpy(np.ascontiguousarray(gold[k])).to(dev) for k in ("st_dg_A", "st_dg_b", "st_dg_c")); c = c.reshape(1)
    G = torch.empty(shape, dtype=torch.complex128, device=dev); sh = _lib.shape_array(shape)
    for stable in (0, 1):
        ms = dev_time(lambda: check(lib.mmh_forward(len(shape), sh, A.data_ptr(), b.data_ptr(), c.data_ptr(), G.data_ptr(), stable, None)))
        print(f"forward shape {shape} stable={stable}: {ms:.3f} ms  {np.prod(shape)/ms/1e6:.2f} G amp/s")
print(f"binomial (10,10) host call: {wall(lambda: mm.strategies.binomial((10, 10), gold['bin_A'], gold['bin_b'], complex(gold['bin_c']), 0.9, 15)):.3f} ms")
print(f"binomial (60,60,60) host call: {wall(lambda: mm.strategies.binomial((60, 60, 60), *[x for x in (np.eye(3)*0.1+0.05, np.ones(3)*0.2, 0.5)], 2.0, 200)):.3f} ms")
for name in ("d3b", "d4"):
    A_, b_, c_ = gd[f"{name}_A"], gd[f"{name}_b"], complex(gd[f"{name}_c"]); cut = tuple(int(x) for x in gd[f"{name}_cut"])
    print(f"diagonal {cut} host call: {wall(lambda: mm.hermite_renormalized_diagonal(A_, b_, c_, cut)):.3f} ms")
A_, b_, c_ = gd["d4_A"], gd["d4_b"], complex(gd["d4_c"])
print(f"diagonal (12,)*4 host call: {wall(lambda: mm.hermite_renormalized_diagonal(A_, b_, c_, (12,) * 4)):.3f} ms")
A_, b_, c_ = gd["l4_A"], gd["l4_b"], complex(gd["l4_c"])
print(f"1leftover oc=11 pnr=(11,11,11) host call: {wall(lambda: mm.hermite_renormalized_1leftoverMode(A_, b_, c_, 11, (11, 11, 11))):.3f} ms")
A_, b_, c_ = gd["d3_A"], gd["d3_b"], complex(gd["d3_c"])
A2, b2 = (np.ascontiguousarray(x) for x in mm.backend.reorder_AB_bargmann(A_, b_))
print(f"diagonal jacobians (12,12,12) host call: {wall(lambda: mm.strategies.grad_hermite_multidimensional_diagonal(A2, b2, c_, np.empty((12, 12, 12), complex))):.3f} ms")
