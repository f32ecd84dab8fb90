#!/usr/bin/env python
"""Per-tile timeline of one cfg2 forward WITH stage overlap (MMH_TRACE_OVERLAP=1): when does each stage-0 box start / finish, and
how long are its steps?  Prints a t1 x (t2+t3) table of start and end times."""
import os, sys, struct
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["MMH_TRACE_FILE"] = "gpurun_out/trace_ov"
os.environ["MMH_TRACE_OVERLAP"] = "1"
os.makedirs("gpurun_out", exist_ok=True)
from mrmustard_b200 import strategies
gold = np.load("tests/golden/vanilla_golden.npz")
A, b, c = gold["cfg2_A"], gold["cfg2_b"], complex(gold["cfg2_c"])
for _ in range(3): strategies.vanilla_numba((50,) * 4, A, b, c)
ts = {}
for stage in (1, 0):
    raw = open(f"gpurun_out/trace_ov.stage{stage}.bin", "rb").read()
    ntiles, S, g0, g1, g2, R, tc, _ = struct.unpack("8i", raw[:32])
    ts[stage] = (np.frombuffer(raw[32:], dtype=np.uint64).reshape(ntiles, S, 8).astype(np.int64), (g0, g1, g2), S)
t0 = min(t[t > 0].min() for t, _, _ in ts.values())
for stage in (1, 0):
    t, g, S = ts[stage]
    rel = np.where(t > 0, (t - t0) / 1e3, np.nan)
    print(f"stage {stage}: grid {g}, first step start {np.nanmin(rel[:, 1, 0]):.1f}, last end {np.nanmax(rel[:, S - 1, 2]):.1f} us")
    if stage == 0:
        g0, g1, g2 = g
        print(" tile (t0,t1,t2): start(s=1)  end(s=1)  end(s=10)  end(s=25)  end(last)   mean period(s>=25)")
        for tile in range(0, g0 * g1 * g2):
            tt = (tile // (g1 * g2), (tile // g2) % g1, tile % g2)
            if tt[1] == tt[2] and tt[1] in (0, 2, 4):
                r = rel[tile]
                print("  %s  %7.1f %7.1f %7.1f %7.1f %7.1f   %.3f" % (tt, r[1, 0], r[1, 2], r[10, 2], r[25, 2], r[S - 1, 2], (r[S - 1, 2] - r[25, 2]) / (S - 1 - 25)))
