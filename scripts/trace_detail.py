#!/usr/bin/env python
"""Step-by-step timeline of a few tiles of the cfg2 stage-0 tiled march (debug aid)."""
import os, sys, struct
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["MMH_TRACE_FILE"] = "gpurun_out/trace"
os.makedirs("gpurun_out", exist_ok=True)
from mrmustard_b200 import strategies
gold = np.load("tests/golden/vanilla_golden.npz")
A, b, c = gold["cfg2_A"], gold["cfg2_b"], complex(gold["cfg2_c"])
strategies.vanilla_numba((50,) * 4, A, b, c); strategies.vanilla_numba((50,) * 4, A, b, c)
raw = open("gpurun_out/trace.stage0.bin", "rb").read()
ntiles, S, g0, g1, g2, R, tc, _ = struct.unpack("8i", raw[:32])
t = np.frombuffer(raw[32:], dtype=np.uint64).reshape(ntiles, S, 8).astype(np.int64)
t0 = t[t > 0].min(); rel = np.where(t > 0, (t - t0) / 1e3, np.nan)
for tile in [int(x) for x in sys.argv[1:]] or [0, 31, 62]:
    lower = [tile - g1 * g2, tile - g2, tile - 1]
    print(f"tile {tile} (lower neighbours {lower}): s | start  pre-barrier  end | halo(s-1) published | lower neighbours' end(s-1)")
    for s in list(range(2, 12)) + [20, 30, 40]:
        nb = " ".join(f"{rel[n, s - 1, 1]:7.2f}" for n in lower if n >= 0)
        print(f"  {s:2d} | {rel[tile, s, 0]:7.2f} {rel[tile, s, 2]:7.2f} {rel[tile, s, 1]:7.2f} | {rel[tile, s - 1, 3]:7.2f} | {nb}")
