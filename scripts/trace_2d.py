#!/usr/bin/env python
"""Per-tile step timing of a (1000,1000) lattice's tiled stage 0 (debug aid)."""
import os, sys, struct
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
os.environ["MMH_TRACE_FILE"] = "gpurun_out/trace2d"
os.makedirs("gpurun_out", exist_ok=True)
from conftest import random_triple
from mrmustard_b200 import strategies
A, b, c = random_triple(2, (), seed=3); A = A * 0.5
strategies.vanilla_numba((1000, 1000), A, b, complex(c)); strategies.vanilla_numba((1000, 1000), A, b, complex(c))
raw = open("gpurun_out/trace2d.stage0.bin", "rb").read()
ntiles, S, g0, g1, g2, R, tc, _ = struct.unpack("8i", raw[:32])
t = np.frombuffer(raw[32:], dtype=np.uint64).reshape(ntiles, S, 8).astype(np.int64)
t0 = t[t > 0].min(); rel = np.where(t > 0, (t - t0) / 1e3, np.nan)
print(f"tiles {ntiles} S={S} R={R} tc={tc}; total {np.nanmax(rel):.1f} us")
for tile in sorted(set([0, 1, 2, ntiles // 2, ntiles - 1])):
    a = rel[tile, 100:900]
    print(" tile %3d: period %.3f us  start->accumulated %.3f  ->divided %.3f  ->prebar %.3f  ->after wait %.3f   halo publish - free %.3f" % (
        tile, (rel[tile, 900, 0] - rel[tile, 100, 0]) / 800, np.nanmean(a[:, 6] - a[:, 0]), np.nanmean(a[:, 7] - a[:, 6]),
        np.nanmean(a[:, 2] - a[:, 7]), np.nanmean(a[:, 1] - a[:, 2]), np.nanmean(a[:, 3] - a[:, 5])))
