#!/usr/bin/env python
"""Per-step timeline of the tiled march on a small lattice with forced tile grids (debug aid)."""
import os, sys, struct
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["MMH_TRACE_FILE"] = "gpurun_out/trace_s"
os.environ["MMH_FORCE_TILED"] = "1"
os.makedirs("gpurun_out", exist_ok=True)
from mrmustard_b200 import strategies
rng = np.random.RandomState(0)
shape = tuple(int(x) for x in sys.argv[1].split(","))
n = len(shape)
A = rng.random((n, n)) + 1j * rng.random((n, n)); A = A + A.T; A /= np.abs(np.linalg.eigvals(A)).max() + 0.2
b = rng.random(n) + 1j * rng.random(n); c = 0.3 + 0.1j
for grid in sys.argv[2:]:
    os.environ["MMH_TILE_G"] = grid
    strategies.vanilla_numba(shape, A, b, c); strategies.vanilla_numba(shape, A, b, c)
    raw = open("gpurun_out/trace_s.stage0.bin", "rb").read()
    ntiles, S, g0, g1, g2, R, tc, _ = struct.unpack("8i", raw[:32])
    t = np.frombuffer(raw[32:], dtype=np.uint64).reshape(ntiles, S, 8).astype(np.int64)
    t0 = t[t > 0].min(); rel = np.where(t > 0, t - t0, -1)
    print(f"shape {shape} grid {g0}x{g1}x{g2} tiles {ntiles} R={R} tc={tc} total {rel.max()/1e3:.1f} us")
    for tile in sorted(set([0, ntiles - 1])):
        start, end, pre = rel[tile, :, 0], rel[tile, :, 1], rel[tile, :, 2]
        print(f"   tile {tile}: mean step {np.diff(end[1:]).mean()/1e3:.3f} us; math+stores {np.mean(pre[2:]-start[2:])/1e3:.3f}; barrier {np.mean(end[2:]-pre[2:])/1e3:.3f}")
