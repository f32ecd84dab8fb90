#!/usr/bin/env python
"""Hop anatomy of the cfg2 tiled march (debug aid): for a few tiles, per step: producers' end(s-1), halo warp's
buffer-free time, canary-pass time, publish time, consumer start(s).  Needs MMH_TRACE_FILE support (8 stamps per step)."""
import os, sys, struct
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["MMH_TRACE_FILE"] = "gpurun_out/trace"
os.makedirs("gpurun_out", exist_ok=True)
from mrmustard_b200 import strategies
gold = np.load("tests/golden/vanilla_golden.npz")
A, b, c = gold["cfg2_A"], gold["cfg2_b"], complex(gold["cfg2_c"])
strategies.vanilla_numba((50,) * 4, A, b, c); strategies.vanilla_numba((50,) * 4, A, b, c)
for stage in (1, 0):
    raw = open(f"gpurun_out/trace.stage{stage}.bin", "rb").read()
    ntiles, S, g0, g1, g2, R, tc, _ = struct.unpack("8i", raw[:32])
    t = np.frombuffer(raw[32:], dtype=np.uint64).reshape(ntiles, S, 8).astype(np.int64)
    t0 = t[t > 0].min(); rel = np.where(t > 0, (t - t0) / 1e3, np.nan)
    print(f"stage {stage}: tiles {ntiles} grid {g0}x{g1}x{g2} S={S} R={R} tc={tc}; total {np.nanmax(rel):.1f} us")
    for tile in range(0, ntiles, max(1, ntiles // 9)):
        print(" tile %3d end(s=1) %.2f end(last) %.2f  mean start->prebar %.3f  mean barrier %.3f" % (
            tile, rel[tile, 1, 1], rel[tile, S - 1, 1], np.nanmean(rel[tile, 1:, 2] - rel[tile, 1:, 0]), np.nanmean(rel[tile, 1:, 1] - rel[tile, 1:, 2])))
    tiles = [ntiles - 1, ntiles // 2] if stage == 0 else [ntiles - 1]
    for tile in tiles:
        tt = [tile // (g1 * g2), (tile // g2) % g1, tile % g2]
        lower = [n for n, cc in zip([tile - g1 * g2, tile - g2, tile - 1], tt) if cc > 0]
        print(f" tile {tile} {tt}: s | producers' end(s-1) | buffer free, canary, publish of halo(s-1) | start(s) prebar(s) end(s)")
        for s in list(range(2, 14)) + [20, 30, 40, S - 1]:
            if s >= S: continue
            nb = " ".join("%7.2f" % rel[n, s - 1, 1] for n in lower)
            print("  %2d | %s | %7.2f %7.2f %7.2f | %7.2f %7.2f %7.2f" % (s, nb, rel[tile, s - 1, 5], rel[tile, s - 1, 4], rel[tile, s - 1, 3],
                                                                      rel[tile, s, 0], rel[tile, s, 2], rel[tile, s, 1]))
    if stage == 0:
        for tile in (0, ntiles // 2):
            a = rel[tile, 5:45]
            print(" tile %d phases (us): start->accumulated %.3f  ->divided %.3f  ->stores+pre (prebar) %.3f  ->barrier %.3f" % (
                tile, np.nanmean(a[:, 6] - a[:, 0]), np.nanmean(a[:, 7] - a[:, 6]), np.nanmean(a[:, 2] - a[:, 7]), np.nanmean(a[:, 1] - a[:, 2])))
