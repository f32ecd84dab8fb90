#!/usr/bin/env python
"""Run one cfg2 lattice with MMH_TRACE_FILE set and print the tile-pipeline timeline (debug aid)."""
import os, sys, struct
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["MMH_TRACE_FILE"] = "gpurun_out/trace"
os.makedirs("gpurun_out", exist_ok=True)
from mrmustard_b200 import strategies
gold = np.load("tests/golden/vanilla_golden.npz")
A, b, c = gold["cfg2_A"], gold["cfg2_b"], complex(gold["cfg2_c"])
strategies.vanilla_numba((50,) * 4, A, b, c)   # warm-up + trace (overwritten)
strategies.vanilla_numba((50,) * 4, A, b, c)
for stage in (1, 0):
    raw = open(f"gpurun_out/trace.stage{stage}.bin", "rb").read()
    hdr = struct.unpack("8i", raw[:32])
    ntiles, S, g0, g1, g2, R, tc, _ = hdr
    t = np.frombuffer(raw[32:], dtype=np.uint64).reshape(ntiles, S, 8).astype(np.int64)
    t0 = t[t > 0].min()
    rel = np.where(t > 0, t - t0, -1)
    print(f"stage {stage}: tiles {ntiles} grid {g0}x{g1}x{g2} S={S} R={R} tc={tc}; total {rel.max()/1e3:.1f} us")
    for tile in sorted(set([0, 1, g2, g1 * g2, ntiles // 2, ntiles - 1])):
        if tile >= ntiles: continue
        start, end, seen, pub = rel[tile, :, 0], rel[tile, :, 1], rel[tile, :, 2], rel[tile, :, 3]
        steps = [1, 2, 3, 4, 5, 10, 20, 30, 40, S - 2, S - 1]
        print(f"  tile {tile}: step: start/end/flagseen/publish (us)")
        for s in steps:
            if s < S: print(f"    s={s:2d}  {start[s]/1e3:8.2f} {end[s]/1e3:8.2f} {seen[s]/1e3:8.2f} {pub[s]/1e3:8.2f}")
        d = np.diff(end[1:])
        print(f"    mean step {d.mean()/1e3:.3f} us, compute (end-start) mean {np.mean(end[1:]-start[1:])/1e3:.3f} us, "
              f"thread0 math+stores {np.mean(seen[1:]-start[1:])/1e3:.3f} us, barrier wait {np.mean(end[1:]-seen[1:])/1e3:.3f} us")
