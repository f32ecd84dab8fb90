#!/usr/bin/env python
"""Summarise an `ncu --page source --csv` dump: per-opcode executed-instruction shares and the hottest SASS lines.

    ncu -i X.ncu-rep --page source --csv > src.csv ; python scripts/ncu_src_summary.py src.csv [--dump]
"""
import csv
import sys
from collections import Counter

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ci = {h: i for i, h in enumerate(hdr)}
ins = []
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    ins.append((r[ci["Address"]], r[ci["Source"]].strip(), int(r[ci["Instructions Executed"]]), int(r[ci["# Samples"]]),
                float(r[ci["Avg. Threads Executed"]] or 0)))
tot = sum(n for _, _, n, _, _ in ins)
tots = sum(s for _, _, _, s, _ in ins)
print(f"{len(ins)} SASS instructions, {tot} warp-instructions executed, {tots} stall samples")
ops = Counter(); ops_s = Counter()
for _, s, n, smp, _ in ins:
    t = s.split()
    op = t[1] if t and t[0].startswith("@") and len(t) > 1 else (t[0] if t else "?")
    op = op.split(".")[0]
    ops[op] += n; ops_s[op] += smp
print("opcode share of executed warp-instructions (top 25):")
for op, n in ops.most_common(25):
    print(f"  {op:10s} {100.0 * n / tot:6.2f}%   stall-samples {100.0 * ops_s[op] / max(tots, 1):6.2f}%")
if "--dump" in sys.argv:
    mx = max(n for _, _, n, _, _ in ins)
    thr = float(sys.argv[sys.argv.index("--dump") + 1]) if len(sys.argv) > sys.argv.index("--dump") + 1 else 0.3
    for a, s, n, smp, thr_ex in ins:
        if n >= thr * mx:
            print(f"{n:10d} {smp:6d} {thr_ex:5.1f}  {s}")
