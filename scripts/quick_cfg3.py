#!/usr/bin/env python
"""cfg3 forward (65,536 x (40,40)) device-resident timing: warp-synchronous lane march vs the shared-memory stage march."""
import hashlib, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from conftest import random_triple
from mrmustard_b200 import _lib
dev = torch.device("cuda:0")
shape = tuple(int(x) for x in (sys.argv[1].split(",") if len(sys.argv) > 1 else (40, 40)))
B = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
D = len(shape); n = int(np.prod(shape)); sh = _lib.shape_array(shape)
A, b, c = random_triple(D, (B,), seed=673)
dA, db, dc = (torch.from_numpy(np.ascontiguousarray(x)).to(dev) for x in (A, b, c))
dG = torch.empty((B, n), dtype=torch.complex128, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def run(): _lib.check(_lib.lib.mmh_forward_batched(B, D, sh, dA.data_ptr(), db.data_ptr(), dc.data_ptr(), dG.data_ptr(), 0, None))
hs = []
for tag, env in (("stage march", {"MMH_NO_LANES": "1"}), ("lane, old chain", {"MMH_NO_CHAIN_ROWS": "1"}), ("lane march", {}), ("lane R=4", {"MMH_LANES_R": "4"})):
    for k in ("MMH_NO_LANES", "MMH_LANES_R", "MMH_NO_CHAIN_ROWS"): os.environ.pop(k, None)
    os.environ.update(env)
    for _ in range(3): run()
    torch.cuda.synchronize(); ms = []
    for _ in range(10):
        flush.fill_(1)
        a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); run(); e.record(); torch.cuda.synchronize(); ms.append(a.elapsed_time(e))
    h = hashlib.sha256(dG.cpu().numpy().tobytes()).hexdigest(); hs.append(h)
    print(f"{shape} x {B} {tag:18s}: median {np.median(ms):.4f} ms  min {min(ms):.4f} ms  {B*n/np.median(ms)/1e6:.1f} G amp/s  "
          f"hbm frac {16*B*n/np.median(ms)/1e6/6534.8:.3f}  {'same bits' if h == hs[0] else 'MISMATCH'}", flush=True)
