#!/usr/bin/env python
"""stable=True probe on one lattice: box wavefront (k_stable_boxes) against the level wavefront with grid barriers (MMH_NO_STABLE_BOXES=1)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mrmustard_b200 import _lib
shapes = [tuple(int(x) for x in a.split(",")) for a in sys.argv[1:]] or [(50,) * 4]
dev = torch.device("cuda:0")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for shape in shapes:
    D = len(shape)
    rng = np.random.default_rng(7 + D)
    A = rng.uniform(-1, 1, (D, D)) + 1j * rng.uniform(-1, 1, (D, D)); A = (A + A.T) / 2; A /= np.abs(np.linalg.eigvals(A)).max() * 1.5
    b = rng.uniform(-1, 1, D) + 1j * rng.uniform(-1, 1, D); c = np.array([0.4 + 0.3j])
    dA, db, dc = (torch.from_numpy(np.ascontiguousarray(x)).to(dev) for x in (A, b, c))
    sh = _lib.shape_array(shape)
    out = {}
    for mode in ("levels", "boxes"):
        if mode == "levels": os.environ["MMH_NO_STABLE_BOXES"] = "1"
        else: os.environ.pop("MMH_NO_STABLE_BOXES", None)
        G = torch.full(shape, float("nan"), dtype=torch.complex128, device=dev)
        def run(): _lib.check(_lib.lib.mmh_forward(D, sh, dA.data_ptr(), db.data_ptr(), dc.data_ptr(), G.data_ptr(), 1, None))
        for _ in range(2): run()
        torch.cuda.synchronize()
        ts = []
        for _ in range(6):
            flush.fill_(1); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); run(); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        out[mode] = (G.clone(), np.median(ts))
    same = torch.equal(out["levels"][0].view(torch.float64).view(torch.int64), out["boxes"][0].view(torch.float64).view(torch.int64))
    print(f"{shape}: levels {out['levels'][1]:.1f} us   boxes {out['boxes'][1]:.1f} us   bit-identical {same}", flush=True)
