#!/usr/bin/env python
"""cfg5 VJP probe: k_vjp_planes against k_vjp_partial (MMH_NO_VJP_PLANES=1) on one (40,)^4 lattice: CUDA-event time with an L2
flush between runs, and agreement of the two results."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mrmustard_b200 import _lib
shape = tuple(int(x) for x in sys.argv[1].split(",")) if len(sys.argv) > 1 else (40,) * 4
D = len(shape)
dev = torch.device("cuda:0")
rng = np.random.default_rng(5)
G = torch.from_numpy(rng.standard_normal(shape) * 1e-2 + 1j * rng.standard_normal(shape) * 1e-2).to(dev)
g = torch.from_numpy(rng.standard_normal(shape) + 1j * rng.standard_normal(shape)).to(dev)
c = torch.tensor([0.4 + 0.3j], dtype=torch.complex128, device=dev)
dA = torch.empty((D, D), dtype=torch.complex128, device=dev); db = torch.empty(D, dtype=torch.complex128, device=dev); dc = torch.empty(1, dtype=torch.complex128, device=dev)
sh = _lib.shape_array(shape)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
res = {}
for mode in ("partial", "planes"):
    if mode == "partial": os.environ["MMH_NO_VJP_PLANES"] = "1"
    else: os.environ.pop("MMH_NO_VJP_PLANES", None)
    def run(): _lib.check(_lib.lib.mmh_vjp(D, sh, G.data_ptr(), c.data_ptr(), g.data_ptr(), dA.data_ptr(), db.data_ptr(), dc.data_ptr(), None))
    for _ in range(3): run()
    ts = []
    for _ in range(10):
        flush.fill_(1); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); run(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    res[mode] = (dA.clone(), db.clone(), dc.clone(), np.median(ts), np.min(ts))
    print(f"{mode}: {np.median(ts):.1f} us (min {np.min(ts):.1f})  = {32 * np.prod(shape) / np.median(ts) / 1e3:.0f} GB/s", flush=True)
for a, b_ in zip(res["partial"][:3], res["planes"][:3]):
    print("max rel diff", float((a - b_).abs().max() / a.abs().max()))
