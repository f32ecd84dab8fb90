"""One-GPU prediction of the cfg3 strong-scaling curve: the per-rank shard sizes of N = 1, 2, 4, 8 ranks timed back to back on one
device (forward and VJP), with the launch variants (CTA size, programmatic dependent launch) selectable through the environment.
    python scripts/strong_scaling_probe.py [steps]
"""
import ctypes, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import random_triple
from mrmustard_b200 import _lib
lib, check = _lib.lib, _lib.check
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 40
dev = torch.device("cuda:0")
A, b, c = random_triple(2, (65536,), seed=673)
sh = _lib.shape_array((40, 40))
st = torch.cuda.current_stream(); sp = ctypes.c_void_p(st.cuda_stream)
base = {}
for env in ({}, {"MMH_NO_PDL": "1"}, {"MMH_LANES_BLOCK": "128"}, {"MMH_LANES_BLOCK": "64"}, {"MMH_LANES_BLOCK": "32"}, {"MMH_NO_FUSE_CHAIN": "1"}):
    for k in ("MMH_NO_PDL", "MMH_LANES_BLOCK", "MMH_NO_FUSE_CHAIN"):
        os.environ.pop(k, None)
    os.environ.update(env)
    row = []
    for N in (1, 2, 4, 8):
        B = 65536 // N
        dA, db, dc = (torch.from_numpy(np.ascontiguousarray(x[:B])).to(dev) for x in (A, b, c))
        dG = torch.empty((B, 1600), dtype=torch.complex128, device=dev)
        dg = torch.randn((B, 1600), dtype=torch.float64, device=dev).to(torch.complex128)
        oA = torch.empty((B, 2, 2), dtype=torch.complex128, device=dev); ob = torch.empty((B, 2), dtype=torch.complex128, device=dev); oc = torch.empty((B,), dtype=torch.complex128, device=dev)
        f = lambda: check(lib.mmh_forward_batched(B, 2, sh, dA.data_ptr(), db.data_ptr(), dc.data_ptr(), dG.data_ptr(), 0, sp))
        v = lambda: check(lib.mmh_vjp_batched(B, 2, sh, dG.data_ptr(), dc.data_ptr(), dg.data_ptr(), oA.data_ptr(), ob.data_ptr(), oc.data_ptr(), sp))
        res = []
        for fn in (f, v):
            for _ in range(5): fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st)
            for _ in range(steps): fn()
            e1.record(st); torch.cuda.synchronize()
            res.append(e0.elapsed_time(e1) / steps * 1e3)
        row.append(res)
        del dG, dg
    f1, v1 = row[0]
    print(f"{str(env):32s} " + "  ".join(f"N={n}: fwd {r[0]:6.1f} us (eff {f1 / n / r[0]:.3f}) vjp {r[1]:6.1f} us (eff {v1 / n / r[1]:.3f})" for n, r in zip((1, 2, 4, 8), row)))
