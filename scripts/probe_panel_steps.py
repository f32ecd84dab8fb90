import os, sys
import numpy as np, torch
sys.path.insert(0, "/root/repo")
sys.path.insert(0, os.getcwd())
from mrmustard_b200 import _lib
dev = torch.device("cuda:0")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
rng = np.random.default_rng(3)
for shape in [(56,)*4, (20,64,64,64), (10,100,100,100), (30,64,64,64)]:
    D = len(shape)
    A = rng.uniform(-1, 1, (D, D)) + 1j * rng.uniform(-1, 1, (D, D)); A = (A + A.T) / 2; A /= np.abs(np.linalg.eigvals(A)).max() * 1.5
    b = rng.uniform(-1, 1, D) + 1j * rng.uniform(-1, 1, D); c = np.array([0.4 + 0.3j])
    dA, db, dc = (torch.from_numpy(np.ascontiguousarray(x)).to(dev) for x in (A, b, c))
    sh = _lib.shape_array(shape)
    out = {}
    for mode in ("default", "coop"):
        if mode == "coop": os.environ["MMH_FORCE_COOP"] = "1"
        else: os.environ.pop("MMH_FORCE_COOP", None)
        G = torch.full(shape, float("nan"), dtype=torch.complex128, device=dev)
        def run(): _lib.check(_lib.lib.mmh_forward(D, sh, dA.data_ptr(), db.data_ptr(), dc.data_ptr(), G.data_ptr(), 0, None))
        for _ in range(2): run()
        torch.cuda.synchronize()
        ts = []
        for _ in range(5):
            flush.fill_(1); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); run(); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        out[mode] = (G.clone(), np.median(ts))
    same = torch.equal(out["default"][0].view(torch.float64).view(torch.int64), out["coop"][0].view(torch.float64).view(torch.int64))
    n = int(np.prod(shape))
    print(f"{shape}: default {out['default'][1]:.1f} us ({16*n/out['default'][1]/1e3:.0f} GB/s)   coop {out['coop'][1]:.1f} us   bit-identical {same}", flush=True)
