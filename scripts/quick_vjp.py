#!/usr/bin/env python
"""Device-resident timing of the VJP kernels on cfg5 (4-mode ket, cutoff 40) and cfg3 (batched) — debug aid."""
import ctypes, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mrmustard_b200 import _lib
lib, check = _lib.lib, _lib.check
dev = torch.device("cuda:0")
gold = np.load("tests/golden/vanilla_golden.npz")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def timeit(fn, reps=10, fl=True):
    fn(); torch.cuda.synchronize(); ms = []
    for _ in range(reps):
        if fl: flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ms.append(a.elapsed_time(b))
    return float(np.median(ms))
which = sys.argv[1] if len(sys.argv) > 1 else "both"
if which in ("cfg5", "both"):
    A, b, c = (torch.from_numpy(np.ascontiguousarray(gold[k])).to(dev) for k in ("cfg5_A", "cfg5_b", "cfg5_c")); c = c.reshape(1)
    shape = (40,) * 4; sh = _lib.shape_array(shape); n = 40 ** 4
    G = torch.empty(n, dtype=torch.complex128, device=dev)
    check(lib.mmh_forward(4, sh, A.data_ptr(), b.data_ptr(), c.data_ptr(), G.data_ptr(), 0, None))
    g = torch.from_numpy(np.random.RandomState(1).standard_normal(shape) + 0j).to(dev)
    oA = torch.empty((4, 4), dtype=torch.complex128, device=dev); ob = torch.empty(4, dtype=torch.complex128, device=dev); oc = torch.empty(1, dtype=torch.complex128, device=dev)
    ms = timeit(lambda: check(lib.mmh_vjp(4, sh, G.data_ptr(), c.data_ptr(), g.data_ptr(), oA.data_ptr(), ob.data_ptr(), oc.data_ptr(), None)))
    ok = np.allclose(oA.cpu().numpy(), gold["cfg5_dA40"], rtol=1e-10, atol=1e-14) and np.allclose(ob.cpu().numpy(), gold["cfg5_db40"], rtol=1e-10, atol=1e-14)
    print(f"cfg5 vjp: {ms*1e3:.1f} us  {n/ms/1e6:.2f} G amp/s  hbm frac {32*n/ms/1e6/6534.8:.3f}  parity {'OK' if ok else 'MISMATCH'}")
if which in ("cfg3", "both"):
    B, n = 65536, 1600
    rng = np.random.RandomState(673)
    G = torch.randn((B, n), dtype=torch.float64, device=dev).to(torch.complex128) * 0.1
    g = torch.randn((B, n), dtype=torch.float64, device=dev).to(torch.complex128)
    c = torch.ones(B, dtype=torch.complex128, device=dev)
    sh = _lib.shape_array((40, 40))
    oA = torch.empty((B, 2, 2), dtype=torch.complex128, device=dev); ob = torch.empty((B, 2), dtype=torch.complex128, device=dev); oc = torch.empty(B, dtype=torch.complex128, device=dev)
    ref = None
    for tag, env in (("partial kernel", {"MMH_NO_LANES": "1"}), ("row walk", {}), ("row walk R=4", {"MMH_LANES_R": "4"}), ("row walk R=3", {"MMH_LANES_R": "3"})):
        for k in ("MMH_NO_LANES", "MMH_LANES_R"): os.environ.pop(k, None)
        os.environ.update(env)
        ms = timeit(lambda: check(lib.mmh_vjp_batched(B, 2, sh, G.data_ptr(), c.data_ptr(), g.data_ptr(), oA.data_ptr(), ob.data_ptr(), oc.data_ptr(), None)), 5, False)
        out = np.concatenate([oA.cpu().numpy().ravel(), ob.cpu().numpy().ravel(), oc.cpu().numpy().ravel()])
        if ref is None: ref = out
        err = np.max(np.abs(out - ref) / (1e-14 / 1e-10 + np.abs(ref)))
        print(f"cfg3 vjp {tag:14s}: {ms*1e3:.1f} us  {B*n/ms/1e6:.2f} G amp/s  hbm frac {32*B*n/ms/1e6/6534.8:.3f}  max rel dev vs partial kernel {err:.1e}", flush=True)
    for k in ("MMH_NO_LANES", "MMH_LANES_R"): os.environ.pop(k, None)
