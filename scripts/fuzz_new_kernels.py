#!/usr/bin/env python
"""Randomised cross-check of the round-2 single-lattice kernels against the kernels they replace (same library, switched by the
environment hooks): k_march_rows vs k_march_tiled2 and k_stable_boxes vs k_stable_coop bit for bit, k_vjp_planes vs k_vjp_partial
within 1e-11 relative.  Shapes are drawn so that every kernel's ragged-edge handling is exercised."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mrmustard_b200 import _lib
dev = torch.device("cuda:0")
rng = np.random.default_rng(int(sys.argv[1]) if len(sys.argv) > 1 else 1)

def triple(D):
    A = rng.uniform(-1, 1, (D, D)) + 1j * rng.uniform(-1, 1, (D, D)); A = (A + A.T) / 2; A /= np.abs(np.linalg.eigvals(A)).max() * 1.4
    b = rng.uniform(-1, 1, D) + 1j * rng.uniform(-1, 1, D); c = np.array([0.4 + 0.3j])
    return tuple(torch.from_numpy(np.ascontiguousarray(x)).to(dev) for x in (A, b, c))

def forward(shape, t, stable, env):
    for k in ("MMH_NO_ROWS", "MMH_NO_STABLE_BOXES"): os.environ.pop(k, None)
    os.environ.update(env)
    G = torch.full(shape, float("nan"), dtype=torch.complex128, device=dev)
    _lib.check(_lib.lib.mmh_forward(len(shape), _lib.shape_array(shape), t[0].data_ptr(), t[1].data_ptr(), t[2].data_ptr(), G.data_ptr(), stable, None))
    torch.cuda.synchronize()
    return G

def same(a, b): return torch.equal(a.view(torch.float64).view(torch.int64), b.view(torch.float64).view(torch.int64))

bad = 0
for it in range(int(os.environ.get("FUZZ_N", "6"))):
    # vanilla, boxes of >= 700 cells: 4-index lattices with a panel of 90k .. 150k cells
    shape = (int(rng.integers(3, 12)),) + tuple(int(x) for x in rng.integers(44, 54, 3))
    t = triple(4)
    ok = same(forward(shape, t, 0, {}), forward(shape, t, 0, {"MMH_NO_ROWS": "1"}))
    print("rows  ", shape, ok, flush=True); bad += not ok
    # stable rule, 2 / 3 / 4 indices
    D = int(rng.integers(2, 5))
    shape = tuple(int(x) for x in rng.integers({2: 300, 3: 50, 4: 18}[D], {2: 700, 3: 90, 4: 30}[D], D))
    t = triple(D)
    ok = same(forward(shape, t, 1, {}), forward(shape, t, 1, {"MMH_NO_STABLE_BOXES": "1"}))
    print("stable", shape, ok, flush=True); bad += not ok
    # VJP planes
    shape = tuple(int(x) for x in rng.integers(20, 45, 4))
    if shape[2] * shape[3] > 1800: shape = shape[:2] + (40, 40)
    G = torch.from_numpy(rng.standard_normal(shape) * 1e-2 + 1j * rng.standard_normal(shape) * 1e-2).to(dev)
    g = torch.from_numpy(rng.standard_normal(shape) + 1j * rng.standard_normal(shape)).to(dev)
    c = torch.tensor([0.4 + 0.3j], dtype=torch.complex128, device=dev)
    outs = []
    for env in ({}, {"MMH_NO_VJP_PLANES": "1"}):
        os.environ.pop("MMH_NO_VJP_PLANES", None); os.environ.update(env)
        dA = torch.empty((4, 4), dtype=torch.complex128, device=dev); db = torch.empty(4, dtype=torch.complex128, device=dev); dc = torch.empty(1, dtype=torch.complex128, device=dev)
        _lib.check(_lib.lib.mmh_vjp(4, _lib.shape_array(shape), G.data_ptr(), c.data_ptr(), g.data_ptr(), dA.data_ptr(), db.data_ptr(), dc.data_ptr(), None))
        torch.cuda.synchronize(); outs.append((dA, db, dc))
    ok = all(torch.allclose(x, y, rtol=1e-11, atol=1e-14) for x, y in zip(*outs))
    print("vjp   ", shape, ok, flush=True); bad += not ok
print("mismatches:", bad)
sys.exit(1 if bad else 0)
