"""GPU tests of the stream contract of the device-pointer entry points (include/mmhermite.h: "enqueue work on `stream` without
synchronising"): first use on a non-blocking stream, calls alternating between two streams that share the per-device scratch,
and one G buffer reused back-to-back with different triples (sentinel pre-fill, exchange-buffer slot rotation)."""
import ctypes
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT, assert_parity, random_triple, sha

pytestmark = pytest.mark.gpu


def _fwd(L, torch, shape, dA, db, dc, dG, st, batch=None):
    sh = L.shape_array(shape)
    sp = ctypes.c_void_p(st.cuda_stream)
    if batch is None:
        L.check(L.lib.mmh_forward(len(shape), sh, dA.data_ptr(), db.data_ptr(), dc.data_ptr(), dG.data_ptr(), 0, sp))
    else:
        L.check(L.lib.mmh_forward_batched(batch, len(shape), sh, dA.data_ptr(), db.data_ptr(), dc.data_ptr(), dG.data_ptr(), 0, sp))


def test_first_call_on_a_non_blocking_stream():
    """A fresh process whose very first library call is a tiled-march lattice on a cudaStreamNonBlocking stream: the sqrt tables,
    the exchange-buffer sentinel and the panel-0 sentinel must all be ordered on that stream (they used to be initialised on the
    legacy default stream, which such a stream does not synchronise with)."""
    code = r"""
import ctypes, sys, hashlib, numpy as np, torch
sys.path.insert(0, %r)
from mrmustard_b200 import _lib as L
gold = np.load(%r)
A, b, c = gold["cfg2_A"], gold["cfg2_b"], gold["cfg2_c"].reshape(1)
dev = torch.device("cuda:0")
dA, db, dc = (torch.from_numpy(np.ascontiguousarray(x)).to(dev) for x in (A, b, c))
dG = torch.empty(50 ** 4, dtype=torch.complex128, device=dev)
torch.cuda.synchronize()
st = torch.cuda.Stream()            # torch side streams are cudaStreamNonBlocking
busy = torch.empty(1 << 28, dtype=torch.uint8, device=dev)
busy.fill_(1)                       # keep the legacy default stream busy while the library initialises
L.check(L.lib.mmh_forward(4, L.shape_array((50,) * 4), dA.data_ptr(), db.data_ptr(), dc.data_ptr(), dG.data_ptr(), 0,
                          ctypes.c_void_p(st.cuda_stream)))
st.synchronize()
got = hashlib.sha256((dG.cpu().numpy() + 0.0).tobytes()).hexdigest()
assert got == str(gold["cfg2_G50_sha"]), "first call on a non-blocking stream is wrong"
print("ok")
""" % (ROOT, os.path.join(ROOT, "tests", "golden", "vanilla_golden.npz"))
    res = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600)
    assert res.returncode == 0 and "ok" in res.stdout, res.stdout + res.stderr


def test_two_streams_share_the_scratch_safely(golden):
    """Forward (tiled march: exchange buffer + sentinels) and VJP (partial-sum scratch) calls alternate between two streams with
    no host synchronisation in between; every result must be the single-stream result."""
    import torch
    from mrmustard_b200 import _lib as L
    import oracle
    dev = torch.device("cuda:0")
    shape = (24, 25, 26, 27)
    n = int(np.prod(shape))
    trip = [random_triple(4, (), seed=40 + k) for k in range(4)]
    want = [oracle.vanilla(shape, A, b, complex(c)) for A, b, c in trip]
    g = np.random.RandomState(3).standard_normal(shape) + 0j
    wvjp = [oracle.vanilla_vjp(w, complex(t[2]), g) for w, t in zip(want, trip)]
    dT = [tuple(torch.from_numpy(np.ascontiguousarray(np.asarray(x, dtype=np.complex128).reshape(-1))).to(dev) for x in t) for t in trip]
    dg = torch.from_numpy(g.reshape(-1)).to(dev)
    dG = [torch.empty(n, dtype=torch.complex128, device=dev) for _ in trip]
    oA = [torch.empty(16, dtype=torch.complex128, device=dev) for _ in trip]
    ob = [torch.empty(4, dtype=torch.complex128, device=dev) for _ in trip]
    oc = [torch.empty(1, dtype=torch.complex128, device=dev) for _ in trip]
    s = [torch.cuda.Stream(), torch.cuda.Stream()]
    torch.cuda.synchronize()
    sh = L.shape_array(shape)
    for rep in range(3):
        for k in range(4):
            st = s[k & 1]
            _fwd(L, torch, shape, *dT[k], dG[k], st)
            L.check(L.lib.mmh_vjp(4, sh, dG[k].data_ptr(), dT[k][2].data_ptr(), dg.data_ptr(), oA[k].data_ptr(), ob[k].data_ptr(),
                                  oc[k].data_ptr(), ctypes.c_void_p(st.cuda_stream)))
    torch.cuda.synchronize()
    for k in range(4):
        assert np.array_equal(dG[k].cpu().numpy().reshape(shape), want[k]), k
        assert_parity(oA[k].cpu().numpy().reshape(4, 4), wvjp[k][0], f"dLdA {k}")
        assert_parity(ob[k].cpu().numpy(), wvjp[k][1], f"dLdb {k}")
        assert_parity(oc[k].cpu().numpy()[0], np.complex128(wvjp[k][2]), f"dLdc {k}")


def test_one_buffer_reused_with_different_triples(golden):
    """The same G buffer filled back-to-back (no host sync) by different triples, single lattices and a pipelined batch: a stale
    amplitude of the previous lattice must never be taken for a delivered one (the panel-0 sentinel pre-fill and the exchange-buffer
    slot rotation are what guarantee it)."""
    import torch
    from mrmustard_b200 import _lib as L
    import oracle
    dev = torch.device("cuda:0")
    shape = (24, 25, 26, 27)
    n = int(np.prod(shape))
    st = torch.cuda.current_stream()
    A, b, c = random_triple(4, (7,), seed=77)
    dA, db, dc = (torch.from_numpy(np.ascontiguousarray(x)).to(dev) for x in (A, b, c))
    dG = torch.empty(n, dtype=torch.complex128, device=dev)
    outs = []
    for k in range(7):   # single lattices into ONE buffer, copied out on the same stream right after each fill
        _fwd(L, torch, shape, dA[k], db[k], dc[k:k + 1], dG, st)
        outs.append(dG.clone())
    torch.cuda.synchronize()
    for k in range(7):
        assert np.array_equal(outs[k].cpu().numpy().reshape(shape), oracle.vanilla(shape, A[k], b[k], complex(c[k]))), k
    # pipelined batch (7 lattices cross the every-4th full-wait boundary) into one buffer, twice with permuted triples
    dGb = torch.empty((7, n), dtype=torch.complex128, device=dev)
    perm = [3, 0, 6, 1, 5, 2, 4]
    _fwd(L, torch, shape, dA, db, dc, dGb, st, batch=7)
    first = dGb.clone()
    dA2, db2, dc2 = dA[perm].contiguous(), db[perm].contiguous(), dc[perm].contiguous()
    _fwd(L, torch, shape, dA2, db2, dc2, dGb, st, batch=7)
    torch.cuda.synchronize()
    f, sec = first.cpu().numpy(), dGb.cpu().numpy()
    for k in range(7):
        want = oracle.vanilla(shape, A[k], b[k], complex(c[k])).reshape(-1)
        assert np.array_equal(f[k], want), k
        assert np.array_equal(sec[perm.index(k)], want), k


def test_cfg2_repeated_on_side_stream(golden):
    import torch
    from mrmustard_b200 import _lib as L
    dev = torch.device("cuda:0")
    A, b, c = golden["cfg2_A"], golden["cfg2_b"], golden["cfg2_c"].reshape(1)
    dA, db, dc = (torch.from_numpy(np.ascontiguousarray(x)).to(dev) for x in (A, b, c))
    dG = torch.empty(50 ** 4, dtype=torch.complex128, device=dev)
    st = torch.cuda.Stream()
    for _ in range(5):
        _fwd(L, torch, (50,) * 4, dA, db, dc, dG, st)
    st.synchronize()
    assert sha(dG.cpu().numpy()) == str(golden["cfg2_G50_sha"])
