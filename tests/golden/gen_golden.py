"""Generate the committed golden vectors from the UNMODIFIED reference (run in the build container).

    python tests/golden/gen_golden.py

Imports /root/reference through oracle/refimport.py (shims only; no reference source is copied) and
writes small .npz fixtures next to this file.  The fixtures pin (1) the C oracle (tests/test_oracle_golden.py,
CPU) and (2) the CUDA path (tests/test_gpu_*.py) to the reference's own numba output on identical
input bytes.  Large outputs (cfg2 (50,)*4, cfg3 65,536x(40,40), cfg5 (40,)*4) are pinned by a sha256 of
the canonicalised bytes (x + 0.0 to fold -0.0 into +0.0) plus a strided sample.

Reference functions used (file:line under /root/reference/mrmustard/math/lattice/strategies/):
  vanilla/core.py:25 vanilla_numba, :127 stable_numba, vanilla/batch.py:27 vanilla_batch_numba,
  vanilla/gradients.py:25 vanilla_vjp_numba, :85 vanilla_batch_vjp_numba, binomial.py:30 binomial.
"""
from __future__ import annotations

import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import refimport  # noqa: E402


def sha(a: np.ndarray) -> str:
    a = np.ascontiguousarray(a) + 0.0  # fold signed zeros
    return hashlib.sha256(a.tobytes()).hexdigest()


def random_triple(n, batch=(), seed=None):
    """Same recipe as the reference's tests/test_math/test_lattice/test_vanilla.py:24-35 (restated)."""
    rng = np.random.RandomState(seed)
    A = rng.random((*batch, n, n)) + 1j * rng.random((*batch, n, n))
    A = A + np.swapaxes(A, -1, -2)
    A /= np.abs(np.linalg.eigvals(A)).max() + 0.2
    b = rng.random((*batch, n)) + 1j * rng.random((*batch, n))
    c = rng.random(batch) + 1j * rng.random(batch)
    return A, b, c


def main():
    refimport.install_shims(with_lab=True)
    from mrmustard import math, settings  # noqa: PLC0415
    from mrmustard.lab import BSgate, DisplacedSqueezed, Ggate, Ket, Sgate, Vacuum  # noqa: PLC0415
    from mrmustard.math.lattice import strategies as S  # noqa: PLC0415

    out = {}

    # ---- lab-derived triples for the BASELINE configs -----------------------------
    A1, b1, c1 = (np.asarray(x, dtype=np.complex128) for x in DisplacedSqueezed(0, r=0.5, alpha=0.3).bargmann_triple())
    u = BSgate((0, 1), theta=0.5, phi=0.2) >> Sgate(0, r=0.3) >> Sgate(1, r=0.2)
    A2, b2, c2 = (np.asarray(x, dtype=np.complex128) for x in u.bargmann_triple())
    ket5 = Vacuum((0, 1, 2, 3)) >> Ggate((0, 1, 2, 3), symplectic=math.random_symplectic(4))
    with settings(SEED=11):
        ket5 = Vacuum((0, 1, 2, 3)) >> Ggate((0, 1, 2, 3), symplectic=math.random_symplectic(4))
        A5, b5, c5 = (np.asarray(x, dtype=np.complex128) for x in ket5.bargmann_triple())
    with settings(SEED=42):
        Ak, bk, ck = (np.asarray(x, dtype=np.complex128) for x in Ket.random((0, 1)).bargmann_triple())
    with settings(SEED=3):
        A8, b8, c8 = (np.asarray(x, dtype=np.complex128) for x in Ket.random(tuple(range(8)), max_r=0.5).bargmann_triple())

    # cross-check the lab path once: fock_array == hermite_renormalized(triple)
    G1 = S.vanilla_numba((200,), A1, b1, complex(c1))
    assert np.array_equal(np.asarray(DisplacedSqueezed(0, r=0.5, alpha=0.3).fock_array(200)), G1)

    out.update(cfg1_A=A1, cfg1_b=b1, cfg1_c=c1, cfg1_G=G1,
               cfg1_G_stable=S.stable_numba((200,), A1, b1, complex(c1)))

    out.update(cfg2_A=A2, cfg2_b=b2, cfg2_c=c2)
    out["cfg2_G12"] = S.vanilla_numba((12,) * 4, A2, b2, complex(c2))
    out["cfg2_G12_stable"] = S.stable_numba((12,) * 4, A2, b2, complex(c2))
    G2 = S.vanilla_numba((50,) * 4, A2, b2, complex(c2))
    out["cfg2_G50_sha"] = np.array(sha(G2))
    out["cfg2_G50_sample"] = G2.ravel()[::9973].copy()
    G2s = S.stable_numba((50,) * 4, A2, b2, complex(c2))
    out["cfg2_G50_stable_sha"] = np.array(sha(G2s))
    out["cfg2_G50_stable_sample"] = G2s.ravel()[::9973].copy()
    # raw-kernel variant of cfg2: random_triple(4, (), seed=1)
    A2r, b2r, c2r = random_triple(4, (), seed=1)
    G2r = S.vanilla_numba((50,) * 4, A2r, b2r, complex(c2r))
    out.update(cfg2r_A=A2r, cfg2r_b=b2r, cfg2r_c=np.asarray(c2r), cfg2r_G50_sha=np.array(sha(G2r)),
               cfg2r_G50_sample=G2r.ravel()[::9973].copy())

    out.update(cfg5_A=A5, cfg5_b=b5, cfg5_c=c5)
    G5s = S.vanilla_numba((8,) * 4, A5, b5, complex(c5))
    out["cfg5_G8"] = G5s
    g5s = np.random.RandomState(1).standard_normal(G5s.shape) + 1j * np.random.RandomState(2).standard_normal(G5s.shape)
    dA, db, dc = S.vanilla_vjp_numba(G5s, complex(c5), g5s)
    out.update(cfg5_g8=g5s, cfg5_dA8=dA, cfg5_db8=db, cfg5_dc8=np.asarray(dc))
    G5 = S.vanilla_numba((40,) * 4, A5, b5, complex(c5))
    out["cfg5_G40_sha"] = np.array(sha(G5))
    out["cfg5_G40_sample"] = G5.ravel()[::4999].copy()
    # cotangent for the full-size VJP is regenerated from the seed in the tests
    g5 = np.random.RandomState(1).standard_normal(G5.shape) + 0j
    dA, db, dc = S.vanilla_vjp_numba(G5, complex(c5), g5)
    out.update(cfg5_dA40=dA, cfg5_db40=db, cfg5_dc40=np.asarray(dc))

    out.update(cfg4_A=A8, cfg4_b=b8, cfg4_c=c8)
    out["cfg4_G3"] = S.vanilla_numba((3,) * 8, A8, b8, complex(c8))

    # ---- random triples over odd shapes (H5: size-1 dims, ragged shapes) ---------------------------
    cases = {
        "r1": (1, (33,), 5), "r2": (2, (7, 5), 673), "r2b": (2, (40, 40), 7), "r3": (3, (4, 4, 4), 673),
        "r3b": (3, (1, 2, 3), 11), "r3c": (3, (3, 1, 5), 12), "r4": (4, (6, 5, 4, 3), 13),
        "r4b": (4, (1, 1, 1, 1), 14), "r4c": (4, (2, 9, 1, 7), 15), "r5": (5, (3, 2, 3, 2, 3), 16),
        "r6": (6, (2,) * 6, 17), "r6b": (6, (5, 2, 2, 5, 2, 2), 18), "r1b": (1, (1,), 19),
    }
    names = []
    for name, (n, shape, seed) in cases.items():
        A, b, c = random_triple(n, (), seed=seed)
        G = S.vanilla_numba(shape, A, b, complex(c))
        Gs = S.stable_numba(shape, A, b, complex(c))
        g = np.random.RandomState(seed + 1000).standard_normal(shape) + 1j * np.random.RandomState(seed + 2000).standard_normal(shape)
        dA, db, dc = S.vanilla_vjp_numba(G, complex(c), g)
        out.update({f"{name}_A": A, f"{name}_b": b, f"{name}_c": np.asarray(c), f"{name}_shape": np.array(shape),
                    f"{name}_G": G, f"{name}_Gs": Gs, f"{name}_g": g, f"{name}_dA": dA, f"{name}_db": db,
                    f"{name}_dc": np.asarray(dc)})
        names.append(name)
    out["random_cases"] = np.array(names)

    # ---- batched -------------------------------------------------------------------------------
    bcases = {"b2": (2, (5,), (7, 6), 673), "b3": (3, (2,), (1, 2, 3), 21), "b4": (4, (3,), (4, 3, 2, 3), 22),
              "b1": (1, (9,), (17,), 23)}
    bn = []
    for name, (n, batch, shape, seed) in bcases.items():
        A, b, c = random_triple(n, batch, seed=seed)
        G = S.vanilla_batch_numba(shape, A, b, c, False)
        Gs = S.vanilla_batch_numba(shape, A, b, c, True)
        g = np.random.RandomState(seed + 1).standard_normal(G.shape) + 1j * np.random.RandomState(seed + 2).standard_normal(G.shape)
        dA, db, dc = S.vanilla_batch_vjp_numba(G, c, g)
        out.update({f"{name}_A": A, f"{name}_b": b, f"{name}_c": c, f"{name}_shape": np.array(shape), f"{name}_G": G,
                    f"{name}_Gs": Gs, f"{name}_g": g, f"{name}_dA": dA, f"{name}_db": db, f"{name}_dc": dc})
        bn.append(name)
    out["batch_cases"] = np.array(bn)

    # ---- cfg3: 65,536 random 2-mode triples, cutoff 40 (inputs are regenerated from the seed) -----
    A3, b3, c3 = random_triple(2, (65536,), seed=673)
    G3 = S.vanilla_batch_numba((40, 40), A3, b3, c3, False)
    out["cfg3_in_sha"] = np.array(sha(np.concatenate([A3.ravel(), b3.ravel(), c3.ravel()])))
    out["cfg3_G_sha"] = np.array(sha(G3))
    out["cfg3_G_first4"] = G3[:4].copy()
    out["cfg3_G_sample"] = G3.ravel()[::1000003].copy()
    # sha per 4096-triple chunk so that sharded runs can be checked rank by rank
    out["cfg3_chunk_sha"] = np.array([sha(G3[i:i + 4096]) for i in range(0, 65536, 4096)])
    g3 = np.random.RandomState(1).standard_normal((64, 40, 40)) + 0j
    dA, db, dc = S.vanilla_batch_vjp_numba(G3[:64].copy(), c3[:64].copy(), g3)
    out.update(cfg3_dA64=dA, cfg3_db64=db, cfg3_dc64=dc)
    G3s = S.vanilla_batch_numba((40, 40), A3[:64].copy(), b3[:64].copy(), c3[:64].copy(), True)
    out["cfg3_Gs64_sha"] = np.array(sha(G3s))
    del G3

    # ---- binomial ------------------------------------------------------------------------------
    out.update(bin_A=Ak, bin_b=bk, bin_c=ck)
    for tag, (cut, max_l2, gc) in {"a": ((5, 5), 0.9999, 12), "b": ((10, 10), 0.9, 15), "c": ((10, 10), 0.5, 19),
                                    "d": ((6, 9), 2.0, 14)}.items():
        G, norm = S.binomial(cut, Ak, bk, complex(ck), max_l2, gc)
        out[f"bin_{tag}_G"] = G
        out[f"bin_{tag}_norm"] = np.asarray(norm)
        out[f"bin_{tag}_args"] = np.array([cut[0], cut[1], max_l2, gc], dtype=np.float64)
    A3d, b3d, c3d = random_triple(3, (), seed=31)
    G, norm = S.binomial((4, 3, 5), A3d, b3d, complex(c3d), 1e9, 10)
    out.update(bin3_A=A3d, bin3_b=b3d, bin3_c=np.asarray(c3d), bin3_G=G, bin3_norm=np.asarray(norm))

    # ---- reference test_vanilla_stable (tests/test_math/test_lattice/test_lattice_functions.py:137-149):
    # Dgate(4+4j) and Sgate(r=4, phi=2) at cutoff 1000 through the stable strategy
    from mrmustard.lab import Dgate  # noqa: PLC0415
    for tag, gate in {"dg": Dgate(0, 4 + 4j), "sg": Sgate(0, r=4.0, phi=2.0)}.items():
        At, bt, ct = (np.asarray(x, dtype=np.complex128) for x in gate.bargmann_triple())
        Gt = S.stable_numba((1000, 1000), At, bt, complex(ct))
        out.update({f"st_{tag}_A": At, f"st_{tag}_b": bt, f"st_{tag}_c": ct, f"st_{tag}_sha": np.array(sha(Gt)),
                    f"st_{tag}_sample": Gt.ravel()[::7919].copy(), f"st_{tag}_absmax": np.asarray(np.abs(Gt).max())})
    from mrmustard.math.lattice.strategies.displacement import displacement  # noqa: PLC0415
    out["st_dg_closed_form_sample"] = displacement((1000, 1000), 4.0 + 4.0j).ravel()[::7919].copy()

    path = os.path.join(HERE, "vanilla_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) / 1e6, "MB;", len(out), "arrays")


if __name__ == "__main__":
    main()
