"""Generate golden vectors for the compactFock diagonal / 1-leftover-mode paths from the UNMODIFIED reference.

    python tests/golden/gen_golden_diagonal.py

Reference entry points used: math.hermite_renormalized_diagonal (backend_numpy.py:423-432 ->
compactFock/inputValidation.py:61-79), compactFock hermite_multidimensional_1leftoverMode
(inputValidation.py:103-122), strategies.fast_diagonal (fast_diagonal.py:32-77), and
math.hermite_renormalized_1leftoverMode (backend_numpy.py:434-446) on DM.random / Ket-derived triples.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import refimport  # noqa: E402


def main():
    refimport.install_shims(with_lab=True)
    from mrmustard import math, settings
    from mrmustard.lab import DM, Dgate
    from mrmustard.math.lattice import strategies as S
    from mrmustard.math.lattice.strategies.compactFock.inputValidation import (
        hermite_multidimensional_1leftoverMode, hermite_multidimensional_diagonal)

    out = {}

    def triple(modes, seed, disp=True):
        with settings(SEED=seed):
            st = DM.random(modes)
            if disp:
                for m in modes:
                    st = st >> Dgate(m, 0.1 * (m + 1))
            return tuple(np.asarray(x, dtype=np.complex128) for x in st.bargmann_triple())

    cases = {"d1": ([0], 1, (9,)), "d2": ([0, 1], 2, (6, 7)), "d3": ([0, 1, 2], 3, (5, 5, 5)),
             "d3b": ([0, 1, 2], 4, (18, 19, 20)), "d4": ([0, 1, 2, 3], 5, (4, 3, 5, 4)), "d3c": ([0, 1, 2], 6, (1, 4, 3)),
             "d2b": ([0, 1], 7, (1, 1))}
    names = []
    for name, (modes, seed, cutoffs) in cases.items():
        A, b, c = triple(modes, seed)
        G = math.hermite_renormalized_diagonal(A, b, c, cutoffs=cutoffs)
        out.update({f"{name}_A": A, f"{name}_b": b, f"{name}_c": np.asarray(c), f"{name}_cut": np.array(cutoffs), f"{name}_G": np.asarray(G)})
        # the same through the already-reordered entry (reorderedAB=False on interleaved inputs)
        A2, b2 = math.backend.reorder_AB_bargmann(A, b)
        G2 = hermite_multidimensional_diagonal(np.asarray(A2), np.asarray(b2), c, cutoffs)[0]
        assert np.array_equal(G, G2)
        names.append(name)
    out["diag_cases"] = np.array(names)
    # forward-mode Jacobians (compactFock/inputValidation.py:82-100 -> diagonal_grad.py)
    from mrmustard.math.lattice.strategies.compactFock.inputValidation import grad_hermite_multidimensional_diagonal
    gnames = []
    for name in ("d1", "d2", "d3", "d3c", "d4"):
        A, b, c = out[f"{name}_A"], out[f"{name}_b"], complex(out[f"{name}_c"])
        cutoffs = tuple(int(x) for x in out[f"{name}_cut"])
        A2, b2 = (np.asarray(x) for x in math.backend.reorder_AB_bargmann(A, b))
        arrs = hermite_multidimensional_diagonal(A2, b2, c, cutoffs)
        dG0, dA, dB = grad_hermite_multidimensional_diagonal(A2, b2, c, *arrs)
        out.update({f"{name}_dG0": dG0, f"{name}_dA": dA, f"{name}_dB": dB})
        gnames.append(name)
    out["grad_cases"] = np.array(gnames)
    # b-batched diagonal (batch on the LAST axis of B and of the output, diagonal_amps.py:223-233)
    A, b, c = triple([0, 1, 2], 8)
    bb = np.stack([b, 0.5 * b, b + 0.05], axis=1)
    Gb = math.hermite_renormalized_diagonal(A, bb, c, cutoffs=(6, 5, 7))
    out.update(db_A=A, db_b=bb, db_c=np.asarray(c), db_cut=np.array((6, 5, 7)), db_G=np.asarray(Gb))

    lcases = {"l2": ([0, 1], 11, 3, (4,)), "l3": ([0, 1, 2], 12, 3, (1, 2)), "l3b": ([0, 1, 2], 13, 11, (11, 11)),
              "l4": ([0, 1, 2, 3], 14, 4, (3, 2, 3)), "l3c": ([0, 1, 2], 15, 2, (2, 3))}
    lnames = []
    for name, (modes, seed, oc, pnr) in lcases.items():
        A, b, c = triple(modes, seed)
        Gl = math.hermite_renormalized_1leftoverMode(A, b, c, output_cutoff=oc, pnr_cutoffs=pnr)       # numpy backend -> fast_diagonal
        Gf = S.fast_diagonal(A, b, c, oc, pnr, False)
        A2, b2 = math.backend.reorder_AB_bargmann(A, b)
        Gc = hermite_multidimensional_1leftoverMode(np.asarray(A2), np.asarray(b2), c, (oc + 1,) + tuple(p + 1 for p in pnr))[0]
        assert np.allclose(Gl, Gc)
        out.update({f"{name}_A": A, f"{name}_b": b, f"{name}_c": np.asarray(c), f"{name}_oc": np.array(oc), f"{name}_pnr": np.array(pnr),
                    f"{name}_G": np.asarray(Gl), f"{name}_Gfast": np.asarray(Gf), f"{name}_Gcompact": np.asarray(Gc)})
        lnames.append(name)
    out["leftover_cases"] = np.array(lnames)
    # The reference's fast_diagonal quirk: its weight loop `range(1, 2*oc + 2*sum(pnr) - L)` (fast_diagonal.py:68) ends before the top
    # weight 2*sum(pnr) whenever 2*output_cutoff - L < 1, leaving the highest-weight conditional density matrices zero; the compactFock
    # strategy (and the diagonal of the vanilla lattice) hold the true values.  One such case, both reference outputs stored.
    A, b, c = triple([0, 1, 2], 17)
    oc, pnr = 1, (2, 2)
    Gf = S.fast_diagonal(A, b, c, oc, pnr, False)
    A2, b2 = math.backend.reorder_AB_bargmann(A, b)
    Gc = hermite_multidimensional_1leftoverMode(np.asarray(A2), np.asarray(b2), c, (oc + 1,) + tuple(p + 1 for p in pnr))[0]
    out.update(lq_A=A, lq_b=b, lq_c=np.asarray(c), lq_oc=np.array(oc), lq_pnr=np.array(pnr), lq_Gfast=np.asarray(Gf), lq_Gcompact=np.asarray(Gc))
    # (The Jacobians of the one-leftover-mode path are generated by tests/golden/gen_golden_leftover_grad.py: the reference's
    #  grad_hermite_multidimensional_1leftoverMode does not compile under the numba 0.65 of this image, so that generator runs the
    #  unmodified source with the JIT disabled.)
    path = os.path.join(HERE, "diagonal_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) / 1e6, "MB")


if __name__ == "__main__":
    main()
