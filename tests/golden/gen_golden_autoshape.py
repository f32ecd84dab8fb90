"""Golden vectors for autoshape_numba from the UNMODIFIED reference.

    python tests/golden/gen_golden_autoshape.py

Reference: mrmustard/math/lattice/autoshape.py:24-154 through the same call State.auto_shape makes (lab/states/base.py:416-431):
the (A, b, c) of `ansatz.conj & ansatz` for kets, of the ansatz itself for density matrices.
Output: tests/golden/autoshape_golden.npz
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
    from mrmustard.lab import DM, Coherent, Dgate, Ket, SqueezedVacuum, Vacuum
    from mrmustard.math.lattice.autoshape import autoshape_numba

    out, names = {}, []

    def add(tag, state, max_prob, max_shape, min_shape):
        if not state.wires.ket or not state.wires.bra:
            ansatz = state.ansatz.conj & state.ansatz
        else:
            ansatz = state.ansatz
        A, b, c = (np.asarray(math.asnumpy(x), dtype=np.complex128) for x in ansatz.triple)
        shape = np.asarray(autoshape_numba(A, b, c, max_prob, max_shape, min_shape))
        out.update({f"{tag}_A": A, f"{tag}_b": b, f"{tag}_c": c, f"{tag}_args": np.array([max_prob, max_shape, min_shape], dtype=np.float64),
                    f"{tag}_shape": shape})
        names.append(tag)
        print(tag, shape)

    add("vac", Vacuum((0, 1)), 0.999, 50, 1)
    add("coh", Coherent(0, alpha=1.0 + 0.5j), 0.999, 50, 1)
    add("sqv", SqueezedVacuum(0, r=0.8, phi=0.4), 0.99999, 100, 1)
    add("coh2", Coherent(0, alpha=2.5) >> Dgate(0, 0.3j), 0.999, 20, 3)       # clipped by max_shape
    for k, (modes, seed, prob) in enumerate([((0, 1), 1, 0.999), ((0, 1, 2), 2, 0.9999), ((0, 1, 2, 3, 4), 3, 0.999), ((0, 1, 2, 3, 4, 5, 6, 7), 4, 0.99)]):
        with settings(SEED=seed):
            add(f"ket{k}", Ket.random(modes), prob, 50, 1)
    for k, (modes, seed, prob) in enumerate([((0,), 5, 0.999), ((0, 1), 6, 0.99999), ((0, 1, 2, 3), 7, 0.999)]):
        with settings(SEED=seed):
            add(f"dm{k}", DM.random(modes) >> Dgate(modes[0], 0.4), prob, 60, 2)
    out["cases"] = np.array(names)
    path = os.path.join(HERE, "autoshape_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) / 1e3, "kB")


if __name__ == "__main__":
    main()
