"""Golden vectors for the Jacobians of the one-leftover-mode amplitudes, from the UNMODIFIED reference run in pure Python.

    python tests/golden/gen_golden_leftover_grad.py          (re-executes itself with NUMBA_DISABLE_JIT=1)

`grad_hermite_multidimensional_1leftoverMode` (compactFock/inputValidation.py:122-142 -> singleLeftoverMode_grad.py:560-724)
does not compile under this image's numba 0.65 (interpreter assertion in peep_hole_list_to_tuple).  With the JIT disabled the
same source runs as plain Python; the only thing missing then is numba's intrinsic `tuple_setitem`
(numba.cpython.unsafe.tuple), which has no Python body.  This GENERATOR rebinds that one name in the four compactFock modules
to its documented meaning (a copy of the tuple with entry i replaced); the reference files themselves are not touched.
Output: tests/golden/leftover_grad_golden.npz.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))

if os.environ.get("NUMBA_DISABLE_JIT") != "1":
    env = dict(os.environ, NUMBA_DISABLE_JIT="1")
    sys.exit(subprocess.call([sys.executable, os.path.abspath(__file__)], env=env))

import numpy as np  # noqa: E402

sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import refimport  # noqa: E402


def main():
    refimport.install_shims(with_lab=True)
    import importlib
    pkg = "mrmustard.math.lattice.strategies.compactFock."
    for name in ("diagonal_amps", "diagonal_grad", "singleLeftoverMode_amps", "singleLeftoverMode_grad"):
        mod = importlib.import_module(pkg + name)
        mod.tuple_setitem = lambda t, i, v: tuple(t[:i]) + (v,) + tuple(t[i + 1:])
    from mrmustard import math, settings
    from mrmustard.lab import DM, Dgate
    from mrmustard.math.lattice.strategies.compactFock.inputValidation import (
        grad_hermite_multidimensional_1leftoverMode, grad_hermite_multidimensional_diagonal,
        hermite_multidimensional_1leftoverMode, hermite_multidimensional_diagonal)

    def triple(modes, seed):
        with settings(SEED=seed):
            st = DM.random(modes)
            for m in modes:
                st = st >> Dgate(m, 0.1 * (m + 1))
            return tuple(np.asarray(x, dtype=np.complex128) for x in st.bargmann_triple())

    out = {}
    # (modes, seed, cutoffs = (c0, tail...))
    cases = {"g2": ([0, 1], 11, (4, 5)), "g3": ([0, 1, 2], 12, (4, 2, 3)), "g3b": ([0, 1, 2], 13, (3, 4, 4)),
             "g3c": ([0, 1, 2], 15, (1, 3, 2)), "g4": ([0, 1, 2, 3], 14, (3, 2, 3, 2)), "g2b": ([0, 1], 16, (6, 1))}
    names = []
    for name, (modes, seed, cut) in cases.items():
        A, b, c = triple(modes, seed)
        A2, b2 = (np.ascontiguousarray(np.asarray(x)) for x in math.backend.reorder_AB_bargmann(A, b))
        arrs = hermite_multidimensional_1leftoverMode(A2, b2, c, cut)
        dG0, dA, dB = grad_hermite_multidimensional_1leftoverMode(A2, b2, c, *arrs)
        out.update({f"{name}_A": A2, f"{name}_b": b2, f"{name}_c": np.asarray(c), f"{name}_cut": np.array(cut),
                    f"{name}_G": np.asarray(arrs[0]), f"{name}_dG0": np.asarray(dG0), f"{name}_dA": np.asarray(dA),
                    f"{name}_dB": np.asarray(dB)})
        names.append(name)
        print(name, cut, np.asarray(arrs[0]).shape, np.asarray(dA).shape, np.asarray(dB).shape)
    out["cases"] = np.array(names)
    # cross-check of the pure-Python route itself: the diagonal Jacobians it produces equal the numba-compiled ones that
    # tests/golden/diagonal_golden.npz already holds (generated with the JIT enabled)
    gd = np.load(os.path.join(HERE, "diagonal_golden.npz"))
    for name in ("d2", "d3"):
        A, b, c = gd[f"{name}_A"], gd[f"{name}_b"], complex(gd[f"{name}_c"])
        cut = tuple(int(x) for x in gd[f"{name}_cut"])
        A2, b2 = (np.asarray(x) for x in math.backend.reorder_AB_bargmann(A, b))
        arrs = hermite_multidimensional_diagonal(A2, b2, c, cut)
        _, dA, dB = grad_hermite_multidimensional_diagonal(A2, b2, c, *arrs)
        assert np.allclose(dA, gd[f"{name}_dA"], rtol=1e-12, atol=1e-15) and np.allclose(dB, gd[f"{name}_dB"], rtol=1e-12, atol=1e-15)
    path = os.path.join(HERE, "leftover_grad_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) / 1e3, "kB")


if __name__ == "__main__":
    main()
