"""Golden vectors for the lattice + derived-variable contraction, generated from the UNMODIFIED
reference: CircuitComponent.fock_array on PolyExpAnsatz objects with num_derived_vars > 0
(lab/circuit_components.py:516-530), unbatched and batched.

    python tests/golden/gen_golden_contract.py
"""
from __future__ import annotations

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import refimport  # noqa: E402


def triple(rng, n, batch=()):
    A = rng.random((*batch, n, n)) + 1j * rng.random((*batch, n, n))
    A = A + np.swapaxes(A, -1, -2)
    A /= np.abs(np.linalg.eigvals(A)).max() + 0.2
    b = rng.random((*batch, n)) + 1j * rng.random((*batch, n))
    return A, b


def main():
    refimport.install_shims(with_lab=True)
    from mrmustard.lab import CircuitComponent  # noqa: PLC0415
    from mrmustard.physics.ansatz import PolyExpAnsatz  # noqa: PLC0415
    from mrmustard.physics.wires import Wires  # noqa: PLC0415

    rng = np.random.RandomState(17)
    out, names = {}, []
    # (name, core shape, derived shape, batch)
    cases = [("u1", (7,), (3,), ()), ("u2", (5, 6), (3, 4), ()), ("u3", (4, 3, 5), (2,), ()), ("u4", (6, 5), (40,), ()),
             ("b2", (5, 6), (3, 4), (3,)), ("b1", (9,), (2, 2), (2, 2))]
    for name, core, der, batch in cases:
        n = len(core) + len(der)
        A, b = triple(rng, n, batch)
        c = rng.random((*batch, *der)) + 1j * rng.random((*batch, *der))
        comp = CircuitComponent(PolyExpAnsatz(A, b, c), Wires(modes_out_ket=set(range(len(core)))))
        F = np.asarray(comp.fock_array(core))
        assert F.shape == (*batch, *core)
        out.update({f"{name}_A": A, f"{name}_b": b, f"{name}_c": c, f"{name}_F": F,
                    f"{name}_core": np.array(core), f"{name}_der": np.array(der)})
        names.append(name)
    out["cases"] = np.array(names)
    path = os.path.join(HERE, "contract_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
