"""Golden vectors for the gate-specific Fock strategies from the UNMODIFIED reference.

    python tests/golden/gen_golden_gates.py

Reference functions (file:line under /root/reference/mrmustard/math/lattice/strategies/):
  displacement.py:24 displacement, :85 grad_displacement, :117 jacobian_displacement, :68 laguerre
  squeezer.py:29 squeezer, :69 squeezer_vjp, :127 squeezed, :150 squeezed_vjp
  beamsplitter.py:37 beamsplitter, :94 stable_beamsplitter, :175 beamsplitter_vjp
Output: tests/golden/gates_golden.npz
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import refimport  # noqa: E402


def main():
    S = refimport.strategies()
    out = {}
    rng = np.random.RandomState(77)

    def cot(shape):
        return rng.standard_normal(shape) + 1j * rng.standard_normal(shape)

    # ---- displacement ---------------------------------------------------------------------------------------
    names = []
    for k, (cut, alpha) in enumerate([((10, 10), 0.3 + 0.2j), ((7, 12), 1.5 - 0.7j), ((12, 7), -0.4 + 1.1j), ((1, 5), 0.5j),
                                      ((40, 40), 2.0 + 0.0j), ((150, 150), 1.2 + 0.9j), ((6, 6), 1e-9 + 0j)]):
        D = S.displacement(cut, complex(alpha))
        tag = f"disp{k}"
        big = D.size > 5000           # large cases: strided samples (stride 7) instead of the full arrays
        pick = (lambda a: a.ravel()[::7].copy()) if big else (lambda a: a)
        out.update({f"{tag}_cut": np.array(cut), f"{tag}_alpha": np.array(alpha), f"{tag}_D": pick(D), f"{tag}_big": np.array(big)})
        if cut[0] == cut[1]:
            ja, jac = S.jacobian_displacement(D, complex(alpha))
            gr, gphi = S.grad_displacement(D, float(abs(alpha)), float(np.angle(alpha)))
            out.update({f"{tag}_ja": pick(ja), f"{tag}_jac": pick(jac), f"{tag}_gr": pick(gr), f"{tag}_gphi": pick(gphi)})
        names.append(tag)
    out["disp_cases"] = np.array(names)

    # ---- squeezer / squeezed -----------------------------------------------------------------------------------
    names = []
    for k, (shape, r, th) in enumerate([((10, 10), 0.4, 0.7), ((8, 13), 1.0, -1.3), ((13, 8), 0.05, 2.0), ((1, 1), 0.3, 0.1),
                                        ((2, 9), 0.8, 0.0), ((60, 60), 0.6, 0.9), ((200, 200), 0.3, -0.4)]):
        G = S.squeezer(shape, float(r), float(th))
        big = G.size > 5000
        pick = (lambda a: a.ravel()[::7].copy()) if big else (lambda a: a)
        g = np.random.RandomState(2000 + k).standard_normal(shape) + 1j * np.random.RandomState(3000 + k).standard_normal(shape)
        dr, dphi = S.squeezer_vjp(G, g, float(r), float(th))
        tag = f"sq{k}"
        out.update({f"{tag}_shape": np.array(shape), f"{tag}_r": np.array(r), f"{tag}_theta": np.array(th), f"{tag}_G": pick(G),
                    f"{tag}_big": np.array(big), f"{tag}_gseed": np.array(2000 + k), f"{tag}_dr": np.array(dr), f"{tag}_dphi": np.array(dphi)})
        names.append(tag)
    out["sq_cases"] = np.array(names)
    names = []
    for k, (cut, r, th) in enumerate([(30, 0.5, 0.3), (1, 0.2, 0.0), (7, 1.2, -2.0), (500, 0.8, 1.0)]):
        G = S.squeezed(int(cut), float(r), float(th))
        g = cot((cut,))
        dr, dphi = S.squeezed_vjp(G, g, float(r), float(th))
        tag = f"sqz{k}"
        out.update({f"{tag}_cut": np.array(cut), f"{tag}_r": np.array(r), f"{tag}_theta": np.array(th), f"{tag}_G": G, f"{tag}_g": g,
                    f"{tag}_dr": np.array(dr), f"{tag}_dphi": np.array(dphi)})
        names.append(tag)
    out["sqz_cases"] = np.array(names)

    # ---- beamsplitter ------------------------------------------------------------------------------------------
    names = []
    for k, (shape, th, ph) in enumerate([((5, 5, 5, 5), 0.5, 0.2), ((4, 6, 5, 7), 1.1, -0.6), ((7, 3, 2, 6), 0.3, 2.5), ((1, 1, 1, 1), 0.4, 0.1),
                                         ((3, 1, 4, 2), 0.9, 0.0), ((14, 14, 14, 14), 0.7, 1.3), ((24, 20, 22, 25), 0.25, -1.0)]):
        G = S.beamsplitter(shape, float(th), float(ph))
        Gs = S.stable_beamsplitter(shape, float(th), float(ph))
        g = cot(shape)
        dth, dph = S.beamsplitter_vjp(G, g, float(th), float(ph))
        tag = f"bs{k}"
        big = int(np.prod(shape)) > 50000
        out.update({f"{tag}_shape": np.array(shape), f"{tag}_theta": np.array(th), f"{tag}_phi": np.array(ph),
                    f"{tag}_dtheta": np.array(dth), f"{tag}_dphi": np.array(dph), f"{tag}_gseed": np.array(1000 + k)})
        if big:      # store a strided sample + the cotangent seed instead of ~5 MB arrays
            g = np.random.RandomState(1000 + k).standard_normal(shape) + 0j
            dth, dph = S.beamsplitter_vjp(G, g, float(th), float(ph))
            out.update({f"{tag}_dtheta": np.array(dth), f"{tag}_dphi": np.array(dph)})
            out.update({f"{tag}_Gsample": G.ravel()[::101].copy(), f"{tag}_Gssample": Gs.ravel()[::101].copy(),
                        f"{tag}_Gabs2": np.array(np.sum(np.abs(G) ** 2)), f"{tag}_Gsabs2": np.array(np.sum(np.abs(Gs) ** 2))})
        else:
            out.update({f"{tag}_G": G, f"{tag}_Gs": Gs, f"{tag}_g": g})
        names.append(tag)
    out["bs_cases"] = np.array(names)
    path = os.path.join(HERE, "gates_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) / 1e6, "MB")


if __name__ == "__main__":
    main()
