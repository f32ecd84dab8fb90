"""Pins the gate-strategy oracle (oracle/gates.py) to the reference's golden vectors (CPU)."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, assert_parity
from oracle import gates as og


@pytest.fixture(scope="module")
def gg():
    return np.load(os.path.join(GOLDEN, "gates_golden.npz"))


def _pick(a, big):
    return a.ravel()[::7] if big else a


def test_displacement(gg):
    for tag in gg["disp_cases"]:
        cut = tuple(int(x) for x in gg[f"{tag}_cut"]); alpha = complex(gg[f"{tag}_alpha"]); big = bool(gg[f"{tag}_big"])
        if big and cut[0] > 60:
            continue    # (150, 150) is covered on the GPU
        D = og.displacement(cut, alpha)
        assert D.shape == cut
        assert_parity(_pick(D, big), gg[f"{tag}_D"], tag)
        if cut[0] == cut[1]:
            ja, jac = og.jacobian_displacement(D, alpha)
            gr, gphi = og.grad_displacement(D, abs(alpha), float(np.angle(alpha)))
            for got, nm in ((ja, "ja"), (jac, "jac"), (gr, "gr"), (gphi, "gphi")):
                assert_parity(_pick(got, big), gg[f"{tag}_{nm}"], f"{tag} {nm}")


def test_squeezer_and_vjp(gg):
    for tag in gg["sq_cases"]:
        shape = tuple(int(x) for x in gg[f"{tag}_shape"]); r = float(gg[f"{tag}_r"]); th = float(gg[f"{tag}_theta"]); big = bool(gg[f"{tag}_big"])
        if big:
            continue    # (200, 200): GPU only
        G = og.squeezer(shape, r, th)
        assert np.array_equal(G + 0.0, gg[f"{tag}_G"] + 0.0), tag      # bit-identical to the numba strategy
        k = int(gg[f"{tag}_gseed"])
        g = np.random.RandomState(k).standard_normal(shape) + 1j * np.random.RandomState(k + 1000).standard_normal(shape)
        dr, dphi = og.squeezer_vjp(G, g, r, th)
        assert_parity(np.float64(dr), np.float64(gg[f"{tag}_dr"]), tag + " dr")
        assert_parity(np.float64(dphi), np.float64(gg[f"{tag}_dphi"]), tag + " dphi")


def test_squeezed_and_vjp(gg):
    for tag in gg["sqz_cases"]:
        cut = int(gg[f"{tag}_cut"]); r = float(gg[f"{tag}_r"]); th = float(gg[f"{tag}_theta"])
        G = og.squeezed(cut, r, th)
        assert np.array_equal(G + 0.0, gg[f"{tag}_G"] + 0.0), tag
        dr, dphi = og.squeezed_vjp(G, gg[f"{tag}_g"], r, th)
        assert_parity(np.float64(dr), np.float64(gg[f"{tag}_dr"]), tag + " dr")
        assert_parity(np.float64(dphi), np.float64(gg[f"{tag}_dphi"]), tag + " dphi")


def test_beamsplitter_and_vjp(gg):
    for tag in gg["bs_cases"]:
        shape = tuple(int(x) for x in gg[f"{tag}_shape"]); th = float(gg[f"{tag}_theta"]); ph = float(gg[f"{tag}_phi"])
        if f"{tag}_G" not in gg.files:
            continue    # large cases: GPU only
        G = og.beamsplitter(shape, th, ph)
        assert np.array_equal(G + 0.0, gg[f"{tag}_G"] + 0.0), tag
        if int(np.prod(shape)) <= 5000:
            assert np.array_equal(og.stable_beamsplitter(shape, th, ph) + 0.0, gg[f"{tag}_Gs"] + 0.0), tag + " stable"
            dth, dph = og.beamsplitter_vjp(G, gg[f"{tag}_g"], th, ph)
            assert_parity(np.float64(dth), np.float64(gg[f"{tag}_dtheta"]), tag + " dtheta")
            assert_parity(np.float64(dph), np.float64(gg[f"{tag}_dphi"]), tag + " dphi")
