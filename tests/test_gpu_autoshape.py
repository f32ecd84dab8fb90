"""GPU parity test of autoshape against the reference's goldens and, through the drop-in, against the
live reference's State.auto_shape."""
import os

import numpy as np
import pytest

from conftest import GOLDEN

pytestmark = pytest.mark.gpu


def test_autoshape_goldens():
    from mrmustard_b200 import strategies as S
    g = np.load(os.path.join(GOLDEN, "autoshape_golden.npz"))
    for tag in g["cases"]:
        mp, mx, mn = g[f"{tag}_args"]
        got = S.autoshape_numba(g[f"{tag}_A"], g[f"{tag}_b"], g[f"{tag}_c"], float(mp), int(mx), int(mn))
        assert got.dtype == np.int64 and np.array_equal(got, g[f"{tag}_shape"]), (tag, got, g[f"{tag}_shape"])


def test_autoshape_vs_oracle_unseen():
    from conftest import random_triple
    from mrmustard_b200 import strategies as S
    from oracle import autoshape as oa
    rng = np.random.RandomState(3)
    for M in (1, 2, 3, 6):
        # a valid Gaussian density-matrix triple: |psi><psi| of a random ket triple (A_dm = conj(A) (+) A)
        A, b, c = random_triple(M, (), seed=20 + M)
        A = A * 0.5
        Adm = np.zeros((2 * M, 2 * M), complex); Adm[:M, :M] = np.conj(A); Adm[M:, M:] = A
        bdm = np.concatenate([np.conj(b), b]) * 0.3
        cdm = abs(complex(c)) ** 2 * 0.05
        for prob in (0.01, 0.04):
            got = S.autoshape_numba(Adm, bdm, cdm, prob, 40, 1)
            assert np.array_equal(got, oa.autoshape(Adm, bdm, cdm, prob, 40, 1)), (M, prob, got)


def test_auto_shape_through_the_dropin():
    from oracle import refimport
    if not refimport.available():
        pytest.skip("no reference install in this tree")
    refimport.install_shims(with_lab=True)
    import mrmustard as mm
    from mrmustard.lab import Ket
    from mrmustard_b200 import _lib, dropin
    with mm.settings(SEED=11):
        k = Ket.random((0, 1, 2))
    want = tuple(k.auto_shape())
    dropin.install()
    try:
        n0 = _lib.launch_count()
        got = tuple(k.auto_shape())
        assert _lib.launch_count() > n0
    finally:
        dropin.uninstall()
    assert got == want
