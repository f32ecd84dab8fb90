"""Pins the autoshape oracle (oracle/autoshape.py) to the reference's golden vectors (CPU)."""
import os

import numpy as np

from conftest import GOLDEN
from oracle import autoshape as oa


def test_autoshape_goldens():
    g = np.load(os.path.join(GOLDEN, "autoshape_golden.npz"))
    for tag in g["cases"]:
        mp, mx, mn = g[f"{tag}_args"]
        got = oa.autoshape(g[f"{tag}_A"], g[f"{tag}_b"], g[f"{tag}_c"], float(mp), int(mx), int(mn))
        assert got.dtype == np.int64 and np.array_equal(got, g[f"{tag}_shape"]), tag
