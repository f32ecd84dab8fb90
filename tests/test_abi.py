"""CPU checks of the C-ABI boundary: the library loads, exports every symbol include/mmhermite.h declares,
validates arguments, and fails loudly (no CPU fallback) when no GPU is present."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    txt = open(os.path.join(ROOT, "include", "mmhermite.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(mmh_[a-z_0-9]+)\s*\(", txt)))


def test_header_symbols_are_exported():
    from mrmustard_b200 import _lib
    declared = _declared_symbols()
    assert declared, "no declarations parsed from include/mmhermite.h"
    for name in declared:
        assert hasattr(_lib.lib, name), f"{name} declared in mmhermite.h but not exported"
    assert sorted(_lib.SYMBOLS) == declared


def test_version_and_error_strings():
    from mrmustard_b200 import _lib
    assert _lib.lib.mmh_version() >= 100
    assert b"ndim" in _lib.lib.mmh_error_string(-1)
    assert b"CPU fallback" in _lib.lib.mmh_error_string(-6)


def test_argument_validation_needs_no_gpu():
    from mrmustard_b200 import _lib
    sh = _lib.shape_array((3, 0))
    assert _lib.lib.mmh_forward(2, sh, None, None, None, None, 0, None) == -2      # bad shape
    assert _lib.lib.mmh_forward(0, sh, None, None, None, None, 0, None) == -1      # bad ndim
    assert _lib.lib.mmh_forward(33, sh, None, None, None, None, 0, None) == -1
    sh = _lib.shape_array((3, 3))
    assert _lib.lib.mmh_forward(2, sh, None, None, None, None, 0, None) == -3      # null pointers
    assert _lib.lib.mmh_forward_batched(-1, 2, sh, None, None, None, None, 0, None) == -4
    assert _lib.lib.mmh_forward_batched(0, 2, sh, None, None, None, None, 0, None) == 0   # empty batch is a no-op


def test_no_cpu_fallback():
    """Without a CUDA device the product path raises; it never computes on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from mrmustard_b200 import strategies
    with pytest.raises(RuntimeError):
        strategies.vanilla_numba((3, 3), np.eye(2) * 0.1, np.ones(2), 1.0)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "mrmustard_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src and "libmmoracle" not in src, f
