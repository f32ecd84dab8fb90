"""Import the UNMODIFIED reference (XanaduAI/MrMustard at /root/reference) in the build container.

TEST INFRASTRUCTURE ONLY.  Used by tests/golden/gen_golden.py to produce the committed golden
vectors and by optional `-m "not gpu"` tests that cross-check the C oracle against the live numba
strategies when /root/reference is present.  Nothing on the product path imports this module, and it
is never used on the GPU box (the reference tree does not exist there).

The shims follow SURVEY.md Appendix B: the reference has no dist metadata, imports opt_einsum (absent
here; only `contract` is used, backend_manager.py:548) and UI packages (ipywidgets/IPython/plotly).
"""
from __future__ import annotations

import importlib.metadata as _md
import os
import sys
import types

_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# baseline/_ref: `pip install --no-index --no-deps --target baseline/_ref /root/reference` (git-ignored, travels to the GPU
# box with the gpurun snapshot) -- the unmodified reference package, used by `bench.py --impl reference`, by bench.py's
# cpu_baseline leg and by tests/test_dropin_gpu.py where /root/reference itself does not exist.
STAGED_ROOT = os.path.join(_REPO, "baseline", "_ref")


def _pick_root() -> str:
    for cand in (os.environ.get("MMH_REFERENCE_ROOT"), "/root/reference", STAGED_ROOT):
        if cand and os.path.isdir(os.path.join(cand, "mrmustard")):
            return cand
    return "/root/reference"


REFERENCE_ROOT = _pick_root()


def stage(force: bool = False) -> bool:
    """Install the unmodified reference into baseline/_ref (build container only; needs /root/reference)."""
    import subprocess
    if os.path.isdir(os.path.join(STAGED_ROOT, "mrmustard")) and not force:
        return True
    if not os.path.isdir("/root/reference/mrmustard"):
        return False
    cmd = [sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--no-deps", "--find-links",
           "/opt/wheelhouse", "--target", STAGED_ROOT, "--upgrade", "/root/reference"]
    return subprocess.call(cmd, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL) == 0


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "mrmustard"))


class _Any(types.ModuleType):
    def __getattr__(self, n):
        if n.startswith("__"):
            raise AttributeError(n)
        m = _Any(self.__name__ + "." + n)
        setattr(self, n, m)
        return m

    def __call__(self, *a, **k):
        return None


_done = False


def install_shims(with_lab: bool = False) -> None:
    global _done
    import numpy as np

    if not available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    if not _done:
        os.environ.setdefault("NUMBA_CACHE_DIR", "/tmp/mmh_numba_cache")
        os.environ.setdefault("NUMBA_NUM_THREADS", str(os.cpu_count() or 1))
        _v = _md.version
        _md.version = lambda n: "1.0.0a1" if n == "mrmustard" else _v(n)
        if "opt_einsum" not in sys.modules:
            oe = types.ModuleType("opt_einsum")
            oe.contract = lambda s, *t, **k: np.einsum(s, *t, optimize=k.get("optimize", False))
            sys.modules["opt_einsum"] = oe
        if REFERENCE_ROOT not in sys.path:
            sys.path.insert(0, REFERENCE_ROOT)
        _done = True
    if with_lab and "ipywidgets" not in sys.modules:
        for n in ["ipywidgets", "IPython", "IPython.display", "IPython.terminal",
                  "IPython.terminal.interactiveshell", "plotly", "plotly.graph_objects",
                  "plotly.graph_objs", "plotly.subplots"]:
            sys.modules[n] = _Any(n)
        sys.modules["IPython.terminal.interactiveshell"].TerminalInteractiveShell = type(
            "TerminalInteractiveShell", (), {})
        sys.modules["IPython"].get_ipython = lambda: None


def strategies():
    """The reference's `mrmustard.math.lattice.strategies` module (numba njit functions)."""
    install_shims()
    from mrmustard.math.lattice import strategies as s  # noqa: PLC0415
    return s


def math():
    install_shims()
    from mrmustard import math as m  # noqa: PLC0415
    return m
