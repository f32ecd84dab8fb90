"""ctypes binding of libmmhermite.so (the C ABI in include/mmhermite.h).

There is no CPU fallback: importing this module without the built CUDA library raises ImportError, and
every compute call on a machine without an sm_100 GPU raises RuntimeError.
"""
from __future__ import annotations

import ctypes
import os
import threading

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(_HERE, "csrc", "libmmhermite.so")

MMH_OK = 0
_EXC = {
    -1: ValueError, -2: ValueError, -3: ValueError, -4: ValueError,
    -5: NotImplementedError, -6: RuntimeError, -7: MemoryError, -8: RuntimeError,
}

# every symbol include/mmhermite.h declares (tests/test_abi.py checks the header against this list)
SYMBOLS = [
    "mmh_version", "mmh_error_string", "mmh_device_count", "mmh_set_device", "mmh_device_synchronize",
    "mmh_launch_count", "mmh_host_alloc", "mmh_host_free",
    "mmh_forward", "mmh_forward_host", "mmh_forward_batched", "mmh_forward_batched_host", "mmh_forward_panel_range",
    "mmh_vjp", "mmh_vjp_host", "mmh_vjp_batched", "mmh_vjp_batched_host",
    "mmh_binomial", "mmh_binomial_host",
    "mmh_diagonal", "mmh_diagonal_host", "mmh_1leftover", "mmh_1leftover_host",
    "mmh_diagonal_grad", "mmh_diagonal_grad_host", "mmh_1leftover_grad", "mmh_1leftover_grad_host",
    "mmh_forward_contract", "mmh_forward_contract_host", "mmh_debug_timeline", "mmh_debug_plan",
    "mmh_squeezer", "mmh_squeezed", "mmh_beamsplitter", "mmh_displacement", "mmh_gate_host",
    "mmh_displacement_jacobian", "mmh_displacement_grad", "mmh_displacement_derivs_host", "mmh_gate_vjp", "mmh_gate_vjp_host",
    "mmh_autoshape", "mmh_autoshape_host", "mmh_fock_contract", "mmh_fock_contract_host", "mmh_fock_reduce", "mmh_overlap",
]


def _load() -> ctypes.CDLL:
    from . import build as _build
    if not os.path.exists(SO_PATH) or (_build.is_stale() and _build.have_nvcc()):
        try:  # (re)build in-tree if a toolchain is present; never fall back to a CPU implementation
            _build.build()
        except Exception as e:  # pragma: no cover
            if os.path.exists(SO_PATH):
                raise
            raise ImportError(
                f"mrmustard_b200: CUDA library {SO_PATH} is missing and could not be built ({e}). "
                "Run `python -m mrmustard_b200.build`; there is no CPU fallback.") from e
    # MMH_LIBRARY: load another build of the same sources (debug hook: scripts/build_racecheck_variant.sh builds the variant whose
    # tile hand-off uses named barriers so that compute-sanitizer's racecheck can follow it)
    lib = ctypes.CDLL(os.environ.get("MMH_LIBRARY") or SO_PATH)
    vp, i64, ci, dbl = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_double
    p64 = ctypes.POINTER(ctypes.c_int64)
    sig = {
        "mmh_version": ([], ci),
        "mmh_error_string": ([ci], ctypes.c_char_p),
        "mmh_device_count": ([ctypes.POINTER(ci)], ci),
        "mmh_set_device": ([ci], ci),
        "mmh_device_synchronize": ([], ci),
        "mmh_launch_count": ([], i64),
        "mmh_host_alloc": ([ctypes.POINTER(vp), i64], ci),
        "mmh_host_free": ([vp], ci),
        "mmh_forward": ([ci, p64, vp, vp, vp, vp, ci, vp], ci),
        "mmh_forward_host": ([ci, p64, vp, vp, vp, vp, ci], ci),
        "mmh_forward_batched": ([i64, ci, p64, vp, vp, vp, vp, ci, vp], ci),
        "mmh_forward_contract": ([i64, ci, p64, ci, vp, vp, vp, vp, ci, vp], ci),
        "mmh_forward_contract_host": ([i64, ci, p64, ci, vp, vp, vp, vp, ci], ci),
        "mmh_debug_timeline": ([vp], ci),
        "mmh_debug_plan": ([ci, ci, p64, ci, ctypes.POINTER(ci)], ci),
        "mmh_forward_panel_range": ([ci, p64, vp, vp, vp, ci, i64, i64, i64, vp], ci),
        "mmh_forward_batched_host": ([i64, ci, p64, vp, vp, vp, vp, ci], ci),
        "mmh_vjp": ([ci, p64, vp, vp, vp, vp, vp, vp, vp], ci),
        "mmh_vjp_host": ([ci, p64, vp, vp, vp, vp, vp, vp], ci),
        "mmh_vjp_batched": ([i64, ci, p64, vp, vp, vp, vp, vp, vp, vp], ci),
        "mmh_vjp_batched_host": ([i64, ci, p64, vp, vp, vp, vp, vp, vp], ci),
        "mmh_binomial": ([ci, p64, vp, vp, vp, dbl, i64, vp, ctypes.POINTER(dbl), vp], ci),
        "mmh_binomial_host": ([ci, p64, vp, vp, vp, dbl, i64, vp, ctypes.POINTER(dbl)], ci),
        "mmh_diagonal": ([ci, p64, vp, vp, i64, vp, vp, vp], ci),
        "mmh_diagonal_host": ([ci, p64, vp, vp, i64, vp, vp], ci),
        "mmh_diagonal_grad": ([ci, p64, vp, vp, vp, vp, vp, vp, vp], ci),
        "mmh_diagonal_grad_host": ([ci, p64, vp, vp, vp, vp, vp, vp], ci),
        "mmh_1leftover_grad": ([ci, p64, vp, vp, vp, vp, vp, vp, vp], ci),
        "mmh_1leftover_grad_host": ([ci, p64, vp, vp, vp, vp, vp, vp], ci),
        "mmh_1leftover": ([ci, p64, vp, vp, vp, vp, vp], ci),
        "mmh_1leftover_host": ([ci, p64, vp, vp, vp, vp], ci),
        "mmh_squeezer": ([i64, i64, dbl, dbl, vp, vp], ci),
        "mmh_squeezed": ([i64, dbl, dbl, vp, vp], ci),
        "mmh_beamsplitter": ([p64, dbl, dbl, ci, vp, vp], ci),
        "mmh_displacement": ([i64, i64, dbl, dbl, vp, vp], ci),
        "mmh_gate_host": ([ci, p64, dbl, dbl, vp], ci),
        "mmh_displacement_jacobian": ([i64, i64, vp, dbl, dbl, vp, vp, vp], ci),
        "mmh_displacement_grad": ([i64, vp, dbl, dbl, vp, vp, vp], ci),
        "mmh_displacement_derivs_host": ([ci, i64, i64, vp, dbl, dbl, vp, vp], ci),
        "mmh_gate_vjp": ([ci, ci, p64, vp, vp, vp, vp], ci),
        "mmh_gate_vjp_host": ([ci, ci, p64, vp, vp, vp], ci),
        "mmh_autoshape": ([ci, vp, vp, vp, dbl, i64, i64, vp, vp], ci),
        "mmh_autoshape_host": ([ci, vp, vp, vp, dbl, i64, i64, p64], ci),
        "mmh_fock_contract": ([ci, p64, ctypes.POINTER(ci), ci, p64, ctypes.POINTER(ci), ci, ctypes.POINTER(ci), vp, vp, vp, p64, vp], ci),
        "mmh_fock_contract_host": ([ci, p64, ctypes.POINTER(ci), ci, p64, ctypes.POINTER(ci), ci, ctypes.POINTER(ci), vp, vp, vp], ci),
        "mmh_fock_reduce": ([ci, p64, p64, vp, vp, vp], ci),
        "mmh_overlap": ([i64, vp, vp, vp, vp], ci),
    }
    for name, (args, res) in sig.items():
        fn = getattr(lib, name)
        fn.argtypes = args
        fn.restype = res
    return lib


lib = _load()


def check(status: int) -> None:
    if status == MMH_OK:
        return
    msg = lib.mmh_error_string(status).decode()
    if status > 0:
        raise RuntimeError(f"libmmhermite CUDA error {status}: {msg}")
    raise _EXC.get(status, RuntimeError)(f"libmmhermite: {msg}")


def shape_array(shape):
    arr = (ctypes.c_int64 * len(shape))(*[int(s) for s in shape])
    return arr


def device_count() -> int:
    n = ctypes.c_int(0)
    lib.mmh_device_count(ctypes.byref(n))
    return n.value


def launch_count() -> int:
    return int(lib.mmh_launch_count())


# ---- pinned result buffers ------------------------------------------------------------------------
# Results handed back to numpy live in page-locked memory so that the D2H copy is a single DMA at PCIe
# speed (a pageable destination is bounced through a driver staging buffer).  Freed blocks are cached by
# size; the cache is bounded.
class _PinnedPool:
    def __init__(self, max_cached_bytes: int = 8 << 30):
        self._free: dict[int, list[int]] = {}
        self._cached = 0
        self._max = max_cached_bytes
        self._lock = threading.Lock()

    def acquire(self, nbytes: int) -> int:
        with self._lock:
            lst = self._free.get(nbytes)
            if lst:
                self._cached -= nbytes
                return lst.pop()
        ptr = ctypes.c_void_p()
        check(lib.mmh_host_alloc(ctypes.byref(ptr), nbytes))
        return ptr.value

    def release(self, ptr: int, nbytes: int) -> None:
        with self._lock:
            if self._cached + nbytes <= self._max:
                self._free.setdefault(nbytes, []).append(ptr)
                self._cached += nbytes
                return
        try:
            lib.mmh_host_free(ctypes.c_void_p(ptr))
        except Exception:  # pragma: no cover - interpreter shutdown
            pass


_pool = _PinnedPool()


class _PinnedBlock:
    __slots__ = ("ptr", "nbytes", "__weakref__")

    def __init__(self, nbytes: int):
        self.nbytes = max(int(nbytes), 16)
        self.ptr = _pool.acquire(self.nbytes)

    @property
    def __array_interface__(self):
        return {"shape": (self.nbytes,), "typestr": "|u1", "data": (self.ptr, False), "version": 3}

    def __del__(self):
        try:
            _pool.release(self.ptr, self.nbytes)
        except Exception:  # pragma: no cover
            pass


def pinned_empty(shape, dtype=np.complex128) -> np.ndarray:
    """An uninitialised numpy array in page-locked host memory (freed/recycled when garbage collected)."""
    shape = tuple(int(s) for s in shape)
    dt = np.dtype(dtype)
    n = int(np.prod(shape, dtype=np.int64)) if shape else 1
    blk = _PinnedBlock(n * dt.itemsize)
    flat = np.asarray(blk)[: n * dt.itemsize].view(dt)
    return flat.reshape(shape)
