"""In-tree build of libmmhermite.so (hand-written CUDA for sm_100a) with nvcc.

    python -m mrmustard_b200.build [--force]

The shared library is written next to the sources (mrmustard_b200/csrc/libmmhermite.so, git-ignored,
shipped to the GPU box by gpurun).  nvcc cross-compiles without a GPU.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
SO = os.path.join(CSRC, "libmmhermite.so")
SOURCES = ["mmh_api.cu", "mmh_forward.cu", "mmh_march.cu", "mmh_lanes.cu", "mmh_box.cu", "mmh_tiled.cu", "mmh_rows.cu", "mmh_stable_boxes.cu", "mmh_vjp.cu", "mmh_diagonal.cu", "mmh_diagonal_rolling.cu", "mmh_gates.cu", "mmh_autoshape.cu", "mmh_einsum.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "--fmad=false",            # forward kernels also use explicit _rn intrinsics; belt and braces
    "-Xcompiler", "-fPIC", "-shared",
    "-Xptxas", "-v",
    "-lcudart",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def have_nvcc() -> bool:
    try:
        _nvcc()
        return True
    except RuntimeError:
        return False


def sources() -> list[str]:
    return [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]


def _source_hash() -> str:
    import hashlib
    h = hashlib.sha256()
    deps = sorted(sources() + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cuh")])
    deps.append(os.path.join(os.path.dirname(HERE), "include", "mmhermite.h"))
    for d in deps:
        h.update(os.path.basename(d).encode())
        h.update(open(d, "rb").read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def is_stale() -> bool:
    """Content-based (not mtime-based): a copied tree (gpurun snapshot) never triggers a spurious rebuild."""
    stamp = SO + ".srchash"
    if not os.path.exists(SO) or not os.path.exists(stamp):
        return True
    return open(stamp).read().strip() != _source_hash()


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not is_stale():
        return SO
    cmd = [_nvcc(), *NVCC_FLAGS, *os.environ.get("MMH_NVCC_EXTRA", "").split(), "-o", SO, *sources()]
    res = subprocess.run(cmd, capture_output=True, text=True)
    log = res.stdout + res.stderr
    with open(os.path.join(CSRC, "build.log"), "w") as f:
        f.write(" ".join(cmd) + "\n" + log)
    if res.returncode != 0:
        sys.stderr.write(log)
        raise RuntimeError("nvcc failed building libmmhermite.so")
    with open(SO + ".srchash", "w") as f:
        f.write(_source_hash())
    if verbose:
        print(log)
    return SO


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
