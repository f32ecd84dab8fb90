#!/usr/bin/env python
"""Single-lattice march probe: k_march_rows (mmh_rows.cu) against k_march_tiled2 (MMH_NO_ROWS=1) on the same random triples --
bit-exact comparison and CUDA-event timing with an L2 flush between runs.

    python scripts/probe_single.py [shape ...]      e.g.  50,50,50,50  40,40,40,40  33,47,29,31
Environment hooks are passed through (MMH_ROWS_G, MMH_ROWS_R, MMH_ROWS_NPD_MIN, ...)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mrmustard_b200 import _lib  # noqa: E402


def triple(D, seed):
    rng = np.random.default_rng(seed)
    A = rng.uniform(-1, 1, (D, D)) + 1j * rng.uniform(-1, 1, (D, D))
    A = (A + A.T) / 2
    A /= np.abs(np.linalg.eigvals(A)).max() * 1.5      # amplitudes stay finite over the whole lattice
    b = rng.uniform(-1, 1, D) + 1j * rng.uniform(-1, 1, D)
    c = np.array([0.4 + 0.3j])
    return A, b, c


def main():
    shapes = [tuple(int(x) for x in a.split(",")) for a in sys.argv[1:]] or [(50,) * 4, (40,) * 4]
    dev = torch.device("cuda:0")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    reps = int(os.environ.get("PROBE_REPS", "10"))
    for shape in shapes:
        D = len(shape)
        A, b, c = triple(D, 7 + D)
        dA, db, dc = (torch.from_numpy(np.ascontiguousarray(x)).to(dev) for x in (A, b, c))
        sh = _lib.shape_array(shape)
        out = {}
        for mode in ("tiled2", "rows"):
            if mode == "tiled2":
                os.environ["MMH_NO_ROWS"] = "1"
            else:
                os.environ.pop("MMH_NO_ROWS", None)
            G = torch.full(shape, float("nan"), dtype=torch.complex128, device=dev)

            def run():
                _lib.check(_lib.lib.mmh_forward(D, sh, dA.data_ptr(), db.data_ptr(), dc.data_ptr(), G.data_ptr(), 0, None))

            for _ in range(3):
                run()
            torch.cuda.synchronize()
            ts, hs = [], []
            import time
            for _ in range(reps):
                flush.fill_(1)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); h0 = time.perf_counter(); run(); h1 = time.perf_counter(); e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1) * 1e3)
                hs.append((h1 - h0) * 1e6)
            # back-to-back calls without a flush: the launch path of call n+1 overlaps the kernels of call n
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20):
                run()
            e1.record(); torch.cuda.synchronize()
            b2b = e0.elapsed_time(e1) * 1e3 / 20
            print(f"   {mode}: host time of the call {np.median(hs):.1f} us, back-to-back {b2b:.1f} us per lattice")
            out[mode] = (G.clone(), np.median(ts), np.min(ts))
        same = torch.equal(out["tiled2"][0].view(torch.float64).view(torch.int64), out["rows"][0].view(torch.float64).view(torch.int64))
        nbytes = 16 * int(np.prod(shape))
        print(f"{shape}: tiled2 {out['tiled2'][1]:8.1f} us (min {out['tiled2'][2]:.1f})   rows {out['rows'][1]:8.1f} us (min {out['rows'][2]:.1f})"
              f"   bit-identical {same}   rows = {nbytes / out['rows'][1] / 1e3:.0f} GB/s", flush=True)
        if not same:
            d = (out["tiled2"][0] != out["rows"][0]) | (torch.isnan(out["rows"][0].real))
            idx = torch.nonzero(d)
            print("   first mismatches:", idx[:8].tolist(), " count", int(d.sum()))


if __name__ == "__main__":
    main()
