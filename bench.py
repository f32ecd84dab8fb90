#!/usr/bin/env python
"""bench.py — Gaussian-to-Fock hot path on B200: Fock amplitudes/s (complex128), with roofline and CPU baseline.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg2|cfg3]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
           bench.py --gpus N --steps K --warmup W

Workload (BASELINE.json configs[1], "cfg2"): the 2-mode BSgate+Sgate unitary's Bargmann triple (exact bytes
from tests/golden/vanilla_golden.npz) -> Fock lattice of shape (50,50,50,50) through `hermite_renormalized`
(vanilla strategy): one lattice of 6.25 M complex128 amplitudes (100 MB) per step per GPU.  With N GPUs every
rank fills its own lattice (independent units, no collective on the data path): weak scaling.

  value : whole-job amplitudes/s with the triple resident in HBM (device-pointer C-ABI call mmh_forward), timed
          with CUDA events per step; L2 is flushed (untimed) between steps because one lattice (100 MB) is
          smaller than the 126 MB L2.
  e2e   : the same metric through the numpy-facing plugin call (strategies.vanilla_numba -> mmh_forward_host):
          host (A,b,c) -> H2D -> kernels -> D2H of the whole lattice into page-locked host memory, inside the
          timed region.
  roofline : dominant kernel's algorithmic bytes (16 B per amplitude) / its measured duration,
          against the measured HBM copy bandwidth (MEASURED_PEAKS.json).
  cpu_baseline : the oracle's C port of vanilla_numba (strict IEEE, -O2) on one host core, same lattice.

`--impl reference` times that CPU port alone (the reference's algorithm is single-threaded for a single
lattice by construction, vanilla/core.py:25-124) and prints the same line with "impl": "reference".
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "Fock amplitudes/s (complex128)"
UNIT = "amplitudes/s"
ALGO_BYTES_PER_AMP_FWD = 16.0   # each complex128 amplitude is written once
ALGO_BYTES_PER_AMP_VJP = 32.0   # read G and dLdG once


def random_triple(n, batch=(), seed=None):
    """Reference synthetic-triple recipe (tests/test_math/test_lattice/test_vanilla.py:24-35), restated."""
    rng = np.random.RandomState(seed)
    A = rng.random((*batch, n, n)) + 1j * rng.random((*batch, n, n))
    A = A + np.swapaxes(A, -1, -2)
    A /= np.abs(np.linalg.eigvals(A)).max() + 0.2
    b = rng.random((*batch, n)) + 1j * rng.random((*batch, n))
    c = rng.random(batch) + 1j * rng.random(batch)
    return A, b, c


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def workload_inputs(workload: str, rank: int):
    if workload == "cfg2":
        gold = np.load(os.path.join(ROOT, "tests", "golden", "vanilla_golden.npz"))
        if rank == 0:
            A, b, c = gold["cfg2_A"], gold["cfg2_b"], gold["cfg2_c"].reshape(1)
            sha = str(gold["cfg2_G50_sha"])
        else:  # other ranks: raw-kernel variants of cfg2, one independent lattice each
            A, b, c = random_triple(4, (), seed=rank)
            c = np.asarray(c).reshape(1)
            sha = str(gold["cfg2r_G50_sha"]) if rank == 1 else None
        return dict(A=np.ascontiguousarray(A), b=np.ascontiguousarray(b), c=np.ascontiguousarray(c),
                    shape=(50, 50, 50, 50), batch=None, sha=sha)
    if workload == "cfg3":
        A, b, c = random_triple(2, (65536,), seed=673 + rank)
        gold = np.load(os.path.join(ROOT, "tests", "golden", "vanilla_golden.npz"))
        return dict(A=A, b=b, c=c, shape=(40, 40), batch=65536, sha=None,
                    first4=gold["cfg3_G_first4"] if rank == 0 else None)
    raise SystemExit(f"unknown workload {workload}")


# ---------------------------------------------------------------------------------------------------
class ClockSampler:
    """Samples SM clock and throttle reasons through NVML while the timed region runs."""

    def __init__(self, index: int):
        self.samples, self.reasons = [], set()
        self.max_mhz = None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        names = {
            "hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
            "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
            "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
            "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4),
        }
        while not self._stop.is_set():
            try:
                self.samples.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.004)

    def start(self):
        if self.nv is not None:
            self._thr = threading.Thread(target=self._run, daemon=True)
            self._thr.start()

    def stop(self):
        self._stop.set()
        if self._thr is not None:
            self._thr.join()
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


def cpu_baseline(w, budget_s=12.0):
    """Oracle C port (kind=port) on the host cores; bounded sample of the same workload."""
    import oracle
    oracle.build()
    shape = w["shape"]
    n_per = int(np.prod(shape))
    if w["batch"] is None:
        t0 = time.perf_counter(); oracle.vanilla(shape, w["A"], w["b"], complex(w["c"][0])); t1 = time.perf_counter()
        reps = max(1, min(20, int(budget_s / max(t1 - t0, 1e-3))))
        best = t1 - t0
        for _ in range(reps):
            t0 = time.perf_counter(); oracle.vanilla(shape, w["A"], w["b"], complex(w["c"][0])); t1 = time.perf_counter()
            best = min(best, t1 - t0)
        return {"value": n_per / best, "unit": UNIT, "cores": 1, "kind": "port",
                "sample": f"full {shape} lattice, best of {reps + 1} runs of the C port of vanilla_numba (single-threaded by construction)"}
    cores = os.cpu_count() or 1
    B = 8192
    A, b, c = w["A"][:B].copy(), w["b"][:B].copy(), w["c"][:B].copy()
    best = 1e30
    for _ in range(3):
        t0 = time.perf_counter(); oracle.vanilla_batch(shape, A, b, c, nthreads=cores); t1 = time.perf_counter()
        best = min(best, t1 - t0)
    return {"value": B * n_per / best, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"first {B} of 65536 triples, shape {shape}, best of 3, pthreads over the batch (= prange)"}


def run_reference(args, w):
    """--impl reference: the reference algorithm's CPU port, rank 0 only."""
    import oracle
    oracle.build()
    shape = w["shape"]
    n_per = int(np.prod(shape))
    cores = 1 if w["batch"] is None else (os.cpu_count() or 1)
    if w["batch"] is None:
        def step():
            oracle.vanilla(shape, w["A"], w["b"], complex(w["c"][0]))
        amps = n_per
        sample = f"one full {shape} lattice per step (C port of vanilla_numba, 1 thread by construction)"
    else:
        B = 4096
        A, b, c = w["A"][:B].copy(), w["b"][:B].copy(), w["c"][:B].copy()
        def step():
            oracle.vanilla_batch(shape, A, b, c, nthreads=cores)
        amps = B * n_per
        sample = f"{B} of 65536 triples per step, shape {shape}, {cores} threads over the batch"
    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    val = amps * args.steps / dt
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": args.workload_name, "shape": list(shape)},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=["cfg2", "cfg3"])
    ap.add_argument("--no-extras", action="store_true", help="skip the secondary measurements (vjp, batched)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg (profiling runs only)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    args.workload_name = {
        "cfg2": "cfg2: 2-mode BSgate+Sgate unitary -> hermite_renormalized shape (50,50,50,50), one lattice per GPU per step",
        "cfg3": "cfg3: hermite_renormalized_batched, 65,536 random 2-mode triples per GPU per step, cutoff 40",
    }[args.workload]

    if args.impl == "reference":
        if rank == 0:
            run_reference(args, workload_inputs(args.workload, 0))
        return

    import torch
    import torch.distributed as dist
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    from mrmustard_b200 import _lib, strategies
    lib, check = _lib.lib, _lib.check

    w = workload_inputs(args.workload, rank)
    shape = w["shape"]
    D = len(shape)
    n_per = int(np.prod(shape))
    B = w["batch"] or 1
    amps_per_step = B * n_per
    sh = _lib.shape_array(shape)

    dA = torch.from_numpy(w["A"]).to(dev)
    db = torch.from_numpy(w["b"]).to(dev)
    dc = torch.from_numpy(w["c"]).to(dev)
    dG = torch.empty((B, n_per), dtype=torch.complex128, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2
    need_flush = amps_per_step * 16 < (200 << 20)
    stream = torch.cuda.current_stream()
    sptr = ctypes.c_void_p(stream.cuda_stream)

    def launch():
        if w["batch"] is None:
            check(lib.mmh_forward(D, sh, dA.data_ptr(), db.data_ptr(), dc.data_ptr(), dG.data_ptr(), 0, sptr))
        else:
            check(lib.mmh_forward_batched(B, D, sh, dA.data_ptr(), db.data_ptr(), dc.data_ptr(), dG.data_ptr(), 0, sptr))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- correctness of what is being timed (untimed) ---------------------------------------------
    launch(); torch.cuda.synchronize()
    host = dG.cpu().numpy()
    import hashlib
    if w.get("sha"):
        got = hashlib.sha256((host.reshape(shape) + 0.0).tobytes()).hexdigest()
        assert got == w["sha"], "bench: device result differs from the reference's golden sha256"
    if w.get("first4") is not None:
        assert np.array_equal(host[:4].reshape(4, *shape), w["first4"]), "bench: cfg3 result differs from golden"
    del host

    # ---- device-resident timing ---------------------------------------------------------------------
    for _ in range(args.warmup):
        if need_flush:
            flush.fill_(1)
        launch()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    l0 = _lib.launch_count()
    barrier()
    for s in range(args.steps):
        if need_flush:
            flush.fill_(s & 0xff)      # untimed: evict the previous lattice from L2
        ev[s][0].record(stream)
        launch()
        ev[s][1].record(stream)
    barrier()
    launches = _lib.launch_count() - l0
    step_ms = [a.elapsed_time(b) for a, b in ev]
    total_ms = float(sum(step_ms))

    # ---- end to end through the numpy-facing plugin call (host buffers, copies inside the timed region) --
    hA, hb, hc = w["A"], w["b"], w["c"]
    def e2e_call():
        if w["batch"] is None:
            return strategies.vanilla_numba(shape, hA, hb, complex(hc[0]))
        return strategies.vanilla_batch_numba(shape, hA, hb, hc)
    e2e_steps = max(3, min(args.steps, 10))
    for _ in range(2):
        r = e2e_call(); del r
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ds = torch.cuda.default_stream()
    t0 = time.perf_counter()
    e0.record(ds)
    chk = 0.0
    for _ in range(e2e_steps):
        r = e2e_call()
        chk += float(r.flat[-1].real)  # touch the host result
        del r
    e1.record(ds)
    barrier()
    e2e_wall_ms = 1e3 * (time.perf_counter() - t0)
    e2e_ms = max(float(e0.elapsed_time(e1)), 0.0)
    e2e_ms = max(e2e_ms, e2e_wall_ms * 0.0)  # event time is the reported one; wall kept alongside
    clocks = sampler.stop()

    # ---- max over ranks -------------------------------------------------------------------------------
    t = torch.tensor([total_ms, e2e_ms, e2e_wall_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, e2e_ms, e2e_wall_ms = (float(x) for x in t.cpu())

    peak, peak_src = load_peaks()
    value = world * amps_per_step * args.steps / (total_ms * 1e-3)
    e2e_value = world * amps_per_step * e2e_steps / (e2e_ms * 1e-3)
    avg_kernel_ms = float(np.mean(step_ms))
    achieved = ALGO_BYTES_PER_AMP_FWD * amps_per_step / (avg_kernel_ms * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    # one step = one mmh_forward[_batched] call = the launches listed here; the roofline figure is the whole step's
    # algorithmic bytes over the whole step's device time (a lower bound of the dominant kernel's own figure)
    kernel_name = ("mmh_forward: memset(panel 0) + k_warp_tail + k_march_tiled2<1,2> + k_march_tiled2<2,3> (dominant, 66% of the serialised "
                   "kernel time; the three kernels overlap)"
                   if w["batch"] is None else "mmh_forward_batched: k_fwd_chain_rows + k_march_lanes<5> (dominant, >90% of the step)")
    traffic_note = None
    if os.path.exists(tpath):
        try:
            tj = json.load(open(tpath)).get(args.workload, {})
            traffic = tj.get("dram_bytes_per_launch")
            traffic_note = f"dram__bytes_read+write of {tj.get('kernel')} from profiles/{tj.get('source')} (ncu --set full), per launch"
        except Exception:
            traffic = None

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": args.workload_name, "shape": list(shape), "batch_per_gpu": B,
                   "amplitudes_per_step_per_gpu": amps_per_step,
                   "l2": "flushed between timed steps (256 MiB write)" if need_flush else "per-step output larger than L2",
                   "parity": "device result sha256-checked against the reference golden before timing"},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "steps": e2e_steps, "ms_per_step": e2e_ms / e2e_steps,
                "wall_ms_per_step": e2e_wall_ms / e2e_steps,
                "h2d_bytes_per_step": int(hA.nbytes + hb.nbytes + hc.nbytes), "d2h_bytes_per_step": int(amps_per_step * 16),
                "api": ("mrmustard_b200.strategies.vanilla_numba -> mmh_forward_host (pinned result buffer)" if w["batch"] is None else
                        "mrmustard_b200.strategies.vanilla_batch_numba -> mmh_forward_batched_host (pinned result buffer)")},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "kernel": kernel_name, "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": traffic, "traffic_note": traffic_note, "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": ALGO_BYTES_PER_AMP_FWD * amps_per_step,
                     "avg_launch_ms": avg_kernel_ms},
    }
    if rank == 0:
        line["cpu_baseline"] = None if args.no_cpu else cpu_baseline(w)
        if not args.no_extras and world == 1:
            line["also"] = extras(torch, dev, lib, check, _lib, stream, sptr, flush)
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def extras(torch, dev, lib, check, _lib, stream, sptr, flush):
    """Secondary measurements on the other BASELINE configs (not bench lines): device-resident, CUDA events."""
    out = {}
    peak, _ = load_peaks()

    def timeit(fn, reps, flush_l2):
        fn(); torch.cuda.synchronize()
        ms = []
        for _ in range(reps):
            if flush_l2:
                flush.fill_(3)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream); fn(); b.record(stream)
            torch.cuda.synchronize()
            ms.append(a.elapsed_time(b))
        return float(np.median(ms))

    try:
        # cfg3: 65,536 triples x (40,40): forward + vjp
        A, b, c = random_triple(2, (65536,), seed=673)
        dA, db, dc = (torch.from_numpy(x).to(dev) for x in (A, b, c))
        shape = (40, 40); sh = _lib.shape_array(shape); n = 1600; B = 65536
        dG = torch.empty((B, n), dtype=torch.complex128, device=dev)
        ms = timeit(lambda: check(lib.mmh_forward_batched(B, 2, sh, dA.data_ptr(), db.data_ptr(), dc.data_ptr(), dG.data_ptr(), 0, sptr)), 5, False)
        out["cfg3_forward"] = {"amps_per_s": B * n / (ms * 1e-3), "ms": ms, "hbm_frac": 16.0 * B * n / (ms * 1e-3) / 1e9 / peak}
        g = torch.randn((B, n), dtype=torch.float64, device=dev).to(torch.complex128)
        oA = torch.empty((B, 2, 2), dtype=torch.complex128, device=dev)
        ob = torch.empty((B, 2), dtype=torch.complex128, device=dev)
        oc = torch.empty((B,), dtype=torch.complex128, device=dev)
        ms = timeit(lambda: check(lib.mmh_vjp_batched(B, 2, sh, dG.data_ptr(), dc.data_ptr(), g.data_ptr(), oA.data_ptr(), ob.data_ptr(), oc.data_ptr(), sptr)), 5, False)
        out["cfg3_vjp"] = {"amps_per_s": B * n / (ms * 1e-3), "ms": ms, "hbm_frac": 32.0 * B * n / (ms * 1e-3) / 1e9 / peak}
        del dG, g
        # cfg5: 4-mode ket, cutoff 40: forward + vjp
        gold = np.load(os.path.join(ROOT, "tests", "golden", "vanilla_golden.npz"))
        A, b, c = gold["cfg5_A"], gold["cfg5_b"], gold["cfg5_c"].reshape(1)
        dA, db, dc = (torch.from_numpy(np.ascontiguousarray(x)).to(dev) for x in (A, b, c))
        shape = (40,) * 4; sh = _lib.shape_array(shape); n = 40 ** 4
        dG = torch.empty((n,), dtype=torch.complex128, device=dev)
        ms = timeit(lambda: check(lib.mmh_forward(4, sh, dA.data_ptr(), db.data_ptr(), dc.data_ptr(), dG.data_ptr(), 0, sptr)), 10, True)
        out["cfg5_forward"] = {"amps_per_s": n / (ms * 1e-3), "ms": ms, "hbm_frac": 16.0 * n / (ms * 1e-3) / 1e9 / peak}
        g = torch.randn((n,), dtype=torch.float64, device=dev).to(torch.complex128)
        oA = torch.empty((4, 4), dtype=torch.complex128, device=dev)
        ob = torch.empty((4,), dtype=torch.complex128, device=dev)
        oc = torch.empty((1,), dtype=torch.complex128, device=dev)
        ms = timeit(lambda: check(lib.mmh_vjp(4, sh, dG.data_ptr(), dc.data_ptr(), g.data_ptr(), oA.data_ptr(), ob.data_ptr(), oc.data_ptr(), sptr)), 10, True)
        out["cfg5_vjp"] = {"amps_per_s": n / (ms * 1e-3), "ms": ms, "hbm_frac": 32.0 * n / (ms * 1e-3) / 1e9 / peak}
        # cfg2 as a batch: 8 x (50,)^4 through hermite_renormalized_batched (consecutive lattices pipelined on the device)
        A, b, c = gold["cfg2_A"], gold["cfg2_b"], gold["cfg2_c"].reshape(1)
        Bq = 8
        dA, db, dc = (torch.from_numpy(np.ascontiguousarray(np.repeat(x[None], Bq, 0))).to(dev) for x in (A, b, c))
        shape = (50,) * 4; sh = _lib.shape_array(shape); n = 50 ** 4
        dG = torch.empty((Bq, n), dtype=torch.complex128, device=dev)
        ms = timeit(lambda: check(lib.mmh_forward_batched(Bq, 4, sh, dA.data_ptr(), db.data_ptr(), dc.data_ptr(), dG.data_ptr(), 0, sptr)), 5, True)
        out["cfg2_batch8_forward"] = {"amps_per_s": Bq * n / (ms * 1e-3), "ms": ms, "hbm_frac": 16.0 * Bq * n / (ms * 1e-3) / 1e9 / peak}
        del dG
        # a batch of 2-mode unitaries at cutoff 20: 148 x (20,)^4 through hermite_renormalized_batched (box march, mmh_box.cu)
        Bq = 148
        A, b, c = random_triple(4, (Bq,), seed=3)
        dA, db, dc = (torch.from_numpy(np.ascontiguousarray(x)).to(dev) for x in (A, b, c))
        shape = (20,) * 4; sh = _lib.shape_array(shape); n = 20 ** 4
        dG = torch.empty((Bq, n), dtype=torch.complex128, device=dev)
        ms = timeit(lambda: check(lib.mmh_forward_batched(Bq, 4, sh, dA.data_ptr(), db.data_ptr(), dc.data_ptr(), dG.data_ptr(), 0, sptr)), 5, True)
        out["batch148_20p4_forward"] = {"amps_per_s": Bq * n / (ms * 1e-3), "ms": ms, "hbm_frac": 16.0 * Bq * n / (ms * 1e-3) / 1e9 / peak}
        del dG
        # lattice + derived-variable contraction, end to end through the numpy-facing call (host buffers):
        # fused (only the contracted array is copied back) vs materialising the lattice on the host and einsum there
        import mrmustard_b200 as mm
        A4, b4, _ = random_triple(4, (), seed=21)
        core, der = (48, 48), (8, 40)
        cpoly = np.random.RandomState(3).standard_normal(der) + 0j
        def wall(fn, reps=3):
            fn(); best = 1e30
            for _ in range(reps):
                t0 = time.perf_counter(); fn(); best = min(best, time.perf_counter() - t0)
            return best * 1e3
        ms_f = wall(lambda: mm.hermite_renormalized_contracted(A4, b4, cpoly, core))
        ms_m = wall(lambda: np.einsum("abk,k->ab", mm.strategies.vanilla_numba(core + der, A4, b4, 1.0).reshape(core + (-1,)), cpoly.reshape(-1)))
        out["contract_48x48_x_8x40"] = {"fused_ms": ms_f, "materialised_ms": ms_m, "lattice_amplitudes": int(np.prod(core + der))}
        # cfg4: (i) M = 4 diagonal and one-leftover-mode sweeps at cutoff 12 through the numpy-facing calls,
        # (ii) the 8-mode ket as one vanilla lattice (12,)^8 = 430 M amplitudes, device resident
        gd = np.load(os.path.join(ROOT, "tests", "golden", "diagonal_golden.npz"))
        Ad, bd, cd = gd["d4_A"], gd["d4_b"], complex(gd["d4_c"])
        out["cfg4_diagonal_M4_cutoff12_ms"] = wall(lambda: mm.hermite_renormalized_diagonal(Ad, bd, cd, (12,) * 4))
        Al, bl, cl = gd["l4_A"], gd["l4_b"], complex(gd["l4_c"])
        out["cfg4_1leftover_M4_cutoff12_ms"] = wall(lambda: mm.hermite_renormalized_1leftoverMode(Al, bl, cl, 11, (11, 11, 11)))
        A8k, b8k, c8k = gold["cfg4_A"], gold["cfg4_b"], complex(gold["cfg4_c"])
        Adm = np.zeros((16, 16), complex); Adm[:8, :8] = np.conj(A8k); Adm[8:, 8:] = A8k
        bdm = np.concatenate([np.conj(b8k), b8k])
        out["cfg4_diagonal_M8_cutoff6_ms"] = wall(lambda: mm.hermite_renormalized_diagonal(Adm, bdm, abs(c8k) ** 2, (6,) * 8), 2)
        A8, b8, c8 = gold["cfg4_A"], gold["cfg4_b"], gold["cfg4_c"].reshape(1)
        dA, db, dc = (torch.from_numpy(np.ascontiguousarray(x)).to(dev) for x in (A8, b8, c8))
        shape = (12,) * 8; sh = _lib.shape_array(shape); n = 12 ** 8
        dG = torch.empty((n,), dtype=torch.complex128, device=dev)
        ms = timeit(lambda: check(lib.mmh_forward(8, sh, dA.data_ptr(), db.data_ptr(), dc.data_ptr(), dG.data_ptr(), 0, sptr)), 2, False)
        out["cfg4_vanilla_12p8_forward"] = {"amps_per_s": n / (ms * 1e-3), "ms": ms, "hbm_frac": 16.0 * n / (ms * 1e-3) / 1e9 / peak}
        del dG
        # cfg1: latency config
        A, b, c = gold["cfg1_A"], gold["cfg1_b"], gold["cfg1_c"].reshape(1)
        dA, db, dc = (torch.from_numpy(np.ascontiguousarray(x)).to(dev) for x in (A, b, c))
        sh = _lib.shape_array((200,))
        dG = torch.empty((200,), dtype=torch.complex128, device=dev)
        ms = timeit(lambda: check(lib.mmh_forward(1, sh, dA.data_ptr(), db.data_ptr(), dc.data_ptr(), dG.data_ptr(), 0, sptr)), 20, False)
        out["cfg1_forward_us"] = ms * 1e3
    except Exception as e:  # extras must never break the contract line
        out["error"] = repr(e)
    return out


if __name__ == "__main__":
    main()
