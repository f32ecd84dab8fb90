#!/usr/bin/env python
"""bench.py — Gaussian-to-Fock hot path on B200: Fock amplitudes/s (complex128) + VJP, roofline, CPU baseline.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg3|cfg2|cfg4]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
           bench.py --gpus N --steps K --warmup W

Default workload (BASELINE.json configs[2], the config north_star quotes its multi-GPU target on): ONE fixed batch of
65,536 random 2-mode Bargmann triples (the reference's recipe, seed 673), cutoff 40 -> 104,857,600 complex128
amplitudes (1.68 GB) per step through `hermite_renormalized_batched` (vanilla strategy), then the batched VJP.
With N GPUs the SAME batch is sharded contiguously over the ranks (mrmustard_b200.sharding.shard_range): STRONG
scaling, no collective on the data path; the optional final gather is timed separately.  Every rank checks its
shard against the reference's golden sha256 (per 4096-triple chunk) before anything is timed.

  value    : whole-job forward amplitudes/s with the triples resident in HBM (device-pointer C-ABI call
             mmh_forward_batched), CUDA events around the K timed steps, max over ranks.  The per-rank output
             (>= 210 MB) is larger than the 126 MB L2, so consecutive steps cannot reuse cached lattices.
  e2e      : the same metric through the numpy-facing plugin call with HOST buffers
             (backend.hermite_renormalized_batched -> mmh_forward_batched_host): H2D of the triples, kernels, D2H of
             the lattices into page-locked host memory, all inside the timed region.
  vjp      : the second half of the metric ("+ VJP"): mmh_vjp_batched on the same shard, with its own roofline
             (32 B per amplitude), e2e (host G and dLdG -> H2D -> kernel -> D2H of the gradients) and CPU baseline.
  roofline : algorithmic bytes (16 B per amplitude forward, 32 B VJP) / measured device
             time per step, against the measured HBM copy bandwidth (MEASURED_PEAKS.json).
  cpu_baseline : the reference's own numba strategies (kind "reference", from baseline/_ref) on all host cores, on a
             bounded sample of the same batch; the oracle's C port (kind "port") only where no reference install exists.
  also     : complete secondary records of the other BASELINE configs (cfg2 single lattice with roofline / e2e /
             cpu_baseline, cfg5 forward + VJP + device-resident train step, cfg1, cfg4 ...), N = 1 only.

`--impl reference` times the reference's numba implementation alone on the same config (rank 0 only under torchrun).
`--workload cfg2` = one (50,)^4 lattice per GPU per step (weak scaling); `--workload cfg4` = ONE (12,)^8 lattice
sharded over the ranks by panel ranges (strong scaling, halo exchange).
"""
from __future__ import annotations

import argparse
import ctypes
import hashlib
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
CFG3_B, CFG3_SHAPE, CFG3_SEED = 65536, (40, 40), 673
REF_SAMPLE_TRIPLES = 8192       # triples per step of the CPU arms on cfg3 (bounded sample of the same batch)


def random_triple(n, batch=(), seed=None):
    """Reference synthetic-triple recipe (tests/test_math/test_lattice/test_vanilla.py:24-35), restated."""
    rng = np.random.RandomState(seed)
    A = rng.random((*batch, n, n)) + 1j * rng.random((*batch, n, n))
    A = A + np.swapaxes(A, -1, -2)
    A /= np.abs(np.linalg.eigvals(A)).max() + 0.2
    b = rng.random((*batch, n)) + 1j * rng.random((*batch, n))
    c = rng.random(batch) + 1j * rng.random(batch)
    return A, b, c


def sha(a) -> str:
    return hashlib.sha256((np.ascontiguousarray(a) + 0.0).tobytes()).hexdigest()


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def golden():
    return np.load(os.path.join(ROOT, "tests", "golden", "vanilla_golden.npz"))


WORKLOADS = {
    "cfg3": "cfg3: hermite_renormalized_batched + VJP over ONE batch of 65,536 random 2-mode triples (seed 673), cutoff 40, "
            "sharded over the GPUs (strong scaling)",
    "cfg2": "cfg2: 2-mode BSgate+Sgate unitary -> hermite_renormalized shape (50,50,50,50), one lattice per GPU per step",
    "cfg4": "cfg4: 8-mode Gaussian ket -> hermite_renormalized shape (12,)*8 (430 M amplitudes), ONE lattice sharded over the "
            "GPUs by panel ranges",
}
SCALING = {"cfg3": "strong", "cfg2": "weak", "cfg4": "strong"}


def config_for(workload: str, world: int) -> dict:
    """The `config` object -- built by this one function for BOTH arms (ours and --impl reference)."""
    if workload == "cfg3":
        return {"workload": WORKLOADS[workload], "shape": list(CFG3_SHAPE), "batch_total": CFG3_B, "modes": 2,
                "amplitudes_per_step": CFG3_B * int(np.prod(CFG3_SHAPE)), "strategy": "vanilla", "sharding": f"batch/{world}"}
    if workload == "cfg2":
        return {"workload": WORKLOADS[workload], "shape": [50] * 4, "batch_total": world, "modes": 2,
                "amplitudes_per_step": world * 50 ** 4, "strategy": "vanilla", "sharding": f"replicas x{world}"}
    return {"workload": WORKLOADS[workload], "shape": [12] * 8, "batch_total": 1, "modes": 8,
            "amplitudes_per_step": 12 ** 8, "strategy": "vanilla", "sharding": f"panel ranges /{world}"}


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
            time.sleep(0.002)

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


# ---------------------------------------------------------------------------------------------------
# CPU arms: the reference's numba strategies when an install exists (baseline/_ref), else the oracle's C port
# ---------------------------------------------------------------------------------------------------
def reference_strategies():
    """(module, kind, note): the unmodified reference's strategies (numba) or None."""
    try:
        from oracle import refimport
        if not refimport.available():
            return None, "port", "no reference install (baseline/_ref) in this tree"
        S = refimport.strategies()
        import numba
        return S, "reference", f"numba {numba.__version__}, NUMBA_NUM_THREADS={numba.config.NUMBA_NUM_THREADS}, from {refimport.REFERENCE_ROOT}"
    except Exception as e:  # pragma: no cover
        return None, "port", f"reference import failed: {e!r}"


class CpuArm:
    """The CPU implementation of one workload: callable steps on a bounded sample, for cpu_baseline and --impl reference."""

    def __init__(self, workload: str):
        self.S, self.kind, self.note = reference_strategies()
        self.cores = os.cpu_count() or 1
        g = golden()
        if self.S is None:
            import oracle
            oracle.build()
            self.oracle = oracle
        if workload == "cfg3":
            A, b, c = random_triple(2, (CFG3_B,), seed=CFG3_SEED)
            n = REF_SAMPLE_TRIPLES
            self.A, self.b, self.c = A[:n].copy(), b[:n].copy(), c[:n].copy()
            self.shape = CFG3_SHAPE
            self.amps = n * int(np.prod(CFG3_SHAPE))
            self.threads = self.cores
            self.sample = (f"first {n} of {CFG3_B} triples per step, shape {CFG3_SHAPE}: vanilla_batch_numba "
                           f"(prange over the batch, {self.cores} threads)")
            self.vjp_sample = self.sample.replace("vanilla_batch_numba", "vanilla_batch_vjp_numba")
            self._G = None
            self._g = np.random.RandomState(1).standard_normal((n, *CFG3_SHAPE)) + 0j
        elif workload == "cfg2":
            self.A, self.b, self.c = g["cfg2_A"], g["cfg2_b"], complex(g["cfg2_c"])
            self.shape = (50,) * 4
            self.amps = 50 ** 4
            self.threads = 1
            self.sample = "one full (50,50,50,50) lattice per step: vanilla_numba (single-threaded by construction, core.py:25-124)"
        elif workload == "cfg5":
            self.A, self.b, self.c = g["cfg5_A"], g["cfg5_b"], complex(g["cfg5_c"])
            self.shape = (40,) * 4
            self.amps = 40 ** 4
            self.threads = 1
            self.sample = "one full (40,40,40,40) lattice per step: vanilla_numba / vanilla_vjp_numba (single-threaded by construction)"
            self.vjp_sample = self.sample
            self._G = None
            self._g = np.random.RandomState(1).standard_normal(self.shape) + 0j
        else:  # cfg4: the full (12,)^8 lattice takes ~8 s single-threaded; bounded sample = the (12,)^7 sub-lattice k_0 = 0
            self.A, self.b, self.c = g["cfg4_A"][1:, 1:].copy(), g["cfg4_b"][1:].copy(), complex(g["cfg4_c"])
            self.shape = (12,) * 7
            self.amps = 12 ** 7
            self.threads = 1
            self.sample = "the (12,)^7 sub-lattice k_0 = 0 of the (12,)^8 lattice per step: vanilla_numba (single-threaded by construction)"
        self.batched = workload == "cfg3"

    def forward(self):
        if self.S is not None:
            if self.batched:
                return self.S.vanilla_batch_numba(self.shape, self.A, self.b, self.c, False, None)
            return self.S.vanilla_numba(self.shape, self.A, self.b, self.c, None)
        if self.batched:
            return self.oracle.vanilla_batch(self.shape, self.A, self.b, self.c, nthreads=self.cores)
        return self.oracle.vanilla(self.shape, self.A, self.b, self.c)

    def vjp(self):
        if self._G is None:
            self._G = np.ascontiguousarray(self.forward())
        if self.S is not None:
            if self.batched:
                return self.S.vanilla_batch_vjp_numba(self._G, self.c, self._g)
            return self.S.vanilla_vjp_numba(self._G, self.c, self._g)
        if self.batched:
            return self.oracle.vanilla_batch_vjp(self._G, self.c, self._g, nthreads=self.cores)
        return self.oracle.vanilla_vjp(self._G, self.c, self._g)

    def time(self, fn, budget_s=10.0, min_reps=2, max_reps=20):
        fn()   # JIT / page-in
        t0 = time.perf_counter(); fn(); first = time.perf_counter() - t0
        reps = max(min_reps, min(max_reps, int(budget_s / max(first, 1e-4))))
        best = first
        for _ in range(reps):
            t0 = time.perf_counter(); fn(); best = min(best, time.perf_counter() - t0)
        return best, reps + 1

    def baseline(self, which="forward", budget_s=10.0):
        fn = self.forward if which == "forward" else self.vjp
        best, reps = self.time(fn, budget_s)
        sample = self.sample if which == "forward" else self.vjp_sample
        return {"value": self.amps / best, "unit": UNIT, "cores": self.threads, "kind": self.kind,
                "sample": f"{sample}; best of {reps}", "note": self.note}


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path, rank 0 only, same config object."""
    arm = CpuArm(args.workload)
    arm.forward()   # JIT compile outside the timed region
    for _ in range(args.warmup):
        arm.forward()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        arm.forward()
    dt = time.perf_counter() - t0
    val = arm.amps * args.steps / dt
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
            "scaling": SCALING[args.workload], "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_for(args.workload, args.gpus),
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": arm.threads, "kind": arm.kind, "sample": arm.sample, "note": arm.note},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    if args.workload in ("cfg3",):
        best, reps = arm.time(arm.vjp, 6.0)
        line["vjp"] = {"value": arm.amps / best, "unit": UNIT, "sample": arm.vjp_sample + f"; best of {reps}"}
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg3", choices=["cfg3", "cfg2", "cfg4"])
    ap.add_argument("--no-extras", action="store_true", help="skip the secondary records in `also`")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline legs (profiling runs only)")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer legs (profiling runs only)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        if rank == 0:
            run_reference(args)
        return
    args.warmup = max(args.warmup, 3)

    import torch
    import torch.distributed as dist
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    from mrmustard_b200 import _lib
    ctx = dict(torch=torch, dist=dist, dev=dev, rank=rank, local_rank=local_rank, world=world, lib=_lib.lib, check=_lib.check,
               _lib=_lib, stream=torch.cuda.current_stream(), args=args)
    ctx["sptr"] = ctypes.c_void_p(ctx["stream"].cuda_stream)
    ctx["flush"] = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    line = {"cfg3": bench_cfg3, "cfg2": bench_cfg2, "cfg4": bench_cfg4}[args.workload](ctx)
    if rank == 0:
        if not args.no_extras and world == 1:
            line["also"] = extras(ctx, skip=args.workload)
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def _barrier(ctx):
    if ctx["world"] > 1:
        ctx["dist"].barrier()
    ctx["torch"].cuda.synchronize()


def _max_over_ranks(ctx, vals):
    torch = ctx["torch"]
    t = torch.tensor(list(vals), dtype=torch.float64, device=ctx["dev"])
    if ctx["world"] > 1:
        ctx["dist"].all_reduce(t, op=ctx["dist"].ReduceOp.MAX)
    return [float(x) for x in t.cpu()]


def _timed_steps(ctx, step, steps, warmup, flush_between=False):
    """W untimed warm-up steps, then exactly K steps bracketed by barrier + synchronize; device time from CUDA events on
    the launching stream.  Returns (total_ms, launches).  With flush_between every step is timed by its own event pair
    and an untimed 256 MiB write evicts L2 in between (per-step outputs smaller than L2)."""
    torch, stream, _lib = ctx["torch"], ctx["stream"], ctx["_lib"]
    for _ in range(warmup):
        if flush_between:
            ctx["flush"].fill_(1)
        step()
    _barrier(ctx)
    l0 = _lib.launch_count()
    if flush_between:
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        for s in range(steps):
            ctx["flush"].fill_(s & 0xff)
            ev[s][0].record(stream); step(); ev[s][1].record(stream)
        _barrier(ctx)
        total = float(sum(a.elapsed_time(b) for a, b in ev))
    else:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(steps):
            step()
        e1.record(stream)
        _barrier(ctx)
        total = float(e0.elapsed_time(e1))
    return total, _lib.launch_count() - l0


def _timed_host_calls(ctx, call, steps, warmup=1):
    """End-to-end leg: `call()` takes host buffers and returns a host result; wall clock between barriers (every call
    synchronises internally), plus the default-stream event time as a cross-check."""
    torch = ctx["torch"]
    for _ in range(warmup):
        r = call(); del r
    _barrier(ctx)
    ds = torch.cuda.default_stream()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record(ds)
    chk = 0.0
    for _ in range(steps):
        r = call()
        first = r[0] if isinstance(r, tuple) else r
        chk += float(np.asarray(first).flat[-1].real)   # touch the host result
        del r, first
    e1.record(ds)
    torch.cuda.synchronize()
    wall_ms = 1e3 * (time.perf_counter() - t0)
    ev_ms = float(e0.elapsed_time(e1))
    _barrier(ctx)
    return max(wall_ms, ev_ms), ev_ms


def _roofline(kernel, bytes_per_step, ms_per_step, workload_key, peak, peak_src):
    achieved = bytes_per_step / (ms_per_step * 1e-3) / 1e9
    traffic, note = None, None
    tpath = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(tpath):
        try:
            tj = json.load(open(tpath)).get(workload_key, {})
            traffic = tj.get("dram_bytes_per_launch")
            note = f"dram__bytes_read+write of {tj.get('kernel')} from profiles/{tj.get('source')} (ncu --set full), per launch"
        except Exception:
            pass
    return {"bound": "hbm", "kernel": kernel, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            "traffic": traffic, "traffic_note": note, "peak_source": peak_src,
            "algorithmic_bytes_per_launch": bytes_per_step, "avg_launch_ms": ms_per_step}


# ---------------------------------------------------------------------------------------------------
def bench_cfg3(ctx):
    torch, dist, dev, rank, world, lib, check, _lib = (ctx[k] for k in ("torch", "dist", "dev", "rank", "world", "lib", "check", "_lib"))
    args, sptr = ctx["args"], ctx["sptr"]
    from mrmustard_b200 import backend, sharding, strategies
    gold = golden()
    shape, n_per = CFG3_SHAPE, int(np.prod(CFG3_SHAPE))
    A, b, c = random_triple(2, (CFG3_B,), seed=CFG3_SEED)
    assert sha(np.concatenate([A.ravel(), b.ravel(), c.ravel()])) == str(gold["cfg3_in_sha"]), "bench: cfg3 inputs differ from the golden recipe"
    lo, hi = sharding.shard_range(CFG3_B, rank, world)
    Bl = hi - lo
    hA, hb, hc = (np.ascontiguousarray(x[lo:hi]) for x in (A, b, c))
    dA, db, dc = (torch.from_numpy(x).to(dev) for x in (hA, hb, hc))
    dG = torch.empty((Bl, n_per), dtype=torch.complex128, device=dev)
    sh = _lib.shape_array(shape)

    def fwd():
        check(lib.mmh_forward_batched(Bl, 2, sh, dA.data_ptr(), db.data_ptr(), dc.data_ptr(), dG.data_ptr(), 0, sptr))

    # ---- parity of what is being timed: EVERY rank checks its shard against the reference's per-chunk golden sha256 -------
    fwd(); torch.cuda.synchronize()
    host = dG.cpu().numpy().reshape(Bl, *shape)
    checked = 0
    if lo % 4096 == 0 and hi % 4096 == 0:
        for ci in range(lo // 4096, hi // 4096):
            assert sha(host[(ci * 4096 - lo):(ci + 1) * 4096 - lo]) == str(gold["cfg3_chunk_sha"][ci]), f"bench: rank {rank} chunk {ci} differs from the reference golden"
            checked += 1
    if rank == 0:
        assert np.array_equal(host[:4], gold["cfg3_G_first4"]), "bench: cfg3 result differs from golden"
    del host
    chk = _max_over_ranks(ctx, [-float(checked)])      # min over ranks of the number of checked chunks
    chunks_checked_min = int(-chk[0])

    # ---- VJP inputs and parity ------------------------------------------------------------------------------------------------
    gen = torch.Generator(device=dev); gen.manual_seed(1234 + rank)
    dg = torch.randn((Bl, n_per), dtype=torch.float64, device=dev, generator=gen).to(torch.complex128)
    if rank == 0:   # the first 64 cotangents are the golden recipe's (tests/golden/gen_golden.py)
        dg[:64] = torch.from_numpy(np.random.RandomState(1).standard_normal((64, n_per)) + 0j).to(dev)
    oA = torch.empty((Bl, 2, 2), dtype=torch.complex128, device=dev)
    ob = torch.empty((Bl, 2), dtype=torch.complex128, device=dev)
    oc = torch.empty((Bl,), dtype=torch.complex128, device=dev)

    def vjp():
        check(lib.mmh_vjp_batched(Bl, 2, sh, dG.data_ptr(), dc.data_ptr(), dg.data_ptr(), oA.data_ptr(), ob.data_ptr(), oc.data_ptr(), sptr))

    vjp(); torch.cuda.synchronize()
    # size-independent property on every rank: dLdc = sum(G * dLdG) / c (gradients.py:81)
    want_dc = (dG * dg).sum(dim=1) / dc
    assert bool(torch.all((oc - want_dc).abs() <= 1e-14 + 1e-10 * want_dc.abs())), f"bench: rank {rank} dLdc identity violated"
    if rank == 0:
        for got, name in ((oA, "cfg3_dA64"), (ob, "cfg3_db64"), (oc, "cfg3_dc64")):
            w = gold[name]; g_ = got[:64].cpu().numpy()
            assert np.all(np.abs(g_ - w) <= 1e-14 + 1e-10 * np.abs(w)), f"bench: {name} outside the parity gate"
    del want_dc

    # ---- device-resident timing -------------------------------------------------------------------------------------------------
    sampler = ClockSampler(ctx["local_rank"]); sampler.start()
    fwd_ms, fwd_launches = _timed_steps(ctx, fwd, args.steps, args.warmup)
    vjp_ms, vjp_launches = _timed_steps(ctx, vjp, args.steps, args.warmup)
    clocks = sampler.stop()

    # ---- optional final gather (NCCL all_gather of the shards), timed separately ---------------------------------------------
    gather_ms = None
    if world > 1:
        full = torch.empty((world * Bl, n_per, 2), dtype=torch.float64, device=dev)
        src = torch.view_as_real(dG)
        for _ in range(2):
            dist.all_gather_into_tensor(full, src)
        _barrier(ctx)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(ctx["stream"])
        for _ in range(3):
            dist.all_gather_into_tensor(full, src)
        e1.record(ctx["stream"])
        _barrier(ctx)
        gather_ms = float(e0.elapsed_time(e1)) / 3
        del full

    # ---- end to end through the numpy-facing plugin calls (host buffers, copies inside the timed region) -------------------
    e2e = vjp_e2e = None
    if not args.no_e2e:
        e2e_steps = max(3, min(args.steps, 5))
        ms, ev_ms = _timed_host_calls(ctx, lambda: backend.hermite_renormalized_batched(hA, hb, hc, shape), e2e_steps)
        Gh = backend.hermite_renormalized_batched(hA, hb, hc, shape)                 # page-locked (the call's own result buffer)
        gh = _lib.pinned_empty(Gh.shape); gh[...] = dg.cpu().numpy().reshape(Gh.shape)
        vms, vev_ms = _timed_host_calls(ctx, lambda: strategies.vanilla_batch_vjp_numba(Gh, hc, gh), e2e_steps)
        ms, vms = _max_over_ranks(ctx, [ms, vms])
        e2e = {"value": CFG3_B * n_per * e2e_steps / (ms * 1e-3), "unit": UNIT, "steps": e2e_steps, "ms_per_step": ms / e2e_steps,
               "h2d_bytes_per_step": int(hA.nbytes + hb.nbytes + hc.nbytes), "d2h_bytes_per_step": int(Bl * n_per * 16),
               "bytes_are": "per rank", "timer": "wall clock between barriers (every call synchronises); max over ranks",
               "api": "mrmustard_b200.backend.hermite_renormalized_batched(A, b, c, shape) [numpy in, numpy out] -> mmh_forward_batched_host"}
        vjp_e2e = {"value": CFG3_B * n_per * e2e_steps / (vms * 1e-3), "unit": UNIT, "steps": e2e_steps, "ms_per_step": vms / e2e_steps,
                   "h2d_bytes_per_step": int(2 * Bl * n_per * 16 + hc.nbytes), "d2h_bytes_per_step": int(oA.numel() * 16 + ob.numel() * 16 + oc.numel() * 16),
                   "bytes_are": "per rank", "host_buffers": "G and dLdG in page-locked numpy arrays (mrmustard_b200._lib.pinned_empty)",
                   "api": "mrmustard_b200.strategies.vanilla_batch_vjp_numba(G, c, dLdG) [numpy in, numpy out] -> mmh_vjp_batched_host"}
        del Gh, gh

    fwd_ms, vjp_ms = _max_over_ranks(ctx, [fwd_ms, vjp_ms])
    peak, peak_src = load_peaks()
    total_amps = CFG3_B * n_per
    fwd_step, vjp_step = fwd_ms / args.steps, vjp_ms / args.steps
    line = {
        "metric": METRIC, "value": total_amps / (fwd_step * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": fwd_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic", "config": config_for("cfg3", world),
        "notes": {"l2": "per-rank step output (>= 210 MB) larger than the 126 MB L2; no flush needed",
                  "parity": f"every rank's shard sha256-checked against the reference golden before timing (>= {chunks_checked_min} chunks of 4096 triples per rank); "
                            "VJP: dLdc identity on every rank + reference goldens on rank 0",
                  "triples_per_gpu": Bl},
        "clocks": clocks, "e2e": e2e, "gpu_launches": int(fwd_launches),
        # per-rank view: the roofline of ONE GPU's kernel on its shard
        "roofline": _roofline("mmh_forward_batched = ONE launch of k_march_lanes<5, fused chain> per step (100% of the step)",
                              ALGO_BYTES_PER_AMP_FWD * Bl * n_per, fwd_step, "cfg3", peak, peak_src),
        "vjp": {"metric": "Fock amplitudes/s through the VJP (complex128)", "value": total_amps / (vjp_step * 1e-3), "unit": UNIT,
                "ms_per_step": vjp_step, "gpu_launches": int(vjp_launches),
                "roofline": _roofline("mmh_vjp_batched = ONE launch of k_vjp_lanes<5> per step", ALGO_BYTES_PER_AMP_VJP * Bl * n_per,
                                      vjp_step, "cfg3_vjp", peak, peak_src),
                "e2e": vjp_e2e},
        "forward_plus_vjp": {"value": total_amps / ((fwd_step + vjp_step) * 1e-3), "unit": UNIT, "ms_per_step": fwd_step + vjp_step},
        "gather": None if gather_ms is None else {"ms": gather_ms, "bytes_per_rank": int(Bl * n_per * 16),
                                                  "what": "NCCL all_gather_into_tensor of the lattice shards (not part of `value`)"},
    }
    if rank == 0 and not args.no_cpu:
        arm = CpuArm("cfg3")
        line["cpu_baseline"] = arm.baseline("forward")
        line["vjp"]["cpu_baseline"] = arm.baseline("vjp", 6.0)
    elif rank == 0:
        line["cpu_baseline"] = None
    return line


# ---------------------------------------------------------------------------------------------------
def cfg2_record(ctx, steps, warmup, with_cpu=True, with_e2e=True):
    """One (50,)^4 lattice per GPU per step: device-resident timing with L2 flushed between steps, e2e, roofline, CPU baseline."""
    torch, dev, rank, world, lib, check, _lib = (ctx[k] for k in ("torch", "dev", "rank", "world", "lib", "check", "_lib"))
    sptr = ctx["sptr"]
    from mrmustard_b200 import strategies
    gold = golden()
    if rank == 0:
        A, b, c = gold["cfg2_A"], gold["cfg2_b"], gold["cfg2_c"].reshape(1)
        want = str(gold["cfg2_G50_sha"])
    else:   # other ranks: the raw-kernel variant of cfg2 -- one golden-pinned triple shared by all of them
        A, b, c = gold["cfg2r_A"], gold["cfg2r_b"], gold["cfg2r_c"].reshape(1)
        want = str(gold["cfg2r_G50_sha"])
    A, b, c = (np.ascontiguousarray(x) for x in (A, b, c))
    shape, n = (50,) * 4, 50 ** 4
    dA, db, dc = (torch.from_numpy(x).to(dev) for x in (A, b, c))
    dG = torch.empty(n, dtype=torch.complex128, device=dev)
    sh = _lib.shape_array(shape)

    def fwd():
        check(lib.mmh_forward(4, sh, dA.data_ptr(), db.data_ptr(), dc.data_ptr(), dG.data_ptr(), 0, sptr))

    fwd(); torch.cuda.synchronize()
    assert sha(dG.cpu().numpy()) == want, f"bench: rank {rank} device result differs from the reference's golden sha256"
    sampler = ClockSampler(ctx["local_rank"]); sampler.start()
    ms, launches = _timed_steps(ctx, fwd, steps, warmup, flush_between=True)
    clocks = sampler.stop()
    e2e = None
    if with_e2e:
        e2e_steps = max(3, min(steps, 10))
        ems, ev = _timed_host_calls(ctx, lambda: strategies.vanilla_numba(shape, A, b, complex(c[0])), e2e_steps, warmup=2)
        (ems,) = _max_over_ranks(ctx, [ems])
        e2e = {"value": world * n * e2e_steps / (ems * 1e-3), "unit": UNIT, "steps": e2e_steps, "ms_per_step": ems / e2e_steps,
               "h2d_bytes_per_step": int(A.nbytes + b.nbytes + c.nbytes), "d2h_bytes_per_step": int(n * 16), "bytes_are": "per rank",
               "api": "mrmustard_b200.strategies.vanilla_numba -> mmh_forward_host (page-locked result buffer)"}
    (ms,) = _max_over_ranks(ctx, [ms])
    peak, peak_src = load_peaks()
    step = ms / steps
    rec = {"value": world * n / (step * 1e-3), "unit": UNIT, "ms_per_step": step, "steps": steps, "gpu_launches": int(launches),
           "clocks": clocks, "e2e": e2e,
           "roofline": _roofline("mmh_forward: k_warp_tail + k_march_tiled2 (stage 1) + k_march_rows (stage 0, dominant: 65 % of the serialised time); the kernels overlap",
                                 ALGO_BYTES_PER_AMP_FWD * n, step, "cfg2", peak, peak_src),
           "notes": {"l2": "flushed between timed steps (untimed 256 MiB write; one lattice = 100 MB < 126 MB L2)",
                     "parity": "every rank's device result sha256-checked against the reference golden before timing"}}
    if with_cpu and rank == 0:
        rec["cpu_baseline"] = CpuArm("cfg2").baseline("forward", 6.0)
    return rec


def bench_cfg2(ctx):
    args, world = ctx["args"], ctx["world"]
    rec = cfg2_record(ctx, args.steps, args.warmup, with_cpu=not args.no_cpu, with_e2e=not args.no_e2e)
    line = {"metric": METRIC, "value": rec["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": rec["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": config_for("cfg2", world), "notes": rec["notes"], "clocks": rec["clocks"],
            "e2e": rec["e2e"], "gpu_launches": rec["gpu_launches"], "roofline": rec["roofline"],
            "cpu_baseline": rec.get("cpu_baseline")}
    return line


def bench_cfg4(ctx):
    """ONE (12,)^8 lattice (430 M amplitudes, 6.9 GB): N = 1 through mmh_forward, N > 1 sharded by panel ranges with halo
    exchange between the ranks (mrmustard_b200.sharding.forward_single_sharded)."""
    torch, dev, rank, world, lib, check, _lib = (ctx[k] for k in ("torch", "dev", "rank", "world", "lib", "check", "_lib"))
    args, sptr = ctx["args"], ctx["sptr"]
    from mrmustard_b200 import sharding
    gold = golden()
    A, b, c = (np.ascontiguousarray(x) for x in (gold["cfg4_A"], gold["cfg4_b"], gold["cfg4_c"].reshape(1)))
    shape, n = (12,) * 8, 12 ** 8
    steps = min(args.steps, 10)
    if world == 1:
        dA, db, dc = (torch.from_numpy(x).to(dev) for x in (A, b, c))
        dG = torch.empty(n, dtype=torch.complex128, device=dev)
        sh = _lib.shape_array(shape)
        step = lambda: check(lib.mmh_forward(8, sh, dA.data_ptr(), db.data_ptr(), dc.data_ptr(), dG.data_ptr(), 0, sptr))
        get = lambda: dG
    else:
        plan = sharding.SingleLatticePlan(shape, A, b, complex(c[0]))
        step = plan.run if os.environ.get("MMH_BENCH_NO_GRAPH") else plan.run_graphed
        get = lambda: plan.G
    step(); torch.cuda.synchronize()
    # parity: the (3,)^8 corner of the lattice is the reference golden cfg4_G3 (a lattice's corner does not depend on the cutoff)
    G = get().view(shape)
    # rank 0 owns the panel offsets f < P / world of every panel k_0 >= 1: the part of the corner with k_1 < k1c lies inside
    k1c = 3
    while world > 1 and k1c > 1 and (k1c - 1) * 12 ** 6 + 2 * sum(12 ** j for j in range(6)) >= (12 ** 7) // world:
        k1c -= 1
    lo_ok = True
    if world == 1 or rank == 0:
        sl = (slice(0, 3), slice(0, k1c)) + tuple(slice(0, 3) for _ in range(6))
        lo_ok = bool(np.array_equal(G[sl].cpu().numpy(), gold["cfg4_G3"][sl]))
    assert lo_ok, "bench: cfg4 corner differs from the reference golden"
    # every rank: its last owned amplitude of the last panel must be finite and non-zero (the halo chain delivered)
    if world > 1:
        probe = complex(plan.G[(shape[0] - 1) * plan.P + plan.f_hi - 1].cpu())
        assert np.isfinite(probe.real) and np.isfinite(probe.imag), f"bench: rank {rank} holds a non-finite amplitude"
    sampler = ClockSampler(ctx["local_rank"]); sampler.start()
    ms, launches = _timed_steps(ctx, step, steps, max(args.warmup, 3) if world == 1 else 3)
    clocks = sampler.stop()
    (ms,) = _max_over_ranks(ctx, [ms])
    peak, peak_src = load_peaks()
    st = ms / steps
    line = {"metric": METRIC, "value": n / (st * 1e-3), "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": max(args.warmup, 3),
            "ms_per_step": st, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_for("cfg4", world), "clocks": clocks, "e2e": None, "gpu_launches": int(launches),
            "roofline": _roofline("k_panel_step (one launch per panel step)", ALGO_BYTES_PER_AMP_FWD * n / world, st, "cfg4", peak, peak_src),
            "notes": {"l2": "6.9 GB lattice, larger than L2", "parity": "(3,)^8 corner bit-identical to the reference golden",
                      "graph": None if world == 1 else ("eager" if getattr(plan, "_graph", None) in (None, False) else "CUDA graph replay"),
                      "graph_error": None if world == 1 else getattr(plan, "_graph_error", None)}}
    if rank == 0 and not args.no_cpu:
        line["cpu_baseline"] = CpuArm("cfg4").baseline("forward", 4.0)
    return line


# ---------------------------------------------------------------------------------------------------
def extras(ctx, skip):
    """Secondary records on the other BASELINE configs (N = 1): device-resident, CUDA events, L2 flushed where the output fits L2."""
    torch, dev, lib, check, _lib, stream, sptr, flush = (ctx[k] for k in ("torch", "dev", "lib", "check", "_lib", "stream", "sptr", "flush"))
    out = {}
    peak, _ = load_peaks()
    gold = golden()

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

    def wall(fn, reps=3):
        fn(); best = 1e30
        for _ in range(reps):
            t0 = time.perf_counter(); fn(); best = min(best, time.perf_counter() - t0)
        return best * 1e3

    def guarded(name, fn):
        try:
            fn()
        except Exception as e:  # extras must never break the contract line
            out[name + "_error"] = repr(e)
        torch.cuda.synchronize()

    def cfg2():
        out["cfg2"] = cfg2_record(ctx, 20, 3, with_cpu=not ctx["args"].no_cpu)

    def cfg5():
        # cfg5: 4-mode ket, cutoff 40: forward + vjp (device resident) and the device-resident "fidelity gradient step"
        import mrmustard_b200 as mm
        from mrmustard_b200 import device as dv
        A, b, c = gold["cfg5_A"], gold["cfg5_b"], gold["cfg5_c"].reshape(1)
        dA, db, dc = (torch.from_numpy(np.ascontiguousarray(x)).to(dev) for x in (A, b, c))
        shape = (40,) * 4; sh = _lib.shape_array(shape); n = 40 ** 4
        dG = torch.empty((n,), dtype=torch.complex128, device=dev)
        ms = timeit(lambda: check(lib.mmh_forward(4, sh, dA.data_ptr(), db.data_ptr(), dc.data_ptr(), dG.data_ptr(), 0, sptr)), 10, True)
        assert sha(dG.cpu().numpy()) == str(gold["cfg5_G40_sha"])
        rec = {"forward": {"amps_per_s": n / (ms * 1e-3), "ms": ms, "hbm_frac": 16.0 * n / (ms * 1e-3) / 1e9 / peak}}
        g = torch.randn((n,), dtype=torch.float64, device=dev).to(torch.complex128)
        oA = torch.empty((4, 4), dtype=torch.complex128, device=dev)
        ob = torch.empty((4,), dtype=torch.complex128, device=dev)
        oc = torch.empty((1,), dtype=torch.complex128, device=dev)
        ms = timeit(lambda: check(lib.mmh_vjp(4, sh, dG.data_ptr(), dc.data_ptr(), g.data_ptr(), oA.data_ptr(), ob.data_ptr(), oc.data_ptr(), sptr)), 10, True)
        rec["vjp"] = {"amps_per_s": n / (ms * 1e-3), "ms": ms, "hbm_frac": 32.0 * n / (ms * 1e-3) / 1e9 / peak}
        # train step, device resident (the torch stand-in of the reference's jax custom_vjp inside Optimizer.minimize):
        # host (A, b, c) -> H2D (336 B) -> lattice -> loss = 1 - |<t|G>|^2 -> backward (mmh_vjp) -> D2H of 21 complex gradients + loss
        t = torch.randn(shape, dtype=torch.float64, device=dev).to(torch.complex128); t /= t.norm()
        hA, hb, hc = (torch.from_numpy(np.ascontiguousarray(x)).pin_memory() for x in (A, b, c.reshape(())))

        def train_step():
            pa, pb, pc = (x.to(dev, non_blocking=True).requires_grad_() for x in (hA, hb, hc))
            G = dv.hermite_renormalized(pa, pb, pc, shape)
            ov = torch.sum(t.conj() * G)
            loss = 1.0 - (ov.real ** 2 + ov.imag ** 2)
            loss.backward()
            return torch.cat([pa.grad.reshape(-1), pb.grad, pc.grad.reshape(1), loss.reshape(1).to(torch.complex128)]).cpu()

        ms_e2e = wall(train_step, 10)
        a, bb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        train_step(); torch.cuda.synchronize()
        a.record(stream); train_step(); bb.record(stream); torch.cuda.synchronize()
        rec["train_step_autograd_e2e"] = {"wall_ms": ms_e2e, "device_ms": a.elapsed_time(bb), "h2d_bytes": 336, "d2h_bytes": 22 * 16,
                                          "api": "mrmustard_b200.device.hermite_renormalized (torch.autograd.Function over mmh_forward / mmh_vjp); loss written in torch ops"}
        # the same step through the fused device path: packed H2D (336 B) -> mmh_forward -> mmh_overlap -> mmh_vjp (constant
        # cotangent conj(target)) -> D2H of 22 complex numbers; host inputs, host outputs, everything inside the timed region
        step = dv.FidelityStep(shape, t)
        hAn, hbn, hcn = np.ascontiguousarray(A), np.ascontiguousarray(b), complex(c[0])
        ms_fused = wall(lambda: step(hAn, hbn, hcn), 20)
        kern = rec["forward"]["ms"] + rec["vjp"]["ms"]
        step_eager = dv.FidelityStep(shape, t, use_graph=False)
        ms_eager = wall(lambda: step_eager(hAn, hbn, hcn), 20)
        rec["train_step_e2e"] = {"wall_ms": ms_fused, "wall_ms_without_cuda_graph": ms_eager, "cuda_graph": bool(getattr(step, "_graph", None)),
                                 "kernel_ms_forward_plus_vjp": kern, "ratio_to_kernels": ms_fused / kern,
                                 "steps_per_s": 1e3 / ms_fused, "h2d_bytes": 336, "d2h_bytes": 22 * 16,
                                 "api": "mrmustard_b200.device.FidelityStep(shape, target)(A, b, c) -> (loss, dLdA, dLdb, dLdc); numpy in, numpy out"}
        ms_np = wall(lambda: mm.strategies.vanilla_vjp_numba(mm.strategies.vanilla_numba(shape, A, b, complex(c[0])), complex(c[0]), np.ones(shape, complex)), 3)
        rec["train_step_numpy_api_wall_ms"] = ms_np
        if not ctx["args"].no_cpu:
            arm = CpuArm("cfg5")
            rec["cpu_baseline_forward"] = arm.baseline("forward", 3.0)
            rec["cpu_baseline_vjp"] = arm.baseline("vjp", 3.0)
        out["cfg5"] = rec

    def cfg3():
        A, b, c = random_triple(2, (CFG3_B,), seed=CFG3_SEED)
        dA, db, dc = (torch.from_numpy(x).to(dev) for x in (A, b, c))
        sh = _lib.shape_array(CFG3_SHAPE); n = 1600; B = CFG3_B
        dG = torch.empty((B, n), dtype=torch.complex128, device=dev)
        ms = timeit(lambda: check(lib.mmh_forward_batched(B, 2, sh, dA.data_ptr(), db.data_ptr(), dc.data_ptr(), dG.data_ptr(), 0, sptr)), 5, False)
        out["cfg3_forward"] = {"amps_per_s": B * n / (ms * 1e-3), "ms": ms, "hbm_frac": 16.0 * B * n / (ms * 1e-3) / 1e9 / peak}

    def batches():
        # cfg2 as a batch: 8 x (50,)^4 through hermite_renormalized_batched (consecutive lattices pipelined on the device)
        A, b, c = gold["cfg2_A"], gold["cfg2_b"], gold["cfg2_c"].reshape(1)
        Bq = 8
        dA, db, dc = (torch.from_numpy(np.ascontiguousarray(np.repeat(x[None], Bq, 0))).to(dev) for x in (A, b, c))
        shape = (50,) * 4; sh = _lib.shape_array(shape); n = 50 ** 4
        dG = torch.empty((Bq, n), dtype=torch.complex128, device=dev)
        ms = timeit(lambda: check(lib.mmh_forward_batched(Bq, 4, sh, dA.data_ptr(), db.data_ptr(), dc.data_ptr(), dG.data_ptr(), 0, sptr)), 5, True)
        out["cfg2_batch8_forward"] = {"amps_per_s": Bq * n / (ms * 1e-3), "ms": ms, "hbm_frac": 16.0 * Bq * n / (ms * 1e-3) / 1e9 / peak}
        del dG
        # a batch of 2-mode unitaries at cutoff 20: 148 x (20,)^4 (box march, mmh_box.cu)
        Bq = 148
        A, b, c = random_triple(4, (Bq,), seed=3)
        dA, db, dc = (torch.from_numpy(np.ascontiguousarray(x)).to(dev) for x in (A, b, c))
        shape = (20,) * 4; sh = _lib.shape_array(shape); n = 20 ** 4
        dG = torch.empty((Bq, n), dtype=torch.complex128, device=dev)
        ms = timeit(lambda: check(lib.mmh_forward_batched(Bq, 4, sh, dA.data_ptr(), db.data_ptr(), dc.data_ptr(), dG.data_ptr(), 0, sptr)), 5, True)
        out["batch148_20p4_forward"] = {"amps_per_s": Bq * n / (ms * 1e-3), "ms": ms, "hbm_frac": 16.0 * Bq * n / (ms * 1e-3) / 1e9 / peak}

    def stable():
        A, b, c = gold["cfg2_A"], gold["cfg2_b"], gold["cfg2_c"].reshape(1)
        dA, db, dc = (torch.from_numpy(np.ascontiguousarray(x)).to(dev) for x in (A, b, c))
        shape = (50,) * 4; sh = _lib.shape_array(shape); n = 50 ** 4
        dG = torch.empty((n,), dtype=torch.complex128, device=dev)
        ms = timeit(lambda: check(lib.mmh_forward(4, sh, dA.data_ptr(), db.data_ptr(), dc.data_ptr(), dG.data_ptr(), 1, sptr)), 5, True)
        assert sha(dG.cpu().numpy()) == str(gold["cfg2_G50_stable_sha"])
        out["cfg2_stable_forward"] = {"amps_per_s": n / (ms * 1e-3), "ms": ms, "hbm_frac": 16.0 * n / (ms * 1e-3) / 1e9 / peak}

    def misc():
        import mrmustard_b200 as mm
        # lattice + derived-variable contraction, numpy-facing call: fused vs materialised on the host
        A4, b4, _ = random_triple(4, (), seed=21)
        core, der = (48, 48), (8, 40)
        cpoly = np.random.RandomState(3).standard_normal(der) + 0j
        ms_f = wall(lambda: mm.hermite_renormalized_contracted(A4, b4, cpoly, core))
        ms_m = wall(lambda: np.einsum("abk,k->ab", mm.strategies.vanilla_numba(core + der, A4, b4, 1.0).reshape(core + (-1,)), cpoly.reshape(-1)))
        out["contract_48x48_x_8x40"] = {"fused_ms": ms_f, "materialised_ms": ms_m, "lattice_amplitudes": int(np.prod(core + der))}
        # cfg4: M = 4 diagonal / one-leftover-mode sweeps at cutoff 12 (numpy-facing calls)
        gd = np.load(os.path.join(ROOT, "tests", "golden", "diagonal_golden.npz"))
        Ad, bd, cd = gd["d4_A"], gd["d4_b"], complex(gd["d4_c"])
        out["cfg4_diagonal_M4_cutoff12_ms"] = wall(lambda: mm.hermite_renormalized_diagonal(Ad, bd, cd, (12,) * 4))
        Al, bl, cl = gd["l4_A"], gd["l4_b"], complex(gd["l4_c"])
        out["cfg4_1leftover_M4_cutoff12_ms"] = wall(lambda: mm.hermite_renormalized_1leftoverMode(Al, bl, cl, 11, (11, 11, 11)))
        # cfg1: latency config
        A, b, c = gold["cfg1_A"], gold["cfg1_b"], gold["cfg1_c"].reshape(1)
        dA, db, dc = (torch.from_numpy(np.ascontiguousarray(x)).to(dev) for x in (A, b, c))
        sh = _lib.shape_array((200,))
        dG = torch.empty((200,), dtype=torch.complex128, device=dev)
        ms = timeit(lambda: check(lib.mmh_forward(1, sh, dA.data_ptr(), db.data_ptr(), dc.data_ptr(), dG.data_ptr(), 0, sptr)), 20, False)
        out["cfg1_forward_us"] = ms * 1e3

    def cfg4():
        import mrmustard_b200 as mm
        # the 8-mode ket of cfg4: as one vanilla lattice (12,)^8 (430 M amplitudes), and through the diagonal strategy
        A8, b8, c8 = gold["cfg4_A"], gold["cfg4_b"], gold["cfg4_c"].reshape(1)
        dA, db, dc = (torch.from_numpy(np.ascontiguousarray(x)).to(dev) for x in (A8, b8, c8))
        shape = (12,) * 8; sh = _lib.shape_array(shape); n = 12 ** 8
        dG = torch.empty((n,), dtype=torch.complex128, device=dev)
        ms = timeit(lambda: check(lib.mmh_forward(8, sh, dA.data_ptr(), db.data_ptr(), dc.data_ptr(), dG.data_ptr(), 0, sptr)), 2, False)
        out["cfg4_vanilla_12p8_forward"] = {"amps_per_s": n / (ms * 1e-3), "ms": ms, "hbm_frac": 16.0 * n / (ms * 1e-3) / 1e9 / peak}
        del dG
        A8k, b8k, c8k = gold["cfg4_A"], gold["cfg4_b"], complex(gold["cfg4_c"])
        Adm = np.zeros((16, 16), complex); Adm[:8, :8] = np.conj(A8k); Adm[8:, 8:] = A8k
        bdm = np.concatenate([np.conj(b8k), b8k])
        out["cfg4_diagonal_M8_cutoff6_ms"] = wall(lambda: mm.hermite_renormalized_diagonal(Adm, bdm, abs(c8k) ** 2, (6,) * 8), 2)
        # cfg4 AS WRITTEN: the 8-mode ket through the diagonal strategy at cutoff 12 (A 16x16, 430 M diagonal amplitudes), device
        # resident, rolling weight-level buffers (mmh_diagonal_rolling.cu); the reference cannot run this config (0.94 TB of aux arrays)
        A2, b2 = (np.ascontiguousarray(x) for x in mm.backend.reorder_AB_bargmann(Adm, bdm))
        dA2, db2, dG0 = (torch.from_numpy(np.ascontiguousarray(np.asarray(x, dtype=np.complex128))).to(dev) for x in (A2, b2, np.array([abs(c8k) ** 2])))
        cut = (12,) * 8
        dO = torch.empty((12 ** 8,), dtype=torch.complex128, device=dev)
        ms = timeit(lambda: check(lib.mmh_diagonal(8, _lib.shape_array(cut), dA2.data_ptr(), db2.data_ptr(), 0, dG0.data_ptr(), dO.data_ptr(), sptr)), 2, False)
        out["cfg4_diagonal_M8_cutoff12"] = {"ms": ms, "diagonal_amps_per_s": 12 ** 8 / (ms * 1e-3), "launches_per_call": 90,
                                            "note": "BASELINE config 4 as written; equals |vanilla (12,)^8|^2 within 1e-10 (tests/test_gpu_diagonal.py)"}
        del dO

    todo = {"cfg2": cfg2, "cfg5": cfg5, "cfg3": cfg3, "batches": batches, "stable": stable, "misc": misc, "cfg4": cfg4}
    for name, fn in todo.items():
        if name != skip:
            guarded(name, fn)
    return out


if __name__ == "__main__":
    main()
