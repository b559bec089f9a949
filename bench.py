#!/usr/bin/env python
"""bench.py — block tensordot throughput (BASELINE.json metric "block tensordot TFLOP/s (% FP64 peak)") on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload T1|T2|T2_8K]

A "step" is ONE pass of the hot path over one batch of synthetic input: the contraction
C = A.tensordot(B, {3}, {0}) of two random U(1) block-sparse rank-4 tensors. The default workload is
BASELINE.json configs[1] literally (T1: 41 charge sectors, total bond dim 2048, fp64: 642 block GEMMs, 0.697 GFLOP);
the DMRG-profile D=4096 contraction (T2, 69.5 GFLOP) is timed in the same run and reported under "workloads".

  value  = algorithmic TFLOP/s (sum over matched block pairs of 2mnk / device time), inputs resident in HBM,
           CUDA events on the engine's stream, L2 flushed between steps.
  e2e    = the same metric through the reference-facing C-ABI call with HOST buffers (qtb_tensordot_host: upload A and B
           from pinned host memory, contract, download C), host wall clock around the synchronous call.
  roofline.peak = cuBLAS DGEMM 8192^3 measured live on the same GPU (MEASURED_PEAKS.json carries no fp64 entry).
  cpu_baseline  = the reference's own CPU implementation (oracle/_ref/ref_harness, the unmodified reference sources
           compiled against libtorch) on the box's host cores, same inputs.

  workloads     = side measurements of the same run (N=1): T2, H_eff.psi at D=4096 (three contractions), and two-site
           DMRG sweeps — Heisenberg L=64 at maximum_bond 256 with the compiled reference's own dmrg() timed on the host
           next to it, and the BASELINE.json metric's first half, Heisenberg L=100 at maximum_bond 4096
           (`seconds_last_sweep`; `--dmrg ''` skips the sweeps, which take ~2 of the ~2.5 minutes of a default run).

N>1 (torchrun): the contraction is an independent object per rank (weak scaling, no data-path collective); time is the
max over ranks of the device time. Under "workloads" ONE H_eff.psi at D=4096 is additionally sharded over the ranks by
charge sector with one NCCL allreduce (strong scaling). stdout carries exactly one JSON line.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

WORKLOADS = {"T1": dict(n_sec=41, D=2048, sigma=6.0), "T2": dict(n_sec=15, D=4096, sigma=1.6),
             "T2_8K": dict(n_sec=17, D=8192, sigma=1.8)}
WORKLOAD_DESC = {
    "T1": "configs[1]: tensordot of two random U(1) block-sparse rank-4 tensors [b,s,s,b*] over one bond, 41 charge "
          "sectors, total bond dim 2048, fp64 (642 block GEMMs)",
    "T2": "DMRG-profile tensordot at D=4096: 15 Gaussian charge sectors (sigma 1.6), rank-4 x rank-4 over one bond, fp64",
    "T2_8K": "DMRG-profile tensordot at D=8192: 17 charge sectors (sigma 1.8), fp64",
}


# ----------------------------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """SM clock + throttle reasons sampled through NVML every 2 ms while the timed region runs (nvidia-smi takes
    ~100 ms per query, longer than the whole timed region of a 70 us step)."""

    def __init__(self, gpu_index: int):
        super().__init__(daemon=True)
        self.gpu = gpu_index
        self.samples = []
        self.reasons = set()
        self._halt = threading.Event()
        self.max_mhz = None
        self.err = None

    def run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = int(vis.split(",")[self.gpu]) if vis and vis.split(",")[self.gpu].isdigit() else self.gpu
            h = nv.nvmlDeviceGetHandleByIndex(idx)
            self.max_mhz = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
            names = {"hw_slowdown": nv.nvmlClocksEventReasonHwSlowdown if hasattr(nv, "nvmlClocksEventReasonHwSlowdown") else 0x8,
                     "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}
            while not self._halt.is_set():
                self.samples.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                except Exception:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for n, bit in names.items():
                    if mask & bit:
                        self.reasons.add(n)
                self._halt.wait(0.002)
        except Exception as e:  # keep the bench alive; the record says why there are no samples
            self.err = repr(e)

    def stop(self):
        self._halt.set()
        self.join(timeout=5)
        out = {"sm_mhz": float(np.median(self.samples)) if self.samples else None, "sm_max_mhz": self.max_mhz,
               "reasons": sorted(self.reasons), "samples": len(self.samples)}
        if self.err:
            out["error"] = self.err
        return out


def dump_qtbt(desc, path):
    """QTBT writer (format shared with oracle/ref_harness.cpp); local so that the product arm never imports oracle/"""
    import struct
    keys = sorted(desc["blocks"].keys())
    with open(path, "wb") as f:
        f.write(b"QTBT0001")
        f.write(struct.pack("<3q", len(desc["sec_sizes"]), len(desc["sel"]), len(keys)))
        f.write(np.asarray([len(s) for s in desc["sec_sizes"]], "<i8").tobytes())
        for s in desc["sec_sizes"]:
            f.write(np.asarray(s, "<i8").tobytes())
        for c in desc["cvals"]:
            f.write(np.asarray(c, "<i8").reshape(-1).tobytes())
        f.write(np.asarray(desc["sel"], "<i8").tobytes())
        for k in keys:
            f.write(np.asarray(k, "<i8").tobytes())
        for k in keys:
            f.write(np.asarray(desc["blocks"][k].shape, "<i8").tobytes())
        for k in keys:
            f.write(np.ascontiguousarray(desc["blocks"][k], "<f8").tobytes())


def algorithmic_flops(a, b, dims_a, dims_b):
    free_a = [i for i in range(len(a["sec_sizes"])) if i not in dims_a]
    free_b = [i for i in range(len(b["sec_sizes"])) if i not in dims_b]
    by = {}
    for i, blk in b["blocks"].items():
        by.setdefault(tuple(i[d] for d in dims_b), []).append(
            (int(np.prod([blk.shape[d] for d in free_b])), int(np.prod([blk.shape[d] for d in dims_b]))))
    fl = 0
    for i, blk in a["blocks"].items():
        m = int(np.prod([blk.shape[d] for d in free_a]))
        for n, k in by.get(tuple(i[d] for d in dims_a), []):
            fl += 2 * m * n * k
    return fl


def reference_cpu_time(a, b, dims_a, dims_b, threads, reps, budget_s=30.0):
    """Times the reference's own btensor::tensordot (oracle/_ref/ref_harness, kind 'reference') or, when the compiled
    reference is absent, the numpy restatement (kind 'port'). Returns (kind, cores, [ms per rep])."""
    harness = os.path.join(ROOT, "oracle", "_ref", "ref_harness")
    if os.path.exists(harness):
        with tempfile.TemporaryDirectory() as td:
            pa, pb, pc = (os.path.join(td, n) for n in ("A.qtbt", "B.qtbt", "C.qtbt"))
            dump_qtbt(a, pa)
            dump_qtbt(b, pb)
            cmd = [harness, "tdot", pa, pb, ",".join(map(str, dims_a)), ",".join(map(str, dims_b)), pc,
                   "--reps", str(reps), "--threads", str(threads)]
            env = dict(os.environ, OMP_NUM_THREADS=str(threads), MKL_NUM_THREADS=str(threads))
            out = subprocess.run(cmd, capture_output=True, text=True, timeout=max(600, budget_s * 20), env=env)
            if out.returncode == 0:
                ms = [float(l.split()[1]) for l in out.stdout.splitlines() if l.startswith("TIME_MS")]
                if ms:
                    return "reference", threads, ms
            sys.stderr.write("ref_harness failed, falling back to the numpy port: " + out.stderr[-500:] + "\n")
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import qtb_oracle as orc  # the one place bench.py executes oracle/ (cpu baseline leg only)
    oa, ob = orc.BT(**{k: a[k] for k in ("sec_sizes", "cvals", "sel", "blocks")}), orc.BT(
        **{k: b[k] for k in ("sec_sizes", "cvals", "sel", "blocks")})
    ms = []
    for _ in range(max(1, reps)):
        t0 = time.perf_counter()
        orc.tensordot(oa, ob, dims_a, dims_b)
        ms.append((time.perf_counter() - t0) * 1e3)
    return "port", 1, ms


# ----------------------------------------------------------------------------------------------------------------------
def run_reference(args, rank, world):
    if rank != 0:
        return
    from quantit_b200 import workloads as wl
    cfg = WORKLOADS[args.workload]
    a, b, da, db = wl.tdot_pair(**cfg)
    flops = algorithmic_flops(a, b, da, db)
    threads = os.cpu_count() or 1
    kind, cores, ms = reference_cpu_time(a, b, da, db, threads, args.steps + args.warmup)
    ms = ms[args.warmup:] if len(ms) > args.warmup else ms
    t = float(np.mean(ms))
    val = flops / (t * 1e-3) / 1e12
    line = {"impl": "reference", "metric": "block_tensordot_tflops", "value": val, "unit": "TFLOP/s", "n_gpus": args.gpus,
            "steps": len(ms), "warmup": args.warmup, "ms_per_step": t, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD_DESC[args.workload], "name": args.workload, **cfg, "flops_per_step": flops},
            "details": {"note": "the reference is a single-host CPU implementation: with --gpus N the host runs the same "
                                "contraction back to back, its rate (value) does not depend on N"},
            "cpu_baseline": {"value": val, "unit": "TFLOP/s", "cores": cores, "kind": kind,
                             "sample": f"{len(ms)} full {args.workload} contractions (whole workload, no sub-sampling)"},
            "e2e": {"value": val, "unit": "TFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def measure_dgemm_peak(torch, n=8192, reps=5):
    x = torch.randn(n, n, dtype=torch.float64, device="cuda")
    y = torch.randn(n, n, dtype=torch.float64, device="cuda")
    for _ in range(2):
        z = x @ y
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        z = x @ y
        e1.record()
        e1.synchronize()
        best = min(best, e0.elapsed_time(e1))
    del x, y, z
    torch.cuda.empty_cache()
    return 2.0 * n ** 3 / (best * 1e-3) / 1e12


def time_workload(torch, qb, ctx, name, steps, warmup, flush_buf, dist=None):
    from quantit_b200 import workloads as wl
    cfg = WORKLOADS[name]
    a, b, da, db = wl.tdot_pair(**cfg)
    A, B = qb.BTensor.from_host(**a), qb.BTensor.from_host(**b)
    info = A.tensordot_info(B, da, db)
    stream = torch.cuda.ExternalStream(ctx.stream)
    c0 = ctx.counters()
    for _ in range(warmup):
        C = A.tensordot(B, da, db)
        del C
    ctx.sync()
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    c1 = ctx.counters()
    evs = []
    with torch.cuda.stream(stream):
        for _ in range(steps):
            flush_buf.zero_()  # L2 flush: 256 MiB write on the same stream, outside the event pair
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            C = A.tensordot(B, da, db)  # public API call: plan-cache hit, pool allocation, ONE grouped-GEMM launch
            e1.record(stream)
            evs.append((e0, e1))
            del C
    ctx.sync()
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    c2 = ctx.counters()
    per = [e0.elapsed_time(e1) for e0, e1 in evs]
    ms = float(np.mean(per))
    return {"a": a, "b": b, "da": da, "db": db, "info": info, "ms": ms, "ms_min": float(np.min(per)),
            "launches": c2["kernel_launches"] - c1["kernel_launches"], "cfg": cfg,
            "bytes_in": wl.stored_bytes(a) + wl.stored_bytes(b)}


def time_e2e(torch, qb, ctx, w, steps, warmup):
    """C-ABI call with host buffers: pinned inputs -> device, contraction, result -> pinned host buffer."""
    import ctypes as C
    from quantit_b200 import engine as eng
    fa, fb = eng.flatten_host(w["a"]), eng.flatten_host(w["b"])
    pin = lambda arr: torch.from_numpy(arr).pin_memory()
    ta, tb = pin(fa["data"]), pin(fb["data"])
    da, db = np.asarray(w["da"], np.int64), np.asarray(w["db"], np.int64)
    nob, numel = C.c_int64(), C.c_int64()
    pi, pf = eng._pi, eng._pf
    base = [ctx.h, fa["nc"], None,
            fa["rank"], pi(fa["nsec"]), pi(fa["ss"]), pi(fa["cv"]), pi(fa["sel"]), fa["nb"], pi(fa["idx"]),
            C.cast(ta.data_ptr(), eng.p_f64),
            fb["rank"], pi(fb["nsec"]), pi(fb["ss"]), pi(fb["cv"]), pi(fb["sel"]), fb["nb"], pi(fb["idx"]),
            C.cast(tb.data_ptr(), eng.p_f64),
            len(da), pi(da), pi(db)]
    eng._check(ctx.lib.qtb_tensordot_host(*base, C.byref(nob), C.byref(numel), None, None))
    r_out = fa["rank"] + fb["rank"] - 2 * len(da)
    cidx = np.zeros(nob.value * r_out, np.int64)
    tc = torch.empty(numel.value, dtype=torch.float64).pin_memory()
    call = lambda: eng._check(ctx.lib.qtb_tensordot_host(*base, C.byref(nob), C.byref(numel), pi(cidx),
                                                          C.cast(tc.data_ptr(), eng.p_f64)))
    for _ in range(warmup):
        call()
    t = []
    for _ in range(steps):
        t0 = time.perf_counter()
        call()
        t.append((time.perf_counter() - t0) * 1e3)
    return float(np.mean(t)), int(ta.numel() + tb.numel()) * 8, int(numel.value) * 8



def time_heff(torch, qb, ctx, n_sec, D, sigma, steps, warmup, flush_buf, dist=None):
    """H_eff.psi = L.W.W.R.psi (details::hamil2site_times_state: three block contractions) on the T3 shapes of
    SURVEY.md section 8d. With a sharded context (N > 1) this is ONE H_eff application split over all ranks by charge
    sector of the bra bond + one allreduce (strong scaling)."""
    from quantit_b200 import workloads as wl
    psi, W, L, R = wl.heff_set(n_sec, D, sigma, seed=5)
    bt = lambda d: qb.BTensor.from_host(**d, ctx=ctx)
    Wb = bt(W)
    H2 = Wb.tensordot(Wb, [2], [0]).permute([0, 1, 3, 4, 2, 5])
    p, l, r = bt(psi), bt(L), bt(R)
    stream = torch.cuda.ExternalStream(ctx.stream)
    for _ in range(warmup):
        qb.hamil2site_times_state(p, H2, l, r)
    ctx.sync()
    if dist is not None:
        dist.barrier()
    c1 = ctx.counters()
    evs = []
    with torch.cuda.stream(stream):
        for _ in range(steps):
            flush_buf.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            out = qb.hamil2site_times_state(p, H2, l, r)
            e1.record(stream)
            evs.append((e0, e1))
            del out
    ctx.sync()
    torch.cuda.synchronize()
    c2 = ctx.counters()
    ms = float(np.mean([a.elapsed_time(b) for a, b in evs]))
    flops_rank = (c2["gemm_flops"] - c1["gemm_flops"]) / steps
    if dist is not None:
        t = torch.tensor([ms, flops_rank], dtype=torch.float64, device="cuda")
        tm = t.clone()
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        ms, flops, max_share = float(tm[0].item()), float(t[1].item()), float(tm[1].item())
    else:
        flops, max_share = flops_rank, flops_rank
    return {"ms_per_step": ms, "flops_per_step": int(flops), "value": flops / (ms * 1e-3) / 1e12, "unit": "TFLOP/s",
            "largest_rank_share": max_share / flops if flops else None,
            "launches_per_step": (c2["kernel_launches"] - c1["kernel_launches"]) / steps}


def time_env_left(torch, qb, ctx, n_sec, D, sigma, steps, warmup, flush_buf):
    """compute_left_env (reference dmrg.cpp:424-456: three block contractions E' = A^dag.(E.A).W) on the T3 shapes of SURVEY.md
    section 8d; the site tensor A is the U factor of the two-site tensor's block SVD."""
    from quantit_b200 import workloads as wl
    psi, W, L, R = wl.heff_set(n_sec, D, sigma, seed=5)
    bt = lambda d: qb.BTensor.from_host(**d, ctx=ctx)
    Wb, l = bt(W), bt(L)
    A, _, _ = qb.svd(bt(psi), 2, 1e-12, 4, D)
    stream = torch.cuda.ExternalStream(ctx.stream)
    for _ in range(warmup):
        qb.compute_left_env(Wb, A, l)
    ctx.sync()
    c1 = ctx.counters()
    evs = []
    with torch.cuda.stream(stream):
        for _ in range(steps):
            flush_buf.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            out = qb.compute_left_env(Wb, A, l)
            e1.record(stream)
            evs.append((e0, e1))
            del out
    ctx.sync()
    torch.cuda.synchronize()
    c2 = ctx.counters()
    ms = float(np.mean([a.elapsed_time(b) for a, b in evs]))
    flops = (c2["gemm_flops"] - c1["gemm_flops"]) / steps
    return {"ms_per_step": ms, "flops_per_step": int(flops), "value": flops / (ms * 1e-3) / 1e12, "unit": "TFLOP/s",
            "launches_per_step": (c2["kernel_launches"] - c1["kernel_launches"]) / steps,
            "workload": f"compute_left_env at D={D} (T3 shapes: {n_sec} charge sectors, Heisenberg MPO): 3 block contractions"}


def reference_dmrg_sweeps(L, maxbond, n_sweeps, cutoff, threads):
    """per-sweep wall milliseconds of the compiled reference's own dmrg() on the host CPU (oracle/_ref/ref_harness heis:
    its Heisenberg bMPO + its random bond-4 bMPS, same L / maximum_bond / cutoff, convergence_criterion 0). None when the
    compiled reference is absent."""
    harness = os.path.join(ROOT, "oracle", "_ref", "ref_harness")
    if not os.path.exists(harness):
        return None
    env = dict(os.environ, OMP_NUM_THREADS=str(threads), MKL_NUM_THREADS=str(threads))
    try:
        out = subprocess.run([harness, "heis", str(L), str(maxbond), repr(cutoff), "0", str(n_sweeps), "0"],
                             capture_output=True, text=True, timeout=900, env=env)
    except subprocess.TimeoutExpired:
        return None
    rows = [l.split() for l in out.stdout.splitlines() if l.startswith("SWEEP")]
    if out.returncode != 0 or not rows:
        return None
    return {"sweep_seconds": [float(r[r.index("ms") + 1]) * 1e-3 for r in rows],
            "sweep_mid_bond": [int(r[r.index("mid_bond") + 1]) for r in rows], "threads": threads,
            "kind": "reference (oracle/_ref/ref_harness heis, the reference's own dmrg())"}


def time_dmrg_sweeps(qb, ctx, L, maxbond, n_sweeps, cutoff=1e-20, with_reference=False):
    """two-site DMRG sweeps of the U(1) Heisenberg chain at a saturated bond dimension (BASELINE.json metric, first half:
    "2-site DMRG sweep time"): inputs from quantit_b200.workloads (no reference code involved), exactly n_sweeps sweeps
    (convergence_criterion = 0), wall seconds per sweep as the reference's dmrg_log_sweeptime records them."""
    from quantit_b200 import workloads as wl
    H = [qb.BTensor.from_host(**h, ctx=ctx) for h in wl.heisenberg_mpo(L)]
    psi = [qb.BTensor.from_host(**p_, ctx=ctx) for p_ in wl.random_mps(L, 4, L % 2, seed=0)]
    log = {}
    c0 = ctx.counters()
    E = qb.dmrg(H, psi, qb.dmrg_options(cutoff, 0.0, maxbond, 4, n_sweeps), oc=0, log=log)
    c1 = ctx.counters()
    out = {"workload": f"U(1) Heisenberg S=1/2 chain L={L}, two-site DMRG from a random bond-4 state, maximum_bond {maxbond}, "
                       f"cutoff {cutoff:g}, exactly {n_sweeps} sweeps; the last sweep runs at the saturated bond dimension",
           "L": L, "maximum_bond": maxbond, "cutoff": cutoff, "energy": E, "sweep_seconds": log["seconds"],
           "sweep_mid_bond": log["mid_bond"], "sweep_energy": log["energy"],
           "updates_per_sweep": 2 * (L - 2), "gemm_flops_total": c1["gemm_flops"] - c0["gemm_flops"],
           "kernel_launches_total": c1["kernel_launches"] - c0["kernel_launches"],
           "seconds_last_sweep": log["seconds"][-1], "mid_bond_last_sweep": log["mid_bond"][-1],
           "unit": "s per sweep", "higher_is_better": False}
    if with_reference:
        ref = reference_dmrg_sweeps(L, maxbond, n_sweeps, cutoff, os.cpu_count() or 1)
        if ref is not None:
            ref["seconds_last_sweep"] = ref["sweep_seconds"][-1]
            out["cpu_reference"] = ref
    return out


def time_svd_sweep(torch, qb, ctx, Ds, ref_max_D=1024):
    """T4 of SURVEY.md section 8d: svd(theta, 2, tol = 1e-10, min = 4, max = D) of a Hubbard-profile two-site tensor
    (U(1) x U(1), ~30-45 sectors per bond), D = 512 ... 8192; the compiled reference's svd on the host CPU beside it up to
    ref_max_D. Wall time of the synchronous call (best of the second and third call; all three are listed)."""
    from quantit_b200 import workloads as wl
    out = {}
    harness = os.path.join(ROOT, "oracle", "_ref", "ref_harness")
    for D in Ds:
        th = wl.hubbard_theta(D, np.random.default_rng(7))
        T = qb.BTensor.from_host(**th, ctx=ctx)
        ts = []
        for _ in range(3):
            ctx.sync()
            t0 = time.perf_counter()
            U, d, V = qb.svd(T, 2, 1e-10, 4, D)
            ctx.sync()
            ts.append(time.perf_counter() - t0)
        rec = {"ms": min(ts[1:]) * 1e3, "ms_calls": [t * 1e3 for t in ts], "kept": int(sum(d.structure()[0][0])), "blocks": T.nblocks,
               "bond_sectors": len(th["sec_sizes"][0]), "stored_MB": wl.stored_bytes(th) / 1e6}
        if D <= ref_max_D and os.path.exists(harness):
            with tempfile.TemporaryDirectory() as td:
                pth = os.path.join(td, "theta.qtbt")
                dump_qtbt(th, pth)
                threads = os.cpu_count() or 1
                r = subprocess.run([harness, "svdt", pth, "2", "1e-10", "4", str(D), "2.0", os.path.join(td, "U"),
                                    os.path.join(td, "d"), os.path.join(td, "V"), "--reps", "1", "--threads", str(threads)],
                                   capture_output=True, text=True, timeout=900)
                tm = [float(l.split()[1]) for l in r.stdout.splitlines() if l.startswith("TIME_MS")]
                if r.returncode == 0 and tm:
                    rec["cpu_reference_ms"] = tm[-1]
                    rec["cpu_threads"] = threads
        out[f"D{D}"] = rec
        del T, U, d, V
        ctx.trim_cache()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="T1", choices=list(WORKLOADS))
    ap.add_argument("--no-extra", action="store_true", help="skip the D=4096 side measurement and the CPU baseline")
    ap.add_argument("--dmrg", default="64,256,7;100,4096,7",
                    help="';'-separated L,maximum_bond,sweeps of the DMRG sweep-time measurements ('100,4096,7' is the "
                         "BASELINE.json configs[2] run, ~2.5 minutes; '' skips them). Runs with maximum_bond <= 512 also time "
                         "the compiled reference's dmrg() on the host CPU")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    # stdout carries exactly ONE JSON line: whatever libraries print while the run is in progress (NCCL writes its
    # version banner to stdout from C) goes to stderr; the real stdout is kept aside for the final line
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)

    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: quantit_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod
        dist_mod.init_process_group("nccl", device_id=torch.device("cuda", local))
        dist = dist_mod
    import quantit_b200 as qb
    ctx = qb.Context(local)
    qb.engine._default_ctx = ctx
    flush_buf = None
    with torch.cuda.stream(torch.cuda.ExternalStream(ctx.stream)):
        flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    sampler = ClockSampler(local)
    sampler.start()
    w = time_workload(torch, qb, ctx, args.workload, args.steps, args.warmup, flush_buf, dist)
    clocks = sampler.stop()

    ms = w["ms"]
    if dist is not None:
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    flops = w["info"]["flops"]
    value = world * flops / (ms * 1e-3) / 1e12

    sharded = None
    if dist is not None and not args.no_extra:
        # strong scaling of ONE H_eff.psi at D=4096 over the ranks: row ranges of the bra bond owned per rank through the
        # three contractions, the engine's own NCCL all-gather of the owned row slabs (quantit_b200/sharding.py). The
        # same call unsharded on one GPU is timed in the same run for the efficiency.
        from quantit_b200.sharding import enable_sharding
        single = time_heff(torch, qb, ctx, 15, 4096, 1.6, 5, 3, flush_buf, dist)
        enable_sharding(ctx)
        sharded = time_heff(torch, qb, ctx, 15, 4096, 1.6, 5, 3, flush_buf, dist)
        ctx.set_sharding(0, 1, None)
        flops1 = single["flops_per_step"] / world  # every rank ran the whole contraction in the unsharded pass
        sharded["flops_per_step"] = int(flops1)
        sharded["value"] = flops1 / (sharded["ms_per_step"] * 1e-3) / 1e12
        sharded["single_gpu_ms_per_step"] = single["ms_per_step"]
        sharded["speedup"] = single["ms_per_step"] / sharded["ms_per_step"]
        sharded["efficiency"] = sharded["speedup"] / world
        sharded["scaling"] = "strong"
        sharded["exchange"] = "ncclAllGather of the owned row slabs issued by the engine (dlopen'ed libnccl)" if getattr(
            ctx, "engine_nccl", False) else "allreduce callback (torch.distributed)"

    e2e_ms, h2d, d2h = time_e2e(torch, qb, ctx, w, max(3, min(args.steps, 10)), 2)
    if dist is not None:
        t = torch.tensor([e2e_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = float(t.item())
    e2e_val = world * flops / (e2e_ms * 1e-3) / 1e12

    line = {"metric": "block_tensordot_tflops", "value": value, "unit": "TFLOP/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD_DESC[args.workload], "name": args.workload, **w["cfg"], "flops_per_step": flops},
            "details": {"block_gemms": w["info"]["pairs"], "out_blocks": w["info"]["out_blocks"],
                        "l2": "flushed between steps (256 MiB write outside the per-step CUDA-event pair)",
                        "parallelism": f"{world} independent contraction(s), one per GPU (the headline workload partitions "
                                       "into independent objects: no data-path collective; the sharded strong-scaling "
                                       "measurement of ONE H_eff.psi is under strong_scaling)"},
            "e2e": {"value": e2e_val, "unit": "TFLOP/s", "ms_per_step": e2e_ms, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "api": "qtb_tensordot_host (C ABI, pinned host buffers)"},
            "gpu_launches": w["launches"], "clocks": clocks}

    if rank == 0:
        peak = measure_dgemm_peak(torch)
        line["roofline"] = {"bound": "tensor", "achieved": flops / (ms * 1e-3) / 1e12, "peak": peak, "unit": "TFLOP/s",
                            "frac": flops / (ms * 1e-3) / 1e12 / peak, "traffic": None,
                            "kernel": "grouped_gemm_kernel (fp64 DMMA.8x8x4)",
                            "peak_source": "cuBLAS DGEMM 8192^3 (torch.matmul fp64) measured live in this run, best of 5; "
                                           "MEASURED_PEAKS.json has no fp64 entry; nominal B200 fp64 tensor peak 40 TFLOP/s",
                            }
        try:
            tr = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))[args.workload]
            line["roofline"]["traffic"] = tr["dram_bytes_per_launch"]
            line["roofline"]["traffic_source"] = tr["source"]
        except Exception:
            pass
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
            line["roofline"]["hbm_peak_gbs_measured"] = peaks.get("hbm_gbs")
        except Exception:
            pass
        if sharded is not None:
            sharded["roofline_frac_aggregate"] = sharded["value"] / (world * peak)
            sharded["workload"] = ("ONE H_eff.psi at D=4096 (T3: 15 charge sectors, Heisenberg MPO) sharded over %d GPUs: "
                                   "row ranges of the bra bond owned per rank, one all-gather of the result" % world)
            line["strong_scaling"] = sharded
        if not args.no_extra and world == 1:
            extra = {}
            for name in ["T2"] if args.workload != "T2" else ["T1"]:
                sw = time_workload(torch, qb, ctx, name, max(3, min(args.steps, 10)), 3, flush_buf, None)
                ach = sw["info"]["flops"] / (sw["ms"] * 1e-3) / 1e12
                extra[name] = {"workload": WORKLOAD_DESC[name], "ms_per_step": sw["ms"], "value": ach, "unit": "TFLOP/s",
                               "flops_per_step": sw["info"]["flops"], "block_gemms": sw["info"]["pairs"],
                               "roofline_frac": ach / peak}
            hf = time_heff(torch, qb, ctx, 15, 4096, 1.6, 5, 3, flush_buf, None)
            hf["roofline_frac"] = hf["value"] / peak
            hf["workload"] = "H_eff.psi at D=4096 (T3: 15 charge sectors, Heisenberg MPO): 3 block contractions"
            extra["heff_D4096"] = hf
            try:
                ev = time_env_left(torch, qb, ctx, 15, 4096, 1.6, 5, 3, flush_buf)
                ev["roofline_frac"] = ev["value"] / peak
                extra["env_left_D4096"] = ev
            except Exception as e:  # noqa: BLE001
                extra["env_left_D4096"] = {"error": repr(e)[:300]}
            try:
                extra["svd_sweep_hubbard"] = time_svd_sweep(torch, qb, ctx, [512, 1024, 2048, 4096, 8192])
            except Exception as e:  # noqa: BLE001
                extra["svd_sweep_hubbard"] = {"error": repr(e)[:300]}
            for spec in [x for x in args.dmrg.split(";") if x.strip()]:
                Ld, Dd, nsw = (int(x) for x in spec.split(","))
                try:  # a side measurement must never cost the headline line
                    extra[f"dmrg_sweep_L{Ld}_D{Dd}"] = time_dmrg_sweeps(qb, ctx, Ld, Dd, nsw, with_reference=(Dd <= 512))
                except Exception as e:  # noqa: BLE001
                    extra[f"dmrg_sweep_L{Ld}_D{Dd}"] = {"error": repr(e)[:300]}
                ctx.trim_cache()
            line["workloads"] = extra
            threads = os.cpu_count() or 1
            kind, cores, cms = reference_cpu_time(w["a"], w["b"], w["da"], w["db"], threads, 12)
            cms = cms[2:] if len(cms) > 4 else cms
            cval = flops / (float(np.mean(cms)) * 1e-3) / 1e12
            line["cpu_baseline"] = {"value": cval, "unit": "TFLOP/s", "cores": cores, "kind": kind,
                                    "ms_per_step": float(np.mean(cms)),
                                    "sample": f"{len(cms)} full {args.workload} contractions on the host CPU "
                                              f"({threads} torch threads), same inputs"}
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
