#!/usr/bin/env python
"""Benchmark of the DF-J/K Fock build (BASELINE.json's metric) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c4|c2|c5|c3]
                    [--secondary c2,c5,c3|none] [--impl reference]

One "step" is one complete Fock build F = H + J - (k/2) K on a resident fitted tensor
(mqcb200_build_fock == the reference's build_fock_df).  The headline workload at every N is
BASELINE.json configs[3], the ~200-atom def2-SVP shape `north_star` states its targets on:
C60H122 RKS-B3LYP/def2-SVP, (n, n_occ, naux) = (1450, 241, 6800), 58 GB packed -- it fits one
B200.  configs[1] (c2), configs[4] (c5, two-spin) and configs[2] (c3, fragment farm) ride along in
the same JSON line as the `workloads` block.  For N>1 (torchrun, one rank per GPU) the build is
sharded over the auxiliary index with ONE exchange of [J|K] per build -- total work fixed, "strong".

value  : builds/s with H, D, C already in HBM (mqcb200_build_fock_device), CUDA events on the
         engine's stream, max over ranks.
e2e    : builds/s through mqcb200_build_fock with pinned HOST buffers (H2D of H, D, C and D2H of F
         inside the timed region); `e2e_pageable` is the same call on pageable NumPy arrays (what
         a Fortran allocatable is).
roofline / kernels : per-kernel device time from CUDA events the engine records on its own
         streams; the headline loop runs the Coulomb kernels concurrently with the exchange
         kernels, so the per-kernel figures come from a second, serial (overlap off) loop.
parity : at EVERY N -- the (sharded) build of the first 64 auxiliary functions against the CPU
         oracle on rank 0, and bit-identity of the full-size F across the ranks.
cpu_baseline : C restatement of build_fock_df on OpenBLAS dgemm (oracle/df_fock_blas.c), 1 thread
         and all host threads, on a bounded auxiliary sub-sample (J and K are sums over naux).
--impl reference : that same CPU restatement as the reference arm (the Fortran reference cannot
         be built in this image), all host threads, one bounded sample per step.
"""
from __future__ import annotations

import argparse
import hashlib
import json
import os
import subprocess
import sys
import threading
import time

# torch.distributed.run exports OMP_NUM_THREADS=1 to its workers, and OpenBLAS sizes its thread
# pool from the environment when it is LOADED: the CPU arm must see every host thread, so the
# variables are put right before NumPy/SciPy come in (the thread count actually used is then set
# explicitly per measurement and stated in the line).
if "reference" in sys.argv[1:]:
    try:
        _n_host = len(os.sched_getaffinity(0))
    except AttributeError:
        _n_host = os.cpu_count() or 1
    for _v in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[_v] = str(_n_host)

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SEED = 20261018
PARITY_NAUX = 64
METRIC = "DF-J/K Fock builds/sec"


# The contract is ONE JSON line on stdout.  Libraries (NCCL's version banner, torchrun
# notices) write to file descriptor 1 behind Python's back, so everything else is sent
# to stderr and only emit() writes to the real stdout.
_REAL_STDOUT = os.fdopen(os.dup(1), "w")
os.dup2(2, 1)
sys.stdout = sys.stderr


def emit(line: dict) -> None:
    _REAL_STDOUT.write(json.dumps(line) + "\n")
    _REAL_STDOUT.flush()


def log(*a):
    print("[bench]", *a, file=sys.stderr, flush=True)


def _env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


def _peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return float(p.get("hbm_gbs", 6650.0)), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled DURING the timed region."""

    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.gpu_index = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.gpu_index}", f"--query-gpu={self.QUERY}",
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, smax, power, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1])); smax.append(float(parts[2])); power.append(float(parts[3]))
            except ValueError:
                continue
            for nm, val in zip(names, parts[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        # "under load" = samples in the upper half of the power range seen
        thr = 0.5 * (min(power) + max(power))
        loaded = [s for s, p in zip(sm, power) if p >= thr] or sm
        return {"sm_mhz": float(np.median(loaded)), "sm_max_mhz": max(smax), "power_w_max": max(power),
                "samples": len(sm), "reasons": sorted(reasons)}


def _workload_cfg(name):
    from metalquicha_b200 import synth
    if name not in synth.CONFIGS:
        raise SystemExit(f"unknown workload {name}; choose from {sorted(synth.CONFIGS)}")
    cfg = dict(synth.CONFIGS[name])
    cfg["two_spin"] = "n_alpha" in cfg
    if cfg["two_spin"]:
        cfg["n_occ"] = cfg["n_alpha"]
    cfg.setdefault("n_beta", 0)
    cfg["k_scale"] = 0.2 if name == "c4" else 1.0      # B3LYP keeps 20 % exact exchange
    cfg["name"] = name
    return cfg


def _config(name, cfg, world, workers=None):
    if name == "c3":
        return {"workload": f"c3: (H2O)64 MBE-3 def2-SVP fragment farm, {256 * world}-fragment batch per step "
                            "(1 monomer : 12 dimers : 243 trimers, 256 per GPU), 12 builds per fragment",
                "n_ao": cfg["n"], "n_occ": cfg["n_occ"], "naux": cfg["naux"],
                "parallelism": f"fragment FIFO over {world} GPU(s) x {workers} worker(s) per GPU, no collective",
                "l2": "fragments are L2-resident by nature (7 MB packed); timed as dispatched, no flush",
                "timing": "host wall clock bracketed by barrier + cudaDeviceSynchronize, max over ranks"}
    nt = (cfg["n"] + 15) // 16
    packed_gb = 8.0 * nt * (nt + 1) / 2 * 256 * cfg["naux"] / 1e9
    return {"workload": f"{name}: {cfg['what']}", "n_ao": cfg["n"], "n_occ": cfg.get("n_occ"),
            "n_beta": cfg.get("n_beta") or None, "naux": cfg["naux"], "k_scale": cfg["k_scale"],
            "parallelism": f"aux-sharded x{world}, one exchange of [J|K] per build" if world > 1 else "single GPU",
            "l2": f"packed tensor {packed_gb:.2f} GB ({packed_gb / world:.2f} GB per GPU) >> 126 MB L2: inputs larger than L2, no flush",
            "timing": "CUDA events on the engine stream, max over ranks"}


# ------------------------------------------------------------------------------------------
# CPU legs: the C/BLAS restatement of the reference's build_fock_df on a bounded sub-sample
# ------------------------------------------------------------------------------------------
def _cpu_sample_size(cfg, override=None):
    if override:
        return int(min(cfg["naux"], override))
    # ~4e11 reference-count flops per sampled build (c4: 197 of 6800 auxiliary functions): about a second on
    # all threads and five on one, so the baseline block costs ~10 s and the reference arm under a minute
    per_aux = 4.0 * cfg["n"] ** 2 * max(cfg["n_occ"] + cfg["n_beta"], 1)
    return int(min(cfg["naux"], max(16, min(PARITY_NAUX * 8, 4.0e11 / per_aux))))


def _cpu_problem(cfg, qs):
    from metalquicha_b200 import synth
    n, n_occ, naux, n_beta = cfg["n"], cfg["n_occ"], cfg["naux"], cfg["n_beta"]
    b = synth.synth_tensor(SEED, n, naux, synth.default_scale(n, naux), 0, qs)
    _, h, density, coeff = synth.synth_problem(SEED, n, n_occ, naux, with_tensor=False)
    coeff_b = synth.synth_orbitals(SEED + 1, n, n_beta) if cfg["two_spin"] else None
    return b, h, density, coeff, coeff_b


def _cpu_build(blas_port, cfg, prob):
    b, h, density, coeff, coeff_b = prob
    if cfg["two_spin"]:
        da, db = coeff @ coeff.T, coeff_b @ coeff_b.T
        return blas_port.build_fock_df_uhf(h, b, da, db, coeff, cfg["n_occ"], coeff_b, cfg["n_beta"],
                                           k_scale=cfg["k_scale"])[0]
    return blas_port.build_fock_df(h, b, density, coeff, cfg["n_occ"], k_scale=cfg["k_scale"])


def _time_cpu(blas_port, cfg, prob, threads, reps):
    ts = []
    with blas_port.blas_threads(threads):
        for _ in range(reps):
            t0 = time.perf_counter()
            _cpu_build(blas_port, cfg, prob)
            ts.append(time.perf_counter() - t0)
    return ts


def run_reference(args, cfg):
    """Reference arm: the CPU restatement of build_fock_df (the Fortran reference cannot be built
    here) with every host thread the BLAS can use; each step is one build on a bounded auxiliary
    sub-sample of the workload, `value` scales it to the full auxiliary range (J and K are sums
    over naux).  `ms_per_step` is the MEASURED time of a step, so steps x ms_per_step is what the
    run really spent."""
    if _env_int("RANK", 0) != 0:
        return
    from oracle import df_fock_blas as blas_port          # the oracle leg: allowed here (tier rule 4)
    qs = _cpu_sample_size(cfg, args.cpu_sample)
    naux = cfg["naux"]
    threads = blas_port.host_threads()
    prob = _cpu_problem(cfg, qs)
    ts = _time_cpu(blas_port, cfg, prob, threads, args.warmup + args.steps)[args.warmup:]
    t_sample = sum(ts) / len(ts)
    t_full = t_sample * naux / qs
    t1 = min(_time_cpu(blas_port, cfg, prob, 1, 2))
    value = 1.0 / t_full
    sample = (f"{qs} of {naux} auxiliary functions per step ({1e3 * t_sample:.1f} ms measured), scaled by {naux}/{qs} "
              f"(J and K are sums over naux); C restatement of build_fock_df on OpenBLAS dgemm, {threads} BLAS threads "
              f"(J loops serial as in the reference)")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "builds/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_sample,
        "ms_per_build_extrapolated": 1e3 * t_full, "sample_fraction": qs / naux,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": _config(cfg["name"], cfg, args.gpus),
        "cpu_baseline": {"value": value, "unit": "builds/s", "cores": threads, "kind": "port", "sample": sample,
                         "value_1thread": 1.0 / (t1 * naux / qs),
                         "note_1thread": "same sample, BLAS pinned to 1 thread: the reference's default "
                                         "sequential-BLAS build (CMakeLists.txt:38-46), best of 2"},
        "e2e": {"value": value, "unit": "builds/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


def _ncu_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed
    `ncu --set full` captures (profiles/ncu_traffic.json), keyed by workload and kernel."""
    path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f)
    return {}


class Ctx:
    pass


def _fp64_peak(torch):
    """Live FP64 peak: cuBLAS DGEMM 8192^3 (its B200 kernel is DMMA.8x8x4 too), best of 5."""
    torch.cuda.synchronize()
    a = torch.randn(8192, 8192, dtype=torch.float64, device="cuda")
    b = torch.randn(8192, 8192, dtype=torch.float64, device="cuda")
    best = 1e9
    for i in range(6):
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record(); a @ b; s1.record(); torch.cuda.synchronize()
        if i:
            best = min(best, s0.elapsed_time(s1))
    del a, b
    torch.cuda.empty_cache()
    return 2 * 8192 ** 3 / best * 1e-9


# ------------------------------------------------------------------------------------------
# whole-molecule build workloads (c2, c4, c5)
# ------------------------------------------------------------------------------------------
def run_builds(ctx, name, primary):
    torch, dist, eng, args = ctx.torch, ctx.dist, ctx.eng, ctx.args
    world, rank = ctx.world, ctx.rank
    from metalquicha_b200 import synth
    from metalquicha_b200.engine import _check, _ptr
    from ctypes import c_double

    cfg = _workload_cfg(name)
    n, n_occ, naux, n_beta, two_spin, k_scale = cfg["n"], cfg["n_occ"], cfg["naux"], cfg["n_beta"], cfg["two_spin"], cfg["k_scale"]
    if args.k_scale is not None and primary:
        k_scale = cfg["k_scale"] = args.k_scale
    steps = args.steps if primary else max(3, min(args.steps, 20))
    scale = synth.default_scale(n, naux)
    _, h, density, coeff = synth.synth_problem(SEED, n, n_occ, naux, with_tensor=False)
    coeff_b = synth.synth_orbitals(SEED + 1, n, n_beta) if two_spin else None
    if two_spin:
        # D_total = C_a C_a^T + C_b C_b^T (input generation; src/scf/mqc_scf_common.f90:98-109)
        density = np.asfortranarray(coeff @ coeff.T + coeff_b @ coeff_b.T)

    q_begin, q_count = synth.shard_range(naux, world, rank)
    t0 = time.perf_counter()
    eng.synth_tensor(n, naux, SEED, scale, q_begin=q_begin, q_count=q_count)
    t_synth = time.perf_counter() - t0
    log(f"{name}: tensor resident ({eng.tensor_bytes() / 1e9:.2f} GB on this rank) in {t_synth:.2f} s")

    def dev(a):      # column-major matrix -> CUDA tensor whose memory is that column-major image
        return torch.from_numpy(np.ascontiguousarray(np.asarray(a).T)).cuda()

    def pinned(a):   # same, in pinned host memory; returns (tensor, column-major numpy view)
        t = torch.from_numpy(np.ascontiguousarray(np.asarray(a).T)).pin_memory()
        return t, t.numpy().T

    d_h, d_d, d_c = dev(h), dev(density), dev(coeff)
    d_f = torch.empty_like(d_h)
    keep = [pinned(h), pinned(density), pinned(coeff), pinned(np.zeros((n, n)))]
    h_np, d_np, c_np, f_np = (k[1] for k in keep)
    pg = [np.asfortranarray(np.array(x, copy=True)) for x in (h, density, coeff)]     # pageable twins
    f_pg = np.empty((n, n), order="F")
    if two_spin:
        d_cb, d_fb = dev(coeff_b), torch.empty_like(d_h)
        keep += [pinned(coeff_b), pinned(np.zeros((n, n)))]
        cb_np, fb_np = keep[4][1], keep[5][1]
        cb_pg, fb_pg = np.asfortranarray(np.array(coeff_b, copy=True)), np.empty((n, n), order="F")

        def dev_step(sync, ks=k_scale):
            eng.build_fock_uhf_device(d_h, d_d, d_c, n_occ, d_cb, n_beta, d_f, d_fb, k_scale=ks, sync=sync)

        def host_step(hh, dd, cc, cb, fa, fb):
            _check(eng._lib.mqcb200_build_fock_uhf(eng.handle, 0, _ptr(hh), _ptr(dd), _ptr(cc), n, n_occ,
                                                   _ptr(cb), n, n_beta, c_double(k_scale), _ptr(fa), _ptr(fb)))

        def e2e_step():
            host_step(h_np, d_np, c_np, cb_np, f_np, fb_np)

        def e2e_pageable_step():
            host_step(pg[0], pg[1], pg[2], cb_pg, f_pg, fb_pg)
        h2d = 8 * (2 * n * n + n * (n_occ + n_beta))
        d2h = 16 * n * n
    else:
        def dev_step(sync, ks=k_scale):
            eng.build_fock_device(d_h, d_d, d_c, n_occ, d_f, k_scale=ks, sync=sync)

        def e2e_step():
            eng.build_fock_df(h_np, d_np, c_np, n_occ, k_scale=k_scale, out=f_np)

        def e2e_pageable_step():
            eng.build_fock_df(pg[0], pg[1], pg[2], n_occ, k_scale=k_scale, out=f_pg)
        h2d = 8 * (2 * n * n + n * n_occ)
        d2h = 8 * n * n

    stream = torch.cuda.ExternalStream(eng.stream())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world > 1:
            t = torch.tensor([x], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        return x

    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def timed_device_loop(k, profile):
        if profile:
            eng.set_profiling(True)
            eng.last_timings()
        barrier()
        e0.record(stream)
        for _ in range(k):
            dev_step(False)
        e1.record(stream)
        e1.synchronize()
        barrier()
        ms = max_over_ranks(e0.elapsed_time(e1)) / k
        ph = None
        if profile:
            ph = {kk: v / k for kk, v in eng.last_timings().items()}
            eng.set_profiling(False)
        return ms, ph

    # ---- device-resident arm (the headline): default engine settings ------------------------
    for _ in range(args.warmup):
        dev_step(True)
    launches_per_build = eng.last_launches()
    sampler = ClockSampler(ctx.local_rank)
    if rank == 0 and primary:
        sampler.start()
    ms_per_step, phase_overlapped = timed_device_loop(steps, True)
    gamma_fused = eng.last_gamma_fused()
    value = 1e3 / ms_per_step
    fock_dev = d_f.cpu().numpy().T.copy()
    log(f"{name}: device-resident {value:.3f} builds/s ({ms_per_step:.3f} ms)")

    # ---- same loop with the Coulomb kernels serialised behind the exchange kernels: per-kernel times
    k_serial = max(2, min(steps, 10))
    eng.set_overlap(False)
    dev_step(True)
    ms_serial, phase = timed_device_loop(k_serial, True)
    eng.set_overlap(True)
    if not np.array_equal(d_f.cpu().numpy().T, fock_dev):
        raise SystemExit(f"{name}: the two-stream and the one-stream build differ bit-wise")

    # ---- the Coulomb kernels alone (k_scale = 0: both passes over B, nothing else on the GPU)
    dev_step(True, 0.0)
    eng.set_profiling(True); eng.last_timings()
    for _ in range(3):
        dev_step(False, 0.0)
    ph_j = {kk: v / 3 for kk, v in eng.last_timings().items()}
    eng.set_profiling(False)
    dev_step(True)

    # ---- end-to-end arms: host buffers through the reference-facing call ----------------------
    def timed_host_loop(step_fn, k):
        for _ in range(max(1, min(args.warmup, 3))):
            step_fn()
        barrier()
        t0_ = time.perf_counter()
        e0.record(stream)
        for _ in range(k):
            step_fn()
        e1.record(stream)
        e1.synchronize()
        wall = 1e3 * (time.perf_counter() - t0_)
        barrier()
        return max_over_ranks(max(e0.elapsed_time(e1), wall)) / k
    e2e_ms = timed_host_loop(e2e_step, steps)
    clocks = sampler.stop() if (rank == 0 and primary) else None
    e2e_equal = bool(np.array_equal(np.asarray(f_np), fock_dev))
    e2e_pg_ms = timed_host_loop(e2e_pageable_step, max(2, min(steps, 10)))
    e2e_pg_equal = bool(np.array_equal(f_pg, fock_dev))

    # ---- bit-identity of the full-size F across ranks + a fingerprint the N=1/2/4/8 lines can be compared on
    digest = hashlib.sha256(np.ascontiguousarray(fock_dev).tobytes()).hexdigest()
    same_across_ranks = True
    if world > 1:
        box = [None] * world
        dist.all_gather_object(box, digest)
        same_across_ranks = all(x == box[0] for x in box)
    fingerprint = {"F[0,0]": float(fock_dev[0, 0]), "F[n/2,n/3]": float(fock_dev[n // 2, n // 3]),
                   "F[n-1,n-1]": float(fock_dev[n - 1, n - 1]), "sum": float(np.sum(fock_dev)),
                   "sum_sq": float(np.sum(fock_dev * fock_dev)),
                   "note": "full-size F of the timed build; agrees across the N=1/2/4/8 lines to ~1e-12 (summation order)"}

    # ---- parity at every N: the first PARITY_NAUX auxiliary functions, sharded like the real build,
    # through the same exchange, against the CPU oracle on rank 0 (the other slot holds them)
    qs = min(PARITY_NAUX, naux)
    pq0, pqc = synth.shard_range(qs, world, rank)
    eng.synth_tensor(n, qs, SEED, scale, q_begin=pq0, q_count=pqc, slot=1)
    if two_spin:
        fa = np.empty((n, n), order="F"); fb = np.empty((n, n), order="F")
        _check(eng._lib.mqcb200_build_fock_uhf(eng.handle, 1, _ptr(pg[0]), _ptr(pg[1]), _ptr(pg[2]), n, n_occ,
                                               _ptr(cb_pg), n, n_beta, c_double(k_scale), _ptr(fa), _ptr(fb)))
        f_par = np.concatenate([fa, fb])
    else:
        f_par = eng.build_fock_df(pg[0], pg[1], pg[2], n_occ, k_scale=k_scale, slot=1)
    eng.clear_tensor(1)
    par_digest = hashlib.sha256(np.ascontiguousarray(f_par).tobytes()).hexdigest()
    par_same = True
    if world > 1:
        box = [None] * world
        dist.all_gather_object(box, par_digest)
        par_same = all(x == box[0] for x in box)
    parity, cpu_baseline = None, None
    if rank == 0 and not args.no_parity:
        from oracle import df_fock_oracle as oracle          # the checker: parity + CPU baseline only
        t0 = time.perf_counter()
        b_s = synth.synth_tensor(SEED, n, naux, scale, q_begin=0, q_count=qs)
        if two_spin:
            j_ref, ka_ref, _ = oracle.jk_df_fast(b_s, density, coeff, n_occ, rhf_factor=1.0)
            _, kb_ref, _ = oracle.jk_df_fast(b_s, density, coeff_b, n_beta, rhf_factor=1.0)
            f_ref = np.concatenate([h + j_ref - k_scale * ka_ref, h + j_ref - k_scale * kb_ref])
        else:
            j_ref, k_ref, _ = oracle.jk_df_fast(b_s, density, coeff, n_occ)
            f_ref = h + j_ref - 0.5 * k_scale * k_ref
        parity = {"max_abs_err_vs_oracle": float(np.max(np.abs(f_par - f_ref))), "tolerance": 1e-10,
                  "sample_naux": qs, "sharded_over": world, "oracle_s": time.perf_counter() - t0,
                  "sample_F_bit_identical_across_ranks": par_same,
                  "full_F_bit_identical_across_ranks": same_across_ranks,
                  "two_stream_equals_one_stream": True, "e2e_equals_device_arm": e2e_equal,
                  "what": f"F of the first {qs} auxiliary functions built through the same sharded path "
                          f"(exchange included) vs oracle.jk_df_fast on rank 0"}
        parity["ok"] = bool(parity["max_abs_err_vs_oracle"] <= 1e-10 and par_same and same_across_ranks)
        log(f"{name}: parity max|F - F_oracle| = {parity['max_abs_err_vs_oracle']:.3e} on {qs} aux, "
            f"ranks identical: {par_same and same_across_ranks}")
        if primary and world == 1 and not args.no_cpu_baseline:
            from oracle import df_fock_blas as blas_port
            qc_ = _cpu_sample_size(cfg, args.cpu_sample)
            prob = (b_s[:, :qc_] if qc_ <= qs else synth.synth_tensor(SEED, n, naux, scale, 0, qc_), h, density, coeff, coeff_b)
            threads = blas_port.host_threads()
            t_all = min(_time_cpu(blas_port, cfg, prob, threads, 2))
            t_one = min(_time_cpu(blas_port, cfg, prob, 1, 1))
            f_cpu = _cpu_build(blas_port, cfg, (b_s[:, :min(qs, qc_)],) + prob[1:])
            cpu_baseline = {
                "value": 1.0 / (t_all * naux / qc_), "unit": "builds/s", "cores": threads, "kind": "port",
                "value_1thread": 1.0 / (t_one * naux / qc_),
                "sample": f"{qc_} of {naux} auxiliary functions ({t_all:.2f} s on {threads} threads, {t_one:.2f} s on 1), "
                          f"scaled by {naux}/{qc_}; C restatement of build_fock_df (oracle/df_fock_blas.c) on OpenBLAS dgemm",
                "sample_ms": {"all_threads": 1e3 * t_all, "one_thread": 1e3 * t_one},
                "agrees_with_numpy_oracle": bool(qc_ < qs or np.max(np.abs(
                    f_cpu - (f_ref[:n] if two_spin else f_ref))) <= 1e-10)}

    # ---- rooflines (from the serial loop) -----------------------------------------------------
    hbm_peak, hbm_src = _peaks()
    npair = n * (n + 1) // 2
    occ_total = n_occ + n_beta
    t_k1, t_k2 = phase["k_half_transform"] * 1e-3, phase["k_accumulate"] * 1e-3
    t_j1, t_j2 = phase["j_gamma"] * 1e-3, phase["j_accumulate"] * 1e-3
    fl_k1 = 2.0 * n * n * occ_total * q_count        # half-transform (all spins), this rank's shard
    fl_k2 = 1.0 * n * n * occ_total * q_count        # SYRK-form accumulation
    by_j = 8.0 * npair * q_count                     # one pass over the packed shard
    fp64_peak = ctx.fp64_peak
    traffic = _ncu_traffic().get(name, {})

    def kern(kname, bound, work, t, peak, unit, scale_):
        ach = work / t * scale_ if t > 0 else 0.0
        return {"kernel": kname, "bound": bound, "achieved": ach, "peak": peak, "unit": unit,
                "frac": ach / peak if peak else None, "ms_per_build": t * 1e3, "traffic": traffic.get(kname)}
    kernels = [
        kern("k_half_transform_kernel", "tensor", fl_k1, t_k1, fp64_peak, "TFLOP/s", 1e-12),
        kern("k_accumulate_kernel", "tensor", fl_k2, t_k2, fp64_peak, "TFLOP/s", 1e-12),
        kern("j_gamma_kernel", "hbm", by_j, ph_j["j_gamma"] * 1e-3, hbm_peak, "GB/s", 1e-9),
        kern("j_accumulate_kernel", "hbm", by_j, ph_j["j_accumulate"] * 1e-3, hbm_peak, "GB/s", 1e-9),
    ]
    kernels[2]["note"] = kernels[3]["note"] = "timed in J-only builds (k_scale = 0): both passes over B, nothing else running"
    roofline = dict(kernels[0])
    roofline["peak_source"] = "measured live: cuBLAS DGEMM 8192^3 best of 5 (FP64 is not in MEASURED_PEAKS.json)"
    k_total = t_k1 + t_k2
    j_alone = (ph_j["j_gamma"] + ph_j["j_accumulate"]) * 1e-3
    summary = {
        "K_tflops": (fl_k1 + fl_k2) / k_total * 1e-12 if k_total > 0 else 0.0,
        "K_frac_of_fp64_peak": (fl_k1 + fl_k2) / k_total * 1e-12 / fp64_peak if k_total > 0 else 0.0,
        "J_gbs": 2 * by_j / j_alone * 1e-9 if j_alone > 0 else 0.0,
        "J_frac_of_hbm_peak": 2 * by_j / j_alone * 1e-9 / hbm_peak if j_alone > 0 else 0.0,
        "J_note": "two passes over the packed tensor (gamma, then J), J-only builds",
        "gamma_fused_in_timed_builds": bool(gamma_fused),
        "J_in_build_ms": {"serial": (t_j1 + t_j2) * 1e3,
                          "note": "inside a full build gamma comes out of the half-transform's epilogue when D = f*C*C^T, "
                                  "so only ONE pass over B remains; with two streams it runs beside the exchange kernels"},
        "J_in_build_gbs_one_pass": by_j / t_j2 * 1e-9 if (gamma_fused and t_j2 > 0) else None,
        "whole_build_frac_of_K_at_fp64_peak": ((fl_k1 + fl_k2) / (fp64_peak * 1e12)) / (ms_per_step * 1e-3),
        "fp64_peak_tflops": fp64_peak, "fp64_issue_bound_tflops": 37.05,
        "hbm_peak_gbs": hbm_peak, "hbm_peak_source": hbm_src,
        "ms_per_build": {"two_streams": ms_per_step, "one_stream": ms_serial},
        "phase_ms_per_build_one_stream": phase, "phase_ms_per_build_two_streams": phase_overlapped,
        "reference_flops_as_executed_tflops": 4.0 * n * n * occ_total * q_count / k_total * 1e-12 if k_total > 0 else 0.0,
        "packed_tensor_gb_per_rank": eng.tensor_bytes() / 1e9,
    }
    out = {
        "value": value, "unit": "builds/s", "ms_per_step": ms_per_step, "steps": steps,
        "config": _config(name, cfg, world),
        "e2e": {"value": 1e3 / e2e_ms, "unit": "builds/s", "ms_per_step": e2e_ms,
                "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "host_buffers": "pinned",
                "bit_identical_to_device_arm": e2e_equal},
        "e2e_pageable": {"value": 1e3 / e2e_pg_ms, "unit": "builds/s", "ms_per_step": e2e_pg_ms,
                         "host_buffers": "pageable NumPy arrays (what a Fortran allocatable is)",
                         "bit_identical_to_device_arm": e2e_pg_equal},
        "gpu_launches": launches_per_build * steps, "launches_per_build": launches_per_build,
        "roofline": roofline, "kernels": kernels, "summary": summary,
        "parity": parity, "cpu_baseline": cpu_baseline, "fock_fingerprint": fingerprint,
        "clocks": clocks, "tensor_setup_s": t_synth,
    }
    return out


# ------------------------------------------------------------------------------------------
# separately timed: host tensor -> packed upload, and the whitening GEMM
# ------------------------------------------------------------------------------------------
def run_setup_timings(ctx, name):
    """set_tensor from a host full-square (n*n, q) array (lower triangles over PCIe, double-
    buffered) on a ~1 GiB sample of slabs, and build_df_tensor's whitening GEMM on the device."""
    eng = ctx.eng
    from metalquicha_b200 import synth
    cfg = _workload_cfg(name)
    n, naux = cfg["n"], cfg["naux"]
    out = {}
    q_up = int(max(1, min(naux, (1 << 30) // (8 * n * n))))
    rng = np.random.default_rng(1)
    slab = rng.standard_normal((n, n)); slab = 0.5 * (slab + slab.T)
    b_host = np.empty((n * n, q_up), order="F")
    b_host[:] = slab.reshape(n * n, order="F")[:, None]
    eng.set_tensor_shard(b_host, n, naux, 0, slot=1)          # warm (allocations, pinning)
    t0 = time.perf_counter()
    eng.set_tensor_shard(b_host, n, naux, 0, slot=1)
    wall = time.perf_counter() - t0
    ms, by = eng.last_set_tensor()
    out["set_tensor"] = {"sample_slabs": q_up, "host_bytes_full_square": 8.0 * n * n * q_up, "pcie_bytes": by,
                         "ms": ms, "wall_ms": 1e3 * wall, "host_gbs_full_square": 8.0 * n * n * q_up / wall * 1e-9,
                         "full_tensor_s_at_this_rate": wall * naux / q_up,
                         "what": "mqcb200_set_tensor_shard from a pageable host array: lower triangles gathered into two "
                                 "pinned buffers, DMA + packing kernel of chunk k overlap the gather of chunk k+1"}
    eng.clear_tensor(1)
    del b_host
    if 8.0 * n * n * naux > 16e9:
        out["whiten"] = {"skipped": "the host-side stand-in for (mu nu|P) is sized for c2; run --workload c2 for this block"}
        return out
    # build_df_tensor's last two stages on the device: metric^(-1/2) (one-sided Jacobi + GEMM), then
    # b = three . half streamed nu-slab by nu-slab (64 orbital columns per push, every push the same
    # host block: the arithmetic does not care, the PCIe traffic is the real one)
    u, _ = np.linalg.qr(rng.standard_normal((naux, naux)))
    lam = np.exp(rng.uniform(-2.0, 2.0, size=naux))
    metric = np.asfortranarray((u * lam[None, :]) @ u.T)
    metric = np.asfortranarray(0.5 * (metric + metric.T))
    t0 = time.perf_counter()
    half = eng.metric_inverse_sqrt(metric)
    wall_metric = time.perf_counter() - t0
    m_ms, m_sweeps = eng.last_metric()
    out["metric_inverse_sqrt"] = {"naux": naux, "device_ms": m_ms, "wall_ms": 1e3 * wall_metric, "jacobi_sweeps": m_sweeps,
                                  "what": "mqcb200_metric_inverse_sqrt: one-sided Jacobi on the GPU + (U s^-1/2) U^T"}
    slab_nu = 64
    block = np.asfortranarray(rng.standard_normal((n * slab_nu, naux)))
    eng.whiten_begin(n, naux, half, slot=1)
    t0 = time.perf_counter()
    nu, pcie = 0, 0.0
    while nu < n:
        cnt = min(slab_nu, n - nu)
        eng.whiten_push(nu, block[: n * cnt, :] if cnt < slab_nu else block, slot=1)
        pcie += 8.0 * (n - nu) * cnt * naux              # only the rows mu >= nu of a slab cross the bus
        nu += cnt
    eng.whiten_end(slot=1)
    wall = time.perf_counter() - t0
    w_ms, w_flops = eng.last_whiten()
    out["whiten"] = {"gemm_ms": w_ms, "gemm_tflops": w_flops / w_ms * 1e-9 if w_ms > 0 else None,
                     "flops": w_flops, "wall_ms_with_pcie": 1e3 * wall, "pcie_bytes": pcie, "host_bytes_full_square": 8.0 * n * n * naux,
                     "reference_flops_2_nao2_naux2": 2.0 * n * n * float(naux) ** 2,
                     "what": "mqcb200_whiten_begin/push/end: b = three . half slab by slab into the packed layout "
                             "(only the packed lower-triangular positions are formed: about half the reference's flops)"}
    eng.clear_tensor(1)
    return out


def run_scf_loop(ctx, name):
    """A whole closed-shell SCF at the workload's shape on the (possibly sharded) resident tensor, every matrix
    staying on the GPU (mqcb200_scf: Fock build + commutator + DIIS + one-sided Jacobi + density per iteration)
    -- and, on one GPU, the same loop driven from the host (C-ABI Fock build + NumPy/LAPACK step) beside it."""
    eng, world, rank = ctx.eng, ctx.world, ctx.rank
    from metalquicha_b200 import synth
    cfg = _workload_cfg(name)
    n, n_occ, naux = cfg["n"], cfg["n_occ"], cfg["naux"]
    rng = np.random.default_rng(n)
    a = rng.standard_normal((n, n)) * (0.15 / np.sqrt(n))
    s = np.asfortranarray(np.eye(n) + a + a.T)
    h = np.asfortranarray(synth.synth_core_hamiltonian(n, n) - 2.0 * np.diag(np.linspace(1.0, 0.0, n)))
    q_begin, q_count = synth.shard_range(naux, world, rank)
    eng.synth_tensor(n, naux, 5, 0.3 * synth.default_scale(n, naux), q_begin=q_begin, q_count=q_count)
    eng.run_scf(h, s, 2 * n_occ, max_iter=3)                      # warm (allocations)
    t0 = time.perf_counter()
    r = eng.run_scf(h, s, 2 * n_occ, max_iter=40)
    t_dev = time.perf_counter() - t0
    out = {"iterations": int(r["iterations"]), "converged": bool(r["converged"]), "electronic": float(r["electronic"]),
           "device_resident_ms_per_iteration": 1e3 * t_dev / (r["iterations"] + 1),
           "what": "mqcb200_scf on synthetic (H, S, B) of this shape: per iteration one Fock build + the SCF step, "
                   "every matrix on the GPU; the +1 is the final rebuild"}
    if world == 1 and rank == 0 and not ctx.args.no_cpu_baseline:
        from oracle import scf_oracle as host_scf                # host-driven baseline leg

        def fock_builder(hh, dd, cc, no):
            f = eng.build_fock_df(np.asfortranarray(hh), np.asfortranarray(dd), np.asfortranarray(cc), no)
            return f, eng.last_energy()
        t0 = time.perf_counter()
        ref = host_scf.run_rhf(h, s, 2 * n_occ, fock_builder, max_iter=40)
        t_host = time.perf_counter() - t0
        out["host_driven_ms_per_iteration"] = 1e3 * t_host / (ref["iterations"] + 1)
        out["host_driven_iterations"] = int(ref["iterations"])
        out["energy_difference"] = abs(float(ref["electronic"]) - float(r["electronic"]))
    return out


# ------------------------------------------------------------------------------------------
# fragment farm (BASELINE configs[2]): independent fragments, one per GPU worker at a time
# ------------------------------------------------------------------------------------------
FARM_BUILDS_PER_FRAGMENT = 12          # guess + ~10 SCF iterations + final rebuild (rhf.f90:546,566,646)
FARM_BATCH = {"c3_monomer": 1, "c3_dimer": 12, "c3": 243}   # 64 : 2016 : 41664 scaled to 256 fragments


def run_farm(ctx, primary):
    """One step = one batch of 256 MBE-3 fragments (monomers/dimers/trimers in the run's
    proportions), largest first, pulled by every GPU worker from a shared FIFO (queue_t
    semantics); each fragment = set its tensor FROM A HOST bmat (n*n, naux) + FARM_BUILDS_PER_FRAGMENT
    Fock builds through the host-buffer C-ABI call (that is how a fragment's SCF drives the
    engine, so `value` and `e2e` coincide for this workload)."""
    torch, dist, args, world, rank, local_rank = ctx.torch, ctx.dist, ctx.args, ctx.world, ctx.rank, ctx.local_rank
    from metalquicha_b200 import B200FockEngine, WorkQueue, farm, synth

    # one batch per step and per GPU: a real MBE-3 run of (H2O)64 has 43 744 fragments, so the
    # sample grows with the machine instead of starving 8 x 4 workers on 256 fragments
    frags = []
    for nm, count in FARM_BATCH.items():
        frags += [nm] * (count * world)
    sizes = [synth.CONFIGS[f]["n"] for f in frags]
    order = farm.sort_fragments_largest_first(sizes)
    problems = {}
    for nm in FARM_BATCH:
        c = synth.CONFIGS[nm]
        b, h, d, co = synth.synth_problem(SEED, c["n"], c["n_occ"], c["naux"], with_tensor=True)
        problems[nm] = (c, h, d, co, b)

    n_workers = max(1, args.workers_per_gpu)
    engines = [B200FockEngine(local_rank) for _ in range(n_workers)]
    store = dist.distributed_c10d._get_default_store() if world > 1 else None
    builds_done = [0] * n_workers
    launches = [0] * n_workers
    mode = {"host_tensor": True}

    def do_fragment(eng, widx, idx):
        c, h, d, co, b = problems[frags[idx]]
        if mode["host_tensor"]:
            eng.set_tensor(b, n=c["n"])                # per-fragment bmat: host -> packed, inside the timed region
        else:
            eng.synth_tensor(c["n"], c["naux"], SEED + idx, synth.default_scale(c["n"], c["naux"]))
        e = 0.0
        for _ in range(FARM_BUILDS_PER_FRAGMENT):
            eng.build_fock_df(h, d, co, c["n_occ"])
            e = eng.last_energy()
            launches[widx] += eng.last_launches() + 1
        builds_done[widx] += FARM_BUILDS_PER_FRAGMENT
        return e

    def run_step(step_id):
        if world > 1:
            q = farm.DistributedWorkQueue(order, store, name=f"farm{step_id}", batch=args.farm_pop)
        else:
            q = WorkQueue(order)
        threads = [threading.Thread(target=farm.worker_loop, args=(q, lambda i, w=w: do_fragment(engines[w], w, i)))
                   for w in range(n_workers)]
        [t.start() for t in threads]
        [t.join() for t in threads]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(tag, k):
        for w in range(max(1, min(args.warmup, 2))):
            run_step(f"{tag}w{w}")
        builds_done[:] = [0] * n_workers
        launches[:] = [0] * n_workers
        barrier()
        t0 = time.perf_counter()
        for s_ in range(k):
            run_step(f"{tag}s{s_}")
        torch.cuda.synchronize()
        t_local = time.perf_counter() - t0
        barrier()
        tt = torch.tensor([t_local, float(sum(builds_done)), float(sum(launches))], dtype=torch.float64, device="cuda")
        if world > 1:
            tmax = tt.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
            tsum = tt.clone(); dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
            return float(tmax[0]), float(tsum[1]), float(tsum[2])
        return float(tt[0]), float(tt[1]), float(tt[2])

    k = args.steps if primary else max(2, min(args.steps, 5))
    sampler = ClockSampler(local_rank)
    if rank == 0 and primary:
        sampler.start()
    t_all, builds_all, launches_all = timed("h", k)
    clocks = sampler.stop() if (rank == 0 and primary) else None
    mode["host_tensor"] = False
    t_dev, builds_dev, _ = timed("d", k)
    for e in engines:
        e.close()
    scf_block = run_farm_scf(ctx, frags, order, k)
    trimer = synth.CONFIGS["c3"]
    cpu_baseline = None
    if rank == 0 and world == 1 and primary and not args.no_cpu_baseline:
        from oracle import df_fock_blas as blas_port
        c, h, d, co, b = problems["c3"]
        ts = []
        with blas_port.blas_threads(blas_port.host_threads()):
            for _ in range(3):
                t1 = time.perf_counter(); blas_port.build_fock_df(h, b, d, co, c["n_occ"]); ts.append(time.perf_counter() - t1)
        cpu_baseline = {"value": 1.0 / min(ts), "unit": "builds/s", "cores": blas_port.host_threads(), "kind": "port",
                        "sample": "one (H2O)3 trimer build (95 % of the batch is trimers), C/OpenBLAS restatement, best of 3"}
    value = builds_all / t_all
    h2d_build = 8 * (2 * trimer["n"] ** 2 + trimer["n"] * trimer["n_occ"])
    tri = trimer["n"] * (trimer["n"] + 1) // 2
    return {
        "value": value, "unit": "builds/s", "ms_per_step": 1e3 * t_all / k, "steps": k,
        "config": _config("c3", trimer, world, n_workers),
        "value_tensor_synthesised_on_device": builds_dev / t_dev,
        "note": "value INCLUDES one mqcb200_set_tensor from a host bmat(n*n, naux) per fragment (lower triangles "
                "over PCIe + packing); value_tensor_synthesised_on_device is round 1's figure (tensor generated on the GPU)",
        "e2e": {"value": value, "unit": "builds/s", "ms_per_step": 1e3 * t_all / k,
                "h2d_bytes_per_step": (h2d_build * FARM_BUILDS_PER_FRAGMENT + 8 * tri * trimer["naux"]) * 256 * world,
                "d2h_bytes_per_step": 8 * trimer["n"] ** 2 * 256 * FARM_BUILDS_PER_FRAGMENT * world,
                "note": "every build goes through mqcb200_build_fock with host buffers"},
        "gpu_launches": int(launches_all),
        "roofline": _farm_roofline(trimer, builds_dev / t_dev / world),
        "cpu_baseline": cpu_baseline, "clocks": clocks, "device_resident_scf": scf_block,
    }


def _farm_roofline(trimer, builds_per_s_per_gpu):
    """One fragment build streams its packed tensor once (fragment_jk_kernel: gamma, J, X, K per slab in one
    pass): 8 * npair * naux algorithmic bytes.  Against the HBM peak that is a few per cent -- the farm is
    bound by launch latency and the host, not by a device roofline; the figure says how far."""
    hbm_peak, hbm_src = _peaks()
    by = 8.0 * (trimer["n"] * (trimer["n"] + 1) // 2) * trimer["naux"]
    ach = by * builds_per_s_per_gpu * 1e-9
    return {"kernel": "fragment_jk_kernel", "bound": "hbm", "achieved": ach, "peak": hbm_peak, "unit": "GB/s",
            "frac": ach / hbm_peak if hbm_peak else None, "traffic": None, "peak_source": hbm_src,
            "note": "per GPU, tensors resident (value_tensor_synthesised_on_device): 79 MFLOP and 7 MB per trimer build "
                    "are launch/latency-bound, see DESIGN.md 4.3"}


def _synthetic_fragment(seed, n, naux, coupling=0.35):
    """(S, H, b) of a fragment-sized synthetic SCF problem that converges like a closed-shell
    molecule does: S = 1 + small symmetric, H with a ladder on the diagonal, b = coupling * synth."""
    from metalquicha_b200 import synth
    rng = np.random.default_rng(seed)
    a = rng.standard_normal((n, n)) * (0.15 / np.sqrt(n))
    s = np.eye(n) + a + a.T
    h = synth.synth_core_hamiltonian(seed, n) - 2.0 * np.diag(np.linspace(1.0, 0.0, n))
    b = coupling * synth.synth_tensor(seed, n, naux)
    return np.asfortranarray(s), np.asfortranarray(h), np.asfortranarray(b)


def run_farm_scf(ctx, frags, order, k):
    """The same 256-fragment batch with each fragment's WHOLE SCF on the GPU (mqcb200_scf_fragment:
    tensor from a host bmat, then FARM_BUILDS_PER_FRAGMENT - 1 iterations of build + commutator + DIIS +
    eigensolver + density on the device, and the final rebuild) -- against the same loop driven from the
    host (engine Fock build through the C ABI + NumPy/LAPACK SCF step), one worker."""
    torch, dist, args, world, rank, local_rank = ctx.torch, ctx.dist, ctx.args, ctx.world, ctx.rank, ctx.local_rank
    from metalquicha_b200 import B200FockEngine, WorkQueue, farm, synth
    n_workers = max(1, args.scf_workers_per_gpu)
    iters = FARM_BUILDS_PER_FRAGMENT - 1
    problems = {}
    for nm in FARM_BATCH:
        c = synth.CONFIGS[nm]
        problems[nm] = (c,) + _synthetic_fragment(900 + c["n"], c["n"], c["naux"])
    engines = [B200FockEngine(local_rank) for _ in range(n_workers)]
    store = dist.distributed_c10d._get_default_store() if world > 1 else None
    done = [0] * n_workers

    def do_fragment(eng, widx, idx):
        c, s_, h_, b_ = problems[frags[idx]]
        eng.set_tensor(b_, n=c["n"])
        r = eng.run_scf_fragment(h_, s_, 2 * c["n_occ"], max_iter=iters, energy_tol=0.0, density_tol=0.0)   # never "converges": fixed work
        done[widx] += r["iterations"] + 1
        return r["electronic"]

    def run_step(step_id):
        q = farm.DistributedWorkQueue(order, store, name=f"scf{step_id}", batch=args.farm_pop) if world > 1 else WorkQueue(order)
        threads = [threading.Thread(target=farm.worker_loop, args=(q, lambda i, w=w: do_fragment(engines[w], w, i)))
                   for w in range(n_workers)]
        [t.start() for t in threads]
        [t.join() for t in threads]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
    run_step("w")
    done[:] = [0] * n_workers
    barrier()
    t0 = time.perf_counter()
    for s_i in range(k):
        run_step(f"s{s_i}")
    torch.cuda.synchronize()
    t_local = time.perf_counter() - t0
    barrier()
    tt = torch.tensor([t_local, float(sum(done))], dtype=torch.float64, device="cuda")
    if world > 1:
        tmax = tt.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = tt.clone(); dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        t_all, builds = float(tmax[0]), float(tsum[1])
    else:
        t_all, builds = float(tt[0]), float(tt[1])
    out = {"builds_per_s": builds / t_all, "fragment_scfs_per_s": builds / FARM_BUILDS_PER_FRAGMENT / t_all,
           "workers_per_gpu": n_workers, "iterations_per_fragment": iters,
           "what": "every build here is followed ON THE GPU by the SCF step the reference does on the host "
                   "(commutator, DIIS, eigensolver, density, convergence test); per iteration a few scalars cross PCIe"}
    # host-driven baseline on rank 0: same fragments, same iteration count, one worker
    if rank == 0 and not args.no_cpu_baseline:
        from oracle import scf_oracle as host_scf          # CPU-side baseline leg (NumPy/LAPACK SCF step)
        eng = engines[0]
        c, s_, h_, b_ = problems["c3"]

        def fock_builder(hh, dd, cc, n_occ):
            f = eng.build_fock_df(np.asfortranarray(hh), np.asfortranarray(dd), np.asfortranarray(cc), n_occ)
            return f, eng.last_energy()
        ts = []
        for _ in range(4):
            t1 = time.perf_counter()
            eng.set_tensor(b_, n=c["n"])
            host_scf.run_rhf(h_, s_, 2 * c["n_occ"], fock_builder, max_iter=iters, energy_tol=0.0, density_tol=0.0)
            ts.append(time.perf_counter() - t1)
        t1 = time.perf_counter()
        eng.set_tensor(b_, n=c["n"])
        eng.run_scf_fragment(h_, s_, 2 * c["n_occ"], max_iter=iters, energy_tol=0.0, density_tol=0.0)
        t_dev_one = time.perf_counter() - t1
        out["one_trimer_scf_ms"] = {"device_resident": 1e3 * t_dev_one, "host_driven": 1e3 * min(ts),
                                    "note": "host-driven = mqcb200_build_fock per iteration + the NumPy/LAPACK SCF step "
                                            "(oracle/scf_oracle.py, the restatement of run_libcint_rhf's loop), one thread"}
    for e in engines:
        e.close()
    out["batched"] = run_farm_scf_batched(ctx, problems, k)
    return out


def run_farm_scf_batched(ctx, problems, k):
    """The same batch once more with the fragments of one kind driven IN LOCK-STEP
    (mqcb200_scf_fragment_batch): up to 32 fragments per call, their tensors back to back on the slot,
    every kernel of an iteration launched once with the fragment index on a grid axis.  Two figures:
    with every fragment's bmat coming from the host inside the timed region, and with the tensors
    already resident (synthesised on the device) -- the SCF machinery alone."""
    torch, dist, args, world, rank, local_rank = ctx.torch, ctx.dist, ctx.args, ctx.world, ctx.rank, ctx.local_rank
    from metalquicha_b200 import B200FockEngine, WorkQueue, farm, synth
    iters = FARM_BUILDS_PER_FRAGMENT - 1
    bsz = max(1, args.scf_batch)
    n_workers = max(1, args.scf_batch_workers_per_gpu)
    batches = []                                            # (kind, fragments in the batch), largest kind first
    reps = 4                                                # four 256-fragment samples per GPU and step: a few calls per worker
    for nm in sorted(FARM_BATCH, key=lambda x: -synth.CONFIGS[x]["n"]):
        left = FARM_BATCH[nm] * world * reps
        while left > 0:
            batches.append((nm, min(bsz, left)))
            left -= min(bsz, left)
    host_b = {nm: np.asfortranarray(np.hstack([problems[nm][3]] * min(bsz, FARM_BATCH[nm] * world * reps))) for nm in FARM_BATCH}
    stacks = {nm: (np.stack([problems[nm][2]] * bsz), np.stack([problems[nm][1]] * bsz)) for nm in FARM_BATCH}
    engines = [B200FockEngine(local_rank) for _ in range(n_workers)]
    store = dist.distributed_c10d._get_default_store() if world > 1 else None
    done = [0] * n_workers
    mode = {"host_tensor": True}

    def do_batch(eng, widx, idx):
        nm, cnt = batches[idx]
        c = problems[nm][0]
        if mode["host_tensor"]:
            eng.set_tensor(host_b[nm][:, :cnt * c["naux"]], n=c["n"])
        else:
            eng.synth_tensor(c["n"], cnt * c["naux"], SEED + idx, 0.35 * synth.default_scale(c["n"], c["naux"]))
        h_all, s_all = stacks[nm]
        r = eng.run_scf_fragment_batch(h_all[:cnt], s_all[:cnt], 2 * c["n_occ"], max_iter=iters, energy_tol=0.0,
                                       density_tol=0.0, want_matrices=False, check_every=iters)
        done[widx] += int(np.sum(r["iterations"] + 1))
        return float(r["electronic"][0])

    order = list(range(len(batches)))

    def run_step(step_id):
        q = farm.DistributedWorkQueue(order, store, name=f"scfb{step_id}", batch=1) if world > 1 else WorkQueue(order)
        threads = [threading.Thread(target=farm.worker_loop, args=(q, lambda i, w=w: do_batch(engines[w], w, i)))
                   for w in range(n_workers)]
        [t.start() for t in threads]
        [t.join() for t in threads]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(tag):
        run_step(tag + "w")
        done[:] = [0] * n_workers
        barrier()
        t0 = time.perf_counter()
        for s_i in range(k):
            run_step(f"{tag}{s_i}")
        torch.cuda.synchronize()
        t_local = time.perf_counter() - t0
        barrier()
        tt = torch.tensor([t_local, float(sum(done))], dtype=torch.float64, device="cuda")
        if world > 1:
            tmax = tt.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
            tsum = tt.clone(); dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
            return float(tsum[1]) / float(tmax[0])
        return float(tt[1]) / float(tt[0])
    rate_host = timed("h")
    mode["host_tensor"] = False
    rate_resident = timed("r")
    for e in engines:
        e.close()
    return {"builds_per_s": rate_host, "builds_per_s_tensors_resident": rate_resident,
            "fragment_scfs_per_s": rate_host / FARM_BUILDS_PER_FRAGMENT, "fragments_per_call": bsz,
            "workers_per_gpu": n_workers, "batches_per_step": len(batches),
            "what": "mqcb200_scf_fragment_batch: fragments of one kind in lock-step, six launches per iteration whatever the "
                    "batch size, one SCF-step CTA per fragment; builds_per_s includes each fragment's bmat coming from the host "
                    "(7 MB of lower triangles per trimer over PCIe), builds_per_s_tensors_resident does not"}


# ------------------------------------------------------------------------------------------
def run_reference_pin(make_plain_engine):
    """Parity against a number the REFERENCE holds, inside the bench line: the reference's own density-fitted
    validation case -- water / 6-31G* fitted with 6-31G*, validation/validation_tests_cpu.json:899-904,
    E = -76.188111755038, manifest tolerance 1e-9 Eh -- with every Fock build of the SCF loop done by the engine
    (bmat from the restated integrals, oracle/gto_integrals.py; the loop is the oracle's restatement of
    run_libcint_rhf: checker code, as in the parity block), then the whole loop device-resident on the same tensor.
    Not timed, rank 0 only, on its own engine handle (no communicator)."""
    import numpy as np
    from oracle import df_fock_oracle as oracle
    from oracle import gto_integrals as gto
    from oracle import scf_oracle as scf
    t0 = time.perf_counter()
    s_mat, h_mat, three, metric, e_nuc, n_electrons, e_ref = gto.df_case_integrals("h2o_631gs")
    eng = make_plain_engine()
    try:
        eng.set_tensor(np.asfortranarray(oracle.whiten(three, metric)))

        def fock_builder(h_, density, coeff, n_occ):
            f = eng.build_fock_df(np.asfortranarray(h_), np.asfortranarray(density), np.asfortranarray(coeff), n_occ)
            return f, eng.last_energy()
        res = scf.run_rhf(h_mat, s_mat, n_electrons, fock_builder, e_nuc=e_nuc)
        dev = eng.run_scf(np.asfortranarray(h_mat), np.asfortranarray(s_mat), n_electrons, e_nuc=e_nuc)
        # the two-spin build against the reference-held unrestricted energy (UHF OH/STO-3G doublet,
        # validation_tests_cpu.json:1844-1849) on a tensor that fits the four-index integrals exactly
        try:
            symbols, coords, n_el_u, mult, e_ref_u = gto.OH_STO3G_UHF
            s_u, h_u, eri_u, e_nuc_u = gto.molecule_integrals(symbols, coords, gto.STO3G_BSE)
            eng.set_tensor(np.asfortranarray(scf.exact_fit_tensor(eri_u)))

            def uhf_builder(h_, d_a, d_b, c_a, n_a, c_b, n_b):
                f_a, f_b = eng.build_fock_df_uhf(np.asfortranarray(h_), np.asfortranarray(d_a), np.asfortranarray(d_b),
                                                 np.asfortranarray(c_a), n_a, np.asfortranarray(c_b), n_b)
                return f_a, f_b, oracle.uhf_electronic_energy(h_, f_a, f_b, d_a, d_b)
            uhf = scf.run_uhf(h_u, s_u, n_el_u, mult, uhf_builder, e_nuc=e_nuc_u)
            err_u = float(abs(uhf["energy"] - e_ref_u))
            two_spin = {"case": "UHF OH/STO-3G doublet (validation/validation_tests_cpu.json:1844-1849), two-spin Fock builds on the GPU",
                        "reference_held_energy": float(e_ref_u), "energy": float(uhf["energy"]), "abs_err": err_u,
                        "spin_squared": float(uhf["spin_squared"]), "ok": bool(uhf["converged"] and err_u <= 1e-9)}
        except Exception as ex:
            two_spin = {"error": str(ex)[:300]}
    finally:
        eng.close()
    err, err_dev = float(abs(res["energy"] - e_ref)), float(abs(dev["energy"] - e_ref))
    return {"case": "H2O RHF/6-31G* density-fitted with 6-31G* (validation/validation_tests_cpu.json:899-904 of the reference)",
            "reference_held_energy": float(e_ref), "tolerance": 1e-9,
            "energy_fock_builds_on_gpu": float(res["energy"]), "abs_err": err, "iterations": int(res["iterations"]),
            "energy_whole_loop_on_gpu": float(dev["energy"]), "abs_err_whole_loop_on_gpu": err_dev,
            "ok": bool(res["converged"] and err <= 1e-9 and dev["converged"] and err_dev <= 1e-7),
            "two_spin": two_spin,
            "wall_s": time.perf_counter() - t0}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="c4")
    ap.add_argument("--secondary", default=None, help="comma list of further workloads for the `workloads` block; 'none' to skip")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--cpu-sample", type=int, default=None, help="auxiliary functions in the CPU-port sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--no-setup-timings", action="store_true")
    ap.add_argument("--no-reference-pin", action="store_true")
    ap.add_argument("--k-scale", type=float, default=None)
    ap.add_argument("--workers-per-gpu", type=int, default=4, help="fragment farm: host workers (engine handles) per GPU")
    ap.add_argument("--farm-pop", type=int, default=4, help="fragment farm: ids taken per round trip to the shared queue")
    ap.add_argument("--scf-workers-per-gpu", type=int, default=8, help="fragment farm, device-resident SCF mode: workers per GPU")
    ap.add_argument("--scf-batch", type=int, default=128, help="fragment farm, batched device SCF: fragments per call")
    ap.add_argument("--scf-batch-workers-per-gpu", type=int, default=2, help="fragment farm, batched device SCF: workers per GPU")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)

    cfg = _workload_cfg(args.workload)
    world = _env_int("WORLD_SIZE", 1)
    if args.steps is None:
        # sized so that the timed region is about a second
        per_rank = 3.0 * cfg["n"] ** 2 * (cfg["n_occ"] + cfg["n_beta"]) * cfg["naux"] / max(1, world)
        args.steps = 5 if args.workload == "c3" else int(min(200, max(5, 1.0 / (per_rank / 28e12 + 2e-4))))
    if args.impl == "reference":
        run_reference(args, cfg)
        return
    if args.secondary is None:
        args.secondary = ",".join(w for w in ("c2", "c5", "c3") if w != args.workload) if args.workload == "c4" else "none"
    secondary = [w for w in args.secondary.split(",") if w and w != "none"]

    import torch
    import torch.distributed as dist
    from metalquicha_b200 import B200FockEngine

    ctx = Ctx()
    ctx.torch, ctx.dist, ctx.args = torch, dist, args
    ctx.world, ctx.rank, ctx.local_rank = world, _env_int("RANK", 0), _env_int("LOCAL_RANK", 0)
    if world == 1 and args.gpus > 1:
        raise SystemExit("--gpus N>1 must be launched with torch.distributed.run (one rank per GPU)")
    torch.cuda.set_device(ctx.local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", ctx.local_rank))
    ctx.fp64_peak = _fp64_peak(torch)

    def make_engine():
        eng = B200FockEngine(ctx.local_rank)
        if world > 1:
            uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
            if ctx.rank == 0:
                uid.copy_(torch.frombuffer(bytearray(B200FockEngine.comm_unique_id()), dtype=torch.uint8))
            dist.broadcast(uid, 0)
            eng.comm_init(world, ctx.rank, bytes(uid.cpu().numpy().tobytes()))
        return eng

    t_start = time.perf_counter()
    results = {}
    for i, name in enumerate([args.workload] + secondary):
        primary = i == 0
        if name == "c3":
            results[name] = run_farm(ctx, primary)
        else:
            ctx.eng = make_engine()
            results[name] = run_builds(ctx, name, primary)
            if ctx.rank == 0 and world == 1 and not args.no_setup_timings and name in ("c2", args.workload):
                try:
                    results[name]["setup_timings"] = run_setup_timings(ctx, name)
                except Exception as ex:          # separately timed extras never sink the headline
                    results[name]["setup_timings"] = {"error": str(ex)[:200]}
            if name == "c2" and not args.no_setup_timings:
                try:
                    results[name]["scf_loop"] = run_scf_loop(ctx, name)
                except Exception as ex:
                    results[name]["scf_loop"] = {"error": str(ex)[:200]}
            ctx.eng.close()
            ctx.eng = None
            torch.cuda.empty_cache()
        log(f"{name} done at {time.perf_counter() - t_start:.1f} s")

    reference_pin = None
    if ctx.rank == 0 and not args.no_reference_pin and not args.no_parity:
        try:                                     # before the barrier: the other ranks wait there meanwhile
            reference_pin = run_reference_pin(lambda: B200FockEngine(ctx.local_rank))
            log(f"reference-held DF energy: |dE| = {reference_pin['abs_err']:.2e} (Fock builds on the GPU), "
                f"{reference_pin['abs_err_whole_loop_on_gpu']:.2e} (whole loop on the GPU)")
        except Exception as ex:                  # a check beside the headline never sinks it
            reference_pin = {"error": str(ex)[:300]}
    if world > 1:
        dist.barrier()
    if ctx.rank == 0:
        head = results[args.workload]
        line = {
            "metric": METRIC, "value": head["value"], "unit": "builds/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": head["ms_per_step"],
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
        }
        for key, val in head.items():
            if key not in ("value", "unit", "ms_per_step", "steps"):
                line[key] = val
        line["workloads"] = {k: v for k, v in results.items() if k != args.workload}
        if reference_pin is not None:
            line["reference_pin"] = reference_pin
        line["bench_wall_s"] = time.perf_counter() - t_start
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
