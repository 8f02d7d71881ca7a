#!/usr/bin/env python
"""Benchmark of the DF-J/K Fock build (BASELINE.json's metric) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c2|c4|c5|c3] [--impl reference]

One "step" is one complete Fock build F = H + J - K/2 on a resident fitted tensor
(mqcb200_build_fock == the reference's build_fock_df).  At N=1 the default workload
is BASELINE.json configs[1]: (H2O)16 RHF/def2-TZVP, (n, n_occ, naux) = (688, 80, 1800).
For N>1 (torchrun, one rank per GPU) the same build is sharded over the auxiliary
index with one NCCL all-reduce of [J;K] per build -- total work fixed, "strong".

value  : builds/s with H, D, C already in HBM (mqcb200_build_fock_device), CUDA-event
         timed on the engine's stream, max over ranks.
e2e    : builds/s through mqcb200_build_fock with pinned HOST buffers: H2D of H, D, C and
         D2H of F inside the timed region.
roofline / kernels : per-kernel device time from CUDA events recorded by the engine on its
         own stream inside the timed region; algorithmic bytes/flops per DESIGN.md.
cpu_baseline : the NumPy/OpenBLAS restatement of the reference's CPU path (oracle/, "port":
         the Fortran reference cannot be built in this image) on a bounded auxiliary
         sub-sample, extrapolated linearly in naux.
--impl reference : the same CPU port as the reference arm.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SEED = 20261018


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


# ------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference's build_fock_df on a bounded sub-sample
# ------------------------------------------------------------------------------------------
def cpu_port_time(n, n_occ, naux, q_sample, reps, two_spin_nb=None):
    """Seconds per FULL build, extrapolated from `q_sample` auxiliary functions."""
    from metalquicha_b200 import synth
    from oracle import df_fock_oracle as oracle
    qs = min(q_sample, naux)
    scale = synth.default_scale(n, naux)
    b = synth.synth_tensor(SEED, n, naux, scale, q_begin=0, q_count=qs)
    _, h, density, coeff = synth.synth_problem(SEED, n, n_occ, naux, with_tensor=False)
    times = []
    fock = None
    for _ in range(reps):
        t0 = time.perf_counter()
        fock = oracle.build_fock_df(h, b, density, coeff, n_occ)
        times.append(time.perf_counter() - t0)
    t = min(times)
    return t * (naux / qs), t, qs, (b, h, density, coeff, fock)


def run_reference(args, cfg):
    rank = _env_int("RANK", 0)
    if rank != 0:
        return
    n, n_occ, naux = cfg["n"], cfg["n_occ"], cfg["naux"]
    cores = os.cpu_count() or 1
    # warm-up + timed "steps", each a bounded sample of the workload
    q_sample = args.cpu_sample
    total_reps = args.warmup + args.steps
    from metalquicha_b200 import synth
    from oracle import df_fock_oracle as oracle
    qs = min(q_sample, naux)
    b = synth.synth_tensor(SEED, n, naux, synth.default_scale(n, naux), 0, qs)
    _, h, density, coeff = synth.synth_problem(SEED, n, n_occ, naux, with_tensor=False)
    times = []
    for r in range(total_reps):
        t0 = time.perf_counter()
        oracle.build_fock_df(h, b, density, coeff, n_occ)
        if r >= args.warmup:
            times.append(time.perf_counter() - t0)
    t_full = (sum(times) / len(times)) * (naux / qs)
    value = 1.0 / t_full
    sample = (f"{qs} of {naux} auxiliary functions per step (J and K are linear in naux), "
              f"time scaled by {naux}/{qs}; NumPy loop-for-loop port of build_fock_df, OpenBLAS threads={cores}")
    line = {
        "impl": "reference", "metric": "DF-J/K Fock builds/sec", "value": value, "unit": "builds/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_full,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": _config(args, cfg),
        "cpu_baseline": {"value": value, "unit": "builds/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "builds/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def _config(args, cfg):
    return {"workload": f"{args.workload}: {cfg['what']}", "n_ao": cfg["n"], "n_occ": cfg.get("n_occ"),
            "naux": cfg["naux"], "parallelism": f"aux-sharded x{args.gpus}" if args.gpus > 1 else "single GPU",
            "l2": "packed tensor (>= 3.4 GB at c2) is larger than the 126 MB L2; no flush needed",
            "timing": "CUDA events on the engine stream, max over ranks"}


# ------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="c2")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--cpu-sample", type=int, default=200, help="auxiliary functions in the CPU-port sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--k-scale", type=float, default=None)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)

    from metalquicha_b200 import synth
    if args.workload not in synth.CONFIGS:
        raise SystemExit(f"unknown workload {args.workload}; choose from {sorted(synth.CONFIGS)}")
    cfg = dict(synth.CONFIGS[args.workload])
    if "n_occ" not in cfg:
        cfg["n_occ"] = cfg["n_alpha"]
    if args.impl == "reference":
        run_reference(args, cfg)
        return

    import torch
    import torch.distributed as dist
    from metalquicha_b200 import B200FockEngine

    world = _env_int("WORLD_SIZE", 1)
    rank = _env_int("RANK", 0)
    local_rank = _env_int("LOCAL_RANK", 0)
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("--gpus N>1 must be launched with torch.distributed.run (one rank per GPU)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    n, n_occ, naux = cfg["n"], cfg["n_occ"], cfg["naux"]
    k_scale = args.k_scale if args.k_scale is not None else (0.2 if args.workload == "c4" else 1.0)
    scale = synth.default_scale(n, naux)
    _, h, density, coeff = synth.synth_problem(SEED, n, n_occ, naux, with_tensor=False)

    eng = B200FockEngine(local_rank)
    q_begin, q_count = synth.shard_range(naux, world, rank)
    if world > 1:
        uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            uid.copy_(torch.frombuffer(bytearray(B200FockEngine.comm_unique_id()), dtype=torch.uint8))
        dist.broadcast(uid, 0)
        eng.comm_init(world, rank, bytes(uid.cpu().numpy().tobytes()))
    t_synth0 = time.perf_counter()
    eng.synth_tensor(n, naux, SEED, scale, q_begin=q_begin, q_count=q_count)
    t_synth = time.perf_counter() - t_synth0

    # device-resident operands (column-major == the transpose of a row-major torch tensor;
    # H and D are symmetric, C is passed as its (n_occ, n) row-major image)
    d_h = torch.from_numpy(np.ascontiguousarray(h.T)).cuda()
    d_d = torch.from_numpy(np.ascontiguousarray(density.T)).cuda()
    d_c = torch.from_numpy(np.ascontiguousarray(coeff.T)).cuda()
    d_f = torch.empty_like(d_h)
    # pinned host operands for the end-to-end arm
    p_h = torch.from_numpy(np.asfortranarray(h).T.copy()).pin_memory()       # memory == column-major h
    p_d = torch.from_numpy(np.asfortranarray(density).T.copy()).pin_memory()
    p_c = torch.from_numpy(np.asfortranarray(coeff).T.copy()).pin_memory()
    p_f = torch.empty(n, n, dtype=torch.float64).pin_memory()
    h_np, d_np, f_np = p_h.numpy().T, p_d.numpy().T, p_f.numpy().T           # column-major views
    c_np = p_c.numpy().T

    stream = torch.cuda.ExternalStream(eng.stream())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        return ms

    # ---- device-resident arm -------------------------------------------------------------
    for _ in range(args.warmup):
        eng.build_fock_device(d_h, d_d, d_c, n_occ, d_f, k_scale=k_scale, sync=True)
    launches_per_build = eng.last_launches()
    eng.set_profiling(True)
    eng.last_timings()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    for _ in range(args.steps):
        eng.build_fock_device(d_h, d_d, d_c, n_occ, d_f, k_scale=k_scale, sync=False)
    e1.record(stream)
    e1.synchronize()
    barrier()
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    clocks = sampler.stop() if rank == 0 else None
    phase = eng.last_timings()                       # sums over the timed steps, this rank
    eng.set_profiling(False)
    ms_per_step = ms_total / args.steps
    value = 1e3 / ms_per_step
    fock_dev = d_f.cpu().numpy().T.copy()

    # ---- end-to-end arm: host buffers through the reference-facing call ----------------------
    for _ in range(max(1, args.warmup)):
        eng.build_fock_df(h_np, d_np, c_np, n_occ, k_scale=k_scale, out=f_np)
    barrier()
    t0 = time.perf_counter()
    e0.record(stream)
    for _ in range(args.steps):
        eng.build_fock_df(h_np, d_np, c_np, n_occ, k_scale=k_scale, out=f_np)
    e1.record(stream)
    e1.synchronize()
    wall_ms = 1e3 * (time.perf_counter() - t0)
    barrier()
    e2e_ms = max_over_ranks(max(e0.elapsed_time(e1), wall_ms)) / args.steps
    h2d = 8 * (2 * n * n + n * n_occ)
    d2h = 8 * n * n
    e2e_equal = bool(np.array_equal(np.asarray(f_np), fock_dev))

    # ---- rooflines ----------------------------------------------------------------------------
    hbm_peak, hbm_src = _peaks()
    npair = n * (n + 1) // 2
    steps = args.steps
    t_k1 = phase["k_half_transform"] / steps * 1e-3
    t_k2 = phase["k_accumulate"] / steps * 1e-3
    t_j1 = phase["j_gamma"] / steps * 1e-3
    t_j2 = phase["j_accumulate"] / steps * 1e-3
    fl_k1 = 2.0 * n * n * n_occ * q_count            # half-transform, per launch (this rank's shard)
    fl_k2 = 1.0 * n * n * n_occ * q_count            # SYRK-form accumulation
    by_j = 8.0 * npair * q_count                     # one pass over the packed shard

    # live FP64 peak: cuBLAS DGEMM (its B200 kernel is DMMA.8x8x4 too), best of 5
    a = torch.randn(8192, 8192, dtype=torch.float64, device="cuda")
    bmat = torch.randn(8192, 8192, dtype=torch.float64, device="cuda")
    best = 1e9
    for i in range(6):
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record(); a @ bmat; s1.record(); torch.cuda.synchronize()
        if i:
            best = min(best, s0.elapsed_time(s1))
    fp64_peak = 2 * 8192 ** 3 / best * 1e-9
    del a, bmat

    def kern(name, bound, work, t, peak, unit, scale_):
        ach = work / t * scale_ if t > 0 else 0.0
        return {"kernel": name, "bound": bound, "achieved": ach, "peak": peak, "unit": unit,
                "frac": ach / peak if peak else None, "ms_per_launch": t * 1e3, "traffic": None}

    kernels = [
        kern("k_half_transform_kernel", "tensor", fl_k1, t_k1, fp64_peak, "TFLOP/s", 1e-12),
        kern("k_accumulate_kernel", "tensor", fl_k2, t_k2, fp64_peak, "TFLOP/s", 1e-12),
        kern("j_gamma_kernel", "hbm", by_j, t_j1, hbm_peak, "GB/s", 1e-9),
        kern("j_accumulate_kernel", "hbm", by_j, t_j2, hbm_peak, "GB/s", 1e-9),
    ]
    roofline = dict(kernels[0])
    roofline["peak_source"] = "measured live: cuBLAS DGEMM 8192^3 best of 5 (FP64 is not in MEASURED_PEAKS.json)"
    k_total = t_k1 + t_k2
    j_total = t_j1 + t_j2
    summary = {
        "K_tflops": (fl_k1 + fl_k2) / k_total * 1e-12 if k_total > 0 else 0.0,
        "K_frac_of_fp64_peak": (fl_k1 + fl_k2) / k_total * 1e-12 / fp64_peak if k_total > 0 else 0.0,
        "J_gbs": 2 * by_j / j_total * 1e-9 if j_total > 0 else 0.0,
        "J_frac_of_hbm_peak": 2 * by_j / j_total * 1e-9 / hbm_peak if j_total > 0 else 0.0,
        "fp64_peak_tflops": fp64_peak, "hbm_peak_gbs": hbm_peak, "hbm_peak_source": hbm_src,
        "phase_ms_per_build": {k: v / steps for k, v in phase.items()},
        "reference_flops_as_executed_tflops": 4.0 * n * n * n_occ * q_count / k_total * 1e-12 if k_total > 0 else 0.0,
    }

    # ---- CPU baseline + in-bench parity on the sampled auxiliary range (rank 0, N=1) -----------
    cpu_baseline = None
    parity = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        t_full, t_s, qs, (b_s, h_s, d_s, c_s, f_ref) = cpu_port_time(n, n_occ, naux, args.cpu_sample, 2)
        cpu_baseline = {"value": 1.0 / t_full, "unit": "builds/s", "cores": cores, "kind": "port",
                        "sample": f"{qs} of {naux} auxiliary functions ({t_s:.2f} s), scaled by {naux}/{qs}; "
                                  "NumPy loop-for-loop port of build_fock_df on OpenBLAS"}
        with B200FockEngine(local_rank) as chk:
            chk.synth_tensor(n, naux, SEED, scale, q_begin=0, q_count=qs)
            f_gpu = chk.build_fock_df(h_s, d_s, c_s, n_occ)
        parity = {"max_abs_err_vs_oracle": float(np.max(np.abs(f_gpu - f_ref))), "sample_naux": qs,
                  "tolerance": 1e-10}

    if world > 1:
        dist.barrier()
    if rank == 0:
        line = {
            "metric": "DF-J/K Fock builds/sec", "value": value, "unit": "builds/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": _config(args, cfg),
            "e2e": {"value": 1e3 / e2e_ms, "unit": "builds/s", "ms_per_step": e2e_ms,
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "bit_identical_to_device_arm": e2e_equal},
            "gpu_launches": launches_per_build * args.steps,
            "roofline": roofline, "kernels": kernels, "summary": summary,
            "cpu_baseline": cpu_baseline, "parity": parity, "clocks": clocks,
            "tensor_setup_s": t_synth,
        }
        print(json.dumps(line))
    eng.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
