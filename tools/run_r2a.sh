#!/bin/bash
# Round-2 single-GPU pass: GPU tests, the default bench line, the reference arm, a launch list.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/r2a_smi.txt 2>&1
nproc >> gpurun_out/r2a_smi.txt; free -g >> gpurun_out/r2a_smi.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2a_pytest.log
tail -15 gpurun_out/r2a_pytest.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; echo "bench rc=$?"
tail -30 gpurun_out/r2a_bench.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/r2a_ref.json 2> gpurun_out/r2a_ref.err; echo "ref rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2a_launches_c2.csv \
  python bench.py --workload c2 --secondary none --steps 2 --warmup 1 --no-cpu-baseline --no-parity --no-setup-timings > gpurun_out/r2a_ncu.log 2>&1; echo "ncu rc=$?"
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/r2a_bench.json"))
except Exception as e:
    print("bench FAILED", e); raise SystemExit
def show(name, d):
    s = d.get("summary") or {}
    print(name, round(d["value"], 3), "builds/s ms", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"], 3),
          "pageable", (d.get("e2e_pageable") or {}).get("value"), "parity", (d.get("parity") or {}).get("max_abs_err_vs_oracle"),
          "cpu", (d.get("cpu_baseline") or {}))
    if s:
        print("   K", round(s["K_tflops"], 2), round(s["K_frac_of_fp64_peak"], 3), "J", round(s["J_gbs"], 1), round(s["J_frac_of_hbm_peak"], 3),
              "ms", s["ms_per_build"], "fused", s["gamma_fused_in_timed_builds"])
        print("   serial", {k: round(v, 4) for k, v in s["phase_ms_per_build_one_stream"].items()})
        print("   2strm ", {k: round(v, 4) for k, v in s["phase_ms_per_build_two_streams"].items()})
    if d.get("setup_timings"): print("   setup", d["setup_timings"])
show("head", d)
for k, v in d.get("workloads", {}).items():
    show(k, v)
print("wall", d.get("bench_wall_s"))
PY
