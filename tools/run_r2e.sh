#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python bench.py --steps 5 --warmup 3 > gpurun_out/r2e_bench.json 2> gpurun_out/r2e_bench.err; echo "bench rc=$?"
grep "\[bench\]" gpurun_out/r2e_bench.err | tail -12; tail -5 gpurun_out/r2e_bench.err | grep -v "\[bench\]"
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/r2e_bench.json"))
except Exception as e:
    print("bench FAILED", e); raise SystemExit
def show(name, d):
    s = d.get("summary") or {}
    print(name, round(d["value"], 3), "builds/s ms", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"], 3),
          "pageable", (d.get("e2e_pageable") or {}).get("value"), "parity", (d.get("parity") or {}).get("max_abs_err_vs_oracle"))
    if s:
        print("   K", round(s["K_tflops"], 2), round(s["K_frac_of_fp64_peak"], 3), "J", round(s["J_gbs"], 1), round(s["J_frac_of_hbm_peak"], 3), "ms", s["ms_per_build"])
    if d.get("setup_timings"): print("   setup", json.dumps(d["setup_timings"])[:1500])
    if d.get("device_resident_scf"): print("   scf", d["device_resident_scf"], "dev-synth", d.get("value_tensor_synthesised_on_device"))
show("head", d)
for k, v in d.get("workloads", {}).items():
    show(k, v)
print("wall", d.get("bench_wall_s"), "cpu", d.get("cpu_baseline"))
PY
