#!/bin/bash
# Round-2 multi-GPU pass (N = number of GPUs on the box): multirank tests + the default bench line under torchrun.
N=${1:-2}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multirank.py -x -q > gpurun_out/r2b_pytest_n$N.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2b_pytest_n$N.log
tail -12 gpurun_out/r2b_pytest_n$N.log
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29711 \
  bench.py --gpus $N --steps ${STEPS:-20} --warmup ${WARMUP:-5} > gpurun_out/r2b_bench_n$N.json 2> gpurun_out/r2b_bench_n$N.err; echo "bench rc=$?"
grep "\[bench\]" gpurun_out/r2b_bench_n$N.err | tail -20
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29712 \
  bench.py --impl reference --gpus $N --steps 3 --warmup 1 > gpurun_out/r2b_ref_n$N.json 2> gpurun_out/r2b_ref_n$N.err; echo "ref rc=$?"
python - $N <<'PY'
import json, sys
N = sys.argv[1]
try:
    d = json.load(open(f"gpurun_out/r2b_bench_n{N}.json"))
except Exception as e:
    print("bench FAILED", e); print(open(f"gpurun_out/r2b_bench_n{N}.err").read()[-3000:]); raise SystemExit
def show(name, d):
    s = d.get("summary") or {}
    print(name, round(d["value"], 3), "builds/s ms", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"], 3),
          "pageable", (d.get("e2e_pageable") or {}).get("value"), "parity", d.get("parity"))
    if s:
        print("   K", round(s["K_tflops"], 2), round(s["K_frac_of_fp64_peak"], 3), "J", round(s["J_gbs"], 1), round(s["J_frac_of_hbm_peak"], 3),
              "ms", s["ms_per_build"], "fused", s["gamma_fused_in_timed_builds"])
        print("   serial", {k: round(v, 4) for k, v in s["phase_ms_per_build_one_stream"].items()})
show("head", d)
for k, v in d.get("workloads", {}).items():
    show(k, v)
print("wall", d.get("bench_wall_s"))
r = json.load(open(f"gpurun_out/r2b_ref_n{N}.json")); print("ref", r["value"], r["cpu_baseline"])
PY
