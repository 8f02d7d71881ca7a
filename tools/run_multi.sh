#!/bin/bash
# Multi-GPU measurement pass on one 8-GPU box (development tool; results -> gpurun_out/).
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
python -m pytest tests/test_gpu_multirank.py -x -q 2>&1 | tail -2
run() { # name nproc port args...
  local name=$1 np=$2 port=$3; shift 3
  $TR --nproc-per-node $np --master-port $port bench.py --gpus $np "$@" > gpurun_out/b_$name.json 2> gpurun_out/b_$name.err
  python - "$name" <<'PY'
import json, sys
name = sys.argv[1]
try:
    d = json.load(open(f"gpurun_out/b_{name}.json"))
except Exception as e:
    print(name, "FAILED", e); print(open(f"gpurun_out/b_{name}.err").read()[-1500:]); sys.exit(0)
print(name, round(d["value"], 3), d["unit"], "ms", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"], 3))
s = d.get("summary")
if s:
    print("   K", round(s["K_tflops"], 2), round(s["K_frac_of_fp64_peak"], 3), "J", round(s["J_gbs"], 1),
          round(s["J_frac_of_hbm_peak"], 3), {k: round(v, 4) for k, v in s["phase_ms_per_build"].items()})
PY
}
run c2_n8 8 29601
run c4_n8 8 29602 --workload c4
run c4_n4 4 29603 --workload c4
run c4_n2 2 29604 --workload c4
run c3_n8 8 29605 --workload c3
run c5_n8 8 29606 --workload c5
run c2_n4 4 29607
run c2_n2 2 29608
