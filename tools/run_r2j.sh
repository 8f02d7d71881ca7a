#!/bin/bash
# Final state of the round on one GPU: GPU tests, smoke, the driver's bench line.
mkdir -p gpurun_out
timeout 120 python -m pytest tests -m gpu -x -q > gpurun_out/r2j_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2j_pytest.log; tail -3 gpurun_out/r2j_pytest.log
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2j_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/r2j_smoke.log
T0=$SECONDS; timeout 110 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2j_bench_n1.json 2> gpurun_out/r2j_bench_n1.err; echo "bench rc=$? wall=$((SECONDS-T0))s"
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2j_bench_n1.json"))
print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "roofline", d["roofline"]["frac"], "parity", d["parity"]["ok"], d["parity"]["max_abs_err_vs_oracle"], "clocks", d["clocks"])
for k, v in d["workloads"].items():
    print(k, round(v["value"], 2), round(v.get("ms_per_step", 0), 4), round(v["e2e"]["value"], 2), (v.get("parity") or {}).get("ok"))
PY
