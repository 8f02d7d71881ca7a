#!/bin/bash
# What the driver does at round end on one GPU: GPU tests, smoke, both bench arms with its flags.
mkdir -p gpurun_out
if [ "$1" != "bench-only" ]; then
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02z_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02z_pytest.log; tail -4 gpurun_out/r02z_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02z_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/r02z_smoke.log
fi
T0=$SECONDS; timeout 1200 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02z_ref_n1.json 2> gpurun_out/r02z_ref_n1.err; echo "ref rc=$? wall=$((SECONDS-T0))s"
T0=$SECONDS; timeout 1800 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02z_bench_n1.json 2> gpurun_out/r02z_bench_n1.err; echo "bench rc=$? wall=$((SECONDS-T0))s"
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02z_bench_n1.json")); r = json.load(open("gpurun_out/r02z_ref_n1.json"))
print("value", d["value"], "e2e", d["e2e"]["value"], "ref", r["value"], "ratio e2e", d["e2e"]["value"] / r["value"], "ref ms_per_step", r["ms_per_step"])
print("parity", d["parity"]["ok"], d["parity"]["max_abs_err_vs_oracle"], "clocks", d["clocks"], "launches", d["gpu_launches"], "wall", d["bench_wall_s"])
for k, v in d["workloads"].items():
    print(k, round(v["value"], 2), round(v["e2e"]["value"], 2), (v.get("parity") or {}).get("ok"))
print(json.dumps(d["workloads"]["c3"]["device_resident_scf"]["batched"]))
print(d["roofline"])
PY
