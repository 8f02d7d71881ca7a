#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2f_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2f_pytest.log; tail -6 gpurun_out/r2f_pytest.log
for cj in 1 0; do
  MQCB200_CORESIDENT_J=$cj timeout 600 python bench.py --steps 10 --warmup 3 --workload c2 --secondary none --no-cpu-baseline --no-setup-timings > gpurun_out/r2f_c2_cj$cj.json 2> gpurun_out/r2f_c2_cj$cj.err; echo "c2 cj=$cj rc=$?"
done
MQCB200_CORESIDENT_J=1 timeout 600 python bench.py --steps 5 --warmup 3 --secondary none --no-cpu-baseline --no-setup-timings > gpurun_out/r2f_c4_cj1.json 2> gpurun_out/r2f_c4_cj1.err; echo "c4 rc=$?"
python - <<'PY'
import json
for f in ("r2f_c2_cj1", "r2f_c2_cj0", "r2f_c4_cj1"):
    try:
        d = json.load(open(f"gpurun_out/{f}.json"))
    except Exception as e:
        print(f, "FAILED", e); print(open(f"gpurun_out/{f}.err").read()[-1500:]); continue
    s = d["summary"]
    print(f, round(d["value"], 3), "ms", s["ms_per_build"], "parity", d["parity"]["max_abs_err_vs_oracle"] if d.get("parity") else None)
    print("   2strm", {k: round(v, 4) for k, v in s["phase_ms_per_build_two_streams"].items()})
PY
