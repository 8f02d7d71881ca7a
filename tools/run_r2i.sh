#!/bin/bash
# Three-class split of the accumulation: GPU suite, then A/B of the planner's choices at c2 and c4.
mkdir -p gpurun_out; rm -f gpurun_out/r2i_ab.log
timeout 200 python -m pytest tests -m gpu -x -q > gpurun_out/r2i_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2i_pytest.log; tail -4 gpurun_out/r2i_pytest.log
run() {  # label, shape..., env assignments come through the environment
  echo "## $1" >> gpurun_out/r2i_ab.log; shift
  timeout 100 python tools/phase_times.py "$@" >> gpurun_out/r2i_ab.log 2>&1
}
run "c2 default (edge class, planner's splits)" 688 80 1800 40
MQCB200_EDGE_SPLITS=1 MQCB200_KSPLITS=15 run "c2 edge class, 15 splits" 688 80 1800 40
MQCB200_NO_EDGE_SPLITS=1 run "c2 two classes (previous schedule)" 688 80 1800 40
run "c4 default (edge class, planner's splits)" 1450 241 6800 4
MQCB200_EDGE_SPLITS=1 MQCB200_KSPLITS=25 run "c4 edge class, 25 splits" 1450 241 6800 4
MQCB200_NO_EDGE_SPLITS=1 run "c4 two classes (previous schedule)" 1450 241 6800 4
python - <<'PY'
import json
lab = None
for line in open("gpurun_out/r2i_ab.log"):
    if line.startswith("##"): lab = line[3:].strip(); continue
    try: d = json.loads(line)
    except Exception: print(lab, "??", line[:200]); continue
    print(f"{lab:45s} {d['ms_per_build']:9.4f} ms  K2 {d['phases_ms']['k_accumulate']:8.4f}  K1 {d['phases_ms']['k_half_transform']:8.4f}  sha {d['fock_sha']}")
PY
