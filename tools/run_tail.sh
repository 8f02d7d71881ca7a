#!/bin/bash
# half-transform: trimmed last block on/off (development)
for shape in "1450 241 850" "1450 241 96" "976 120 300" "400 150 200"; do
  for off in "" 1; do
    if [ -n "$off" ]; then export MQCB200_NO_TRIM=1; else unset MQCB200_NO_TRIM; fi
    echo "$shape notrim=$off $(python tools/phase_times.py $shape 10 2>&1 | tail -1 | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print(d["phases_ms"]["k_half_transform"], d["phases_ms"]["k_accumulate"], d["ms_per_build"], d["k1_tflops"], d["fock_sha"])')"
  done
done
unset MQCB200_NO_TRIM
python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -x -q -m gpu 2>&1 | tail -3
