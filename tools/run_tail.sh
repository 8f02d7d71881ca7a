#!/bin/bash
# SYRK row trim: compare with the previous run's numbers (development)
for shape in "1450 241 850" "976 120 300" "688 80 1800" "688 80 225" "130 33 97" "200 57 300"; do
  echo "$shape $(python tools/phase_times.py $shape 10 2>&1 | tail -1 | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print(d["phases_ms"]["k_half_transform"], d["phases_ms"]["k_accumulate"], d["ms_per_build"], d["k2_tflops_alg"], d["fock_sha"])')"
done
python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -x -q -m gpu 2>&1 | tail -3
