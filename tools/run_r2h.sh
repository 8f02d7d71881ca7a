#!/bin/bash
# Edge tiles of the accumulation (consume_edge): GPU suite with the new switch in the bit-identity tests, then A/B.
mkdir -p gpurun_out
timeout 240 python -m pytest tests -m gpu -x -q > gpurun_out/r2h_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2h_pytest.log; tail -4 gpurun_out/r2h_pytest.log
for shape in "688 80 1800 40" "1450 241 6800 6"; do
  for sw in "" MQCB200_NO_EDGE_TILES; do
    if [ -n "$sw" ]; then export $sw=1; fi
    echo "edge_tiles=${sw:-on}" >> gpurun_out/r2h_ab.log
    timeout 120 python tools/phase_times.py $shape >> gpurun_out/r2h_ab.log 2>&1
    if [ -n "$sw" ]; then unset $sw; fi
  done
done
cat gpurun_out/r2h_ab.log | cut -c1-400
