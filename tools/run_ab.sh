#!/bin/bash
# A/B on one box: the half-transform's block trim and tail split, through bench.py (development)
B="python bench.py --secondary none --steps 6 --warmup 3 --no-cpu-baseline --no-parity --no-setup-timings"
for w in c4 c5; do
  for sw in "" MQCB200_NO_TRIM MQCB200_NO_TAIL_SPLIT; do
    if [ -n "$sw" ]; then export $sw=1; fi
    $B --workload $w 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read()); s = d['summary']['phase_ms_per_build_one_stream']
print('$w', '${sw:-default}', round(d['value'], 4), 'builds/s', round(d['ms_per_step'], 3), 'ms  K1', round(s['k_half_transform'], 3), 'K2', round(s['k_accumulate'], 3), d['clocks']['power_w_max'], d['clocks']['reasons'])"
    if [ -n "$sw" ]; then unset $sw; fi
  done
done
