#!/bin/bash
# repeat the exchange-timeout recovery scenario of tests/test_gpu_multirank.py (development)
python - <<'PY'
import re
src = open("tests/test_gpu_multirank.py").read()
m = re.search(r'RECOVERY_WORKER = r"""(.*?)"""', src, re.S)
open("/tmp/recovery_worker.py", "w").write(m.group(1))
PY
for i in $(seq 1 ${1:-8}); do
  MQC_ROOT=$PWD MQCB200_XGPU_TIMEOUT_S=2 timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 \
    --master-addr 127.0.0.1 --master-port $((29800+i)) /tmp/recovery_worker.py > /tmp/rw_$i.log 2>&1
  echo "iter $i rc=$? $(grep -h '^rank' /tmp/rw_$i.log | tr '\n' '|')"
  if ! grep -q "rank 0.*repeat=True" /tmp/rw_$i.log || ! grep -q "rank 1.*repeat=True" /tmp/rw_$i.log; then tail -30 /tmp/rw_$i.log; fi
done
