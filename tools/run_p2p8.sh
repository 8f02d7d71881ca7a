#!/bin/bash
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --nproc-per-node 8"
show() { python -c "
import json,sys
try:
    d=json.loads(open('$2').read().strip().splitlines()[-1])
    print('$1', round(d['value'],2), 'e2e', round(d['e2e']['value'],2), {k: round(v,4) for k,v in d['summary']['phase_ms_per_build'].items()})
except Exception as e:
    print('$1 FAILED', e); print(open('$3').read()[-1500:])
"; }
timeout 240 $TR --master-port 29711 bench.py --gpus 8 --no-cpu-baseline > gpurun_out/p8_c2_p2p.json 2> gpurun_out/p8_c2_p2p.err; show "c2 P2P " gpurun_out/p8_c2_p2p.json gpurun_out/p8_c2_p2p.err
MQCB200_P2P_ALLREDUCE=0 timeout 240 $TR --master-port 29712 bench.py --gpus 8 --no-cpu-baseline > gpurun_out/p8_c2_nccl.json 2> gpurun_out/p8_c2_nccl.err; show "c2 NCCL" gpurun_out/p8_c2_nccl.json gpurun_out/p8_c2_nccl.err
timeout 240 $TR --master-port 29713 bench.py --gpus 8 --no-cpu-baseline --workload c4 > gpurun_out/p8_c4_p2p.json 2> gpurun_out/p8_c4_p2p.err; show "c4 P2P " gpurun_out/p8_c4_p2p.json gpurun_out/p8_c4_p2p.err
timeout 240 $TR --master-port 29714 bench.py --gpus 8 --no-cpu-baseline --workload c5 > gpurun_out/p8_c5_p2p.json 2> gpurun_out/p8_c5_p2p.err; show "c5 P2P " gpurun_out/p8_c5_p2p.json gpurun_out/p8_c5_p2p.err
