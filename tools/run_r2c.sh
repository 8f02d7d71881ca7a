#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_device_scf.py tests/test_gpu_golden_scf.py tests/test_gpu_parity.py -x -q > gpurun_out/r2c_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c_pytest.log
tail -25 gpurun_out/r2c_pytest.log
timeout 300 python tools/scf_bench.py > gpurun_out/r2c_scf_bench.log 2>&1; cat gpurun_out/r2c_scf_bench.log | tail -20
