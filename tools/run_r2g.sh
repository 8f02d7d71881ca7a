#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2g_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2g_pytest.log; tail -4 gpurun_out/r2g_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2g_smoke.log 2>&1; echo "smoke rc=$?"; cat gpurun_out/r2g_smoke.log
python tools/prof_extras.py > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none -k regex:"whiten_gemm|dgemm_batched|scf_step|scf_orthogonalizer" -c 12 \
  -o gpurun_out/r02_prof_extras -f python tools/prof_extras.py > gpurun_out/r02_ncu_e.log 2>&1; echo "ncu extras rc=$?"
python tools/ncu_summary.py gpurun_out/r02_prof_extras.ncu-rep > gpurun_out/r02_ncu_summary_extras.md 2>/dev/null
rm -f gpurun_out/r02_prof_extras.ncu-rep
grep -c "^## " gpurun_out/r02_ncu_summary_extras.md
