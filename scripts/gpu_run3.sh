#!/bin/bash
# round-2 GPU call 3: full GPU test suite (per-test time-out), two-kernel step A/B, host-cost profile, HBM-kernel timings + ncu
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --timeout 120 > gpurun_out/r02_pytest_gpu_b.log 2>&1
tail -15 gpurun_out/r02_pytest_gpu_b.log
timeout 300 python scripts/ab_step.py > gpurun_out/r02_ab_step.log 2>&1
cat gpurun_out/r02_ab_step.log
timeout 300 python scripts/hbm_kernels.py > gpurun_out/r02_hbm_kernels.log 2>&1
cat gpurun_out/r02_hbm_kernels.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"prob_|confusion_|softce_" -c 60 -o gpurun_out/r02_hbm_kernels python scripts/hbm_kernels.py > gpurun_out/r02_hbm_ncu.log 2>&1
tail -3 gpurun_out/r02_hbm_ncu.log
ls -la gpurun_out/*.ncu-rep
