#!/bin/bash
# round-2 GPU call 4: GPU tests with the C++ torch binding, host-cost profile, HBM-kernel timings, ncu --set full of the HBM kernels
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q --timeout 120 > gpurun_out/r02_pytest_gpu_c.log 2>&1
tail -8 gpurun_out/r02_pytest_gpu_c.log
timeout 300 python scripts/prof_host.py > gpurun_out/r02_prof_host_c.log 2>&1
head -4 gpurun_out/r02_prof_host_c.log; tail -2 gpurun_out/r02_prof_host_c.log
HBM_REPS=10 timeout 300 python scripts/hbm_kernels.py > gpurun_out/r02_hbm_kernels.log 2>&1
cat gpurun_out/r02_hbm_kernels.log
HBM_REPS=1 timeout 900 ncu --set full --clock-control none -k regex:"prob_|confusion_|softce_" -c 45 -o /tmp/r02_hbm python scripts/hbm_kernels.py > gpurun_out/r02_hbm_ncu.log 2>&1
tail -3 gpurun_out/r02_hbm_ncu.log
ncu -i /tmp/r02_hbm.ncu-rep --page raw --csv > gpurun_out/r02_hbm_raw.csv 2>/dev/null
python scripts/ncu_summary.py /tmp/r02_hbm.ncu-rep "ncu --set full --clock-control none, scripts/hbm_kernels.py (HBM_REPS=1): the HBM-bound kernels, cold L2" > gpurun_out/r02_ncu_hbm_summary.txt 2>&1
ls -la gpurun_out/ /tmp/r02_hbm.ncu-rep
