#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q --timeout 120 > gpurun_out/r02_pytest_gpu_e.log 2>&1
tail -4 gpurun_out/r02_pytest_gpu_e.log
timeout 600 python scripts/ab_hbm.py > gpurun_out/r02_ab_hbm.log 2>&1
cat gpurun_out/r02_ab_hbm.log
