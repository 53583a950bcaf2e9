#!/bin/bash
mkdir -p gpurun_out
L=$PWD/maxsquareloss_b200/lib
{
timeout 300 python -m pytest tests/test_gpu_loss.py tests/test_gpu_edge.py -m gpu -x -q --timeout 120 2>&1 | tail -3
for n in 1 2 4; do AB_N=$n python scripts/ab_fused.py; done
for n in 1 2 4; do AB_N=$n MSQ_B200_LIB=$L/libmsq_nopair.so python scripts/ab_fused.py; done
for n in 2; do AB_N=$n python scripts/ab_fused.py; AB_N=$n MSQ_B200_LIB=$L/libmsq_nopair.so python scripts/ab_fused.py; done
python scripts/ab_hbm.py 2>&1 | head -4
} > gpurun_out/r02_ab_pair.log 2>&1
cat gpurun_out/r02_ab_pair.log
