#!/bin/bash
# round-2 GPU call: flag hand-off of the weights, backward set-up before the wait, L2 prefetch of the forward's tile
mkdir -p gpurun_out
MSQ_B200_LIB=$PWD/maxsquareloss_b200/lib/variants/libmsq_trace.so timeout 300 python scripts/trace_step.py 2>&1 | tee gpurun_out/r02_trace_step_b.txt
for n in 2 1 4; do AB_N=$n timeout 400 python scripts/ab_variants.py run 2>&1 | grep "^libmsq_base"; done
timeout 1200 python -m pytest tests -m gpu -x -q --timeout 300 2>&1 | tail -3
