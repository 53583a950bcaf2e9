#!/bin/bash
mkdir -p gpurun_out
V=$PWD/maxsquareloss_b200/lib/variants
for n in 2 1 4; do for late in 1 0; do echo "== batch $n late $late"; AB_QUICK=1 AB_LATE=$late AB_N=$n timeout 120 python scripts/ab_queue.py 2>&1 | tail -3; done; done
AB_STEPS=3000 MSQ_B200_LIB=$V/libmsq_trace.so timeout 120 python scripts/trace_step.py 2>&1 | tee gpurun_out/r02_trace_two_merged.txt | tail -32
timeout 1200 python -m pytest tests -m gpu -x -q --timeout 120 2>&1 | tail -3
