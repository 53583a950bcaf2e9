#!/bin/bash
# round-2 GPU call: finalisation with one load round trip; timeline + timing + parity
mkdir -p gpurun_out
V=$PWD/maxsquareloss_b200/lib/variants
MSQ_B200_LIB=$V/libmsq_trace.so timeout 300 python scripts/trace_step.py 2>&1 | tee gpurun_out/r02_trace_step_e.txt
for n in 2 1 4; do AB_N=$n timeout 400 python scripts/ab_variants.py run 2>&1 | grep "^libmsq_base"; done
timeout 1200 python -m pytest tests -m gpu -x -q --timeout 300 2>&1 | tail -3
