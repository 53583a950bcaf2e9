#!/bin/bash
# round-2 GPU call: timeline of the step (globaltimer stamps), default build timing, full parity suite on the default build
mkdir -p gpurun_out
MSQ_B200_LIB=$PWD/maxsquareloss_b200/lib/variants/libmsq_trace.so timeout 300 python scripts/trace_step.py 2>&1 | tee gpurun_out/r02_trace_step.txt
AB_MODE=0 MSQ_B200_LIB=$PWD/maxsquareloss_b200/lib/variants/libmsq_trace.so timeout 300 python scripts/trace_step.py 2>&1 | tee gpurun_out/r02_trace_step_ms.txt
AB_N=2 timeout 400 python scripts/ab_variants.py run 2>&1 | grep "^libmsq"
timeout 1200 python -m pytest tests -m gpu -x -q --timeout 300 2>&1 | tail -3
