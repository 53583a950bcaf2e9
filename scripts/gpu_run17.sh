#!/bin/bash
# round-2 GPU call: where the finalisation kernel's time goes; polling interval of the flag hand-off
mkdir -p gpurun_out
V=$PWD/maxsquareloss_b200/lib/variants
AB_ONECALL=0 MSQ_B200_LIB=$V/libmsq_trace.so timeout 300 python scripts/trace_step.py 2>&1 | tee gpurun_out/r02_trace_step_c.txt
MSQ_B200_LIB=$V/libmsq_trace1500.so timeout 300 python scripts/trace_step.py 2>&1 | tee gpurun_out/r02_trace_step_d.txt
AB_N=2 timeout 400 python scripts/ab_variants.py run 2>&1 | grep "^libmsq"
