#!/bin/bash
mkdir -p gpurun_out
V=$PWD/maxsquareloss_b200/lib/variants
AB_STEPS=3000 AB_FASTHOST=1 AB_ONECALL=1 MSQ_B200_LIB=$V/libmsq_trace.so timeout 300 python scripts/trace_step.py 2>&1 | tee gpurun_out/r02_trace_two_fast.txt
AB_STEPS=3000 AB_FASTHOST=0 AB_ONECALL=1 MSQ_B200_LIB=$V/libmsq_trace.so timeout 300 python scripts/trace_step.py 2>&1 | tee gpurun_out/r02_trace_two_slow.txt
