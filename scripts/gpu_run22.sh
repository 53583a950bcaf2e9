#!/bin/bash
mkdir -p gpurun_out
V=$PWD/maxsquareloss_b200/lib/variants
AB_ONECALL=1 MSQ_B200_LIB=$V/libmsq_trace.so timeout 300 python scripts/trace_step.py 2>&1 | tee gpurun_out/r02_trace_two_onecall.txt
AB_ONECALL=0 MSQ_B200_LIB=$V/libmsq_trace.so timeout 300 python scripts/trace_step.py 2>&1 | tee gpurun_out/r02_trace_two_separate.txt
AB_N=2 timeout 600 python scripts/ab_variants.py run 2>&1 | grep "^libmsq" | grep -v trace
