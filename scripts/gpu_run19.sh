#!/bin/bash
mkdir -p gpurun_out
V=$PWD/maxsquareloss_b200/lib/variants
AB_ONECALL=1 MSQ_B200_LIB=$V/libmsq_trace.so timeout 300 python scripts/trace_step.py 2>&1 | tee gpurun_out/r02_trace_step_f.txt
AB_ONECALL=0 MSQ_B200_LIB=$V/libmsq_trace.so timeout 300 python scripts/trace_step.py 2>&1 | tee gpurun_out/r02_trace_step_g.txt
AB_ONECALL=1 AB_STEPS=201 MSQ_B200_LIB=$V/libmsq_trace.so timeout 300 python scripts/trace_step.py 2>&1 | tee gpurun_out/r02_trace_step_h.txt
