#!/bin/bash
mkdir -p gpurun_out
V=$PWD/maxsquareloss_b200/lib/variants
for n in 2 1 4; do
for v in base spare0; do echo "== $v batch $n"; AB_N=$n MSQ_B200_LIB=$V/libmsq_$v.so timeout 300 python scripts/ab_queue.py 2>&1 | grep "iters  4000 unthrottled\|iters   300"; done
done
AB_STEPS=3000 MSQ_B200_LIB=$V/libmsq_trace.so timeout 300 python scripts/trace_step.py 2>&1 | tee gpurun_out/r02_trace_two_spare.txt | tail -32
