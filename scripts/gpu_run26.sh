#!/bin/bash
mkdir -p gpurun_out
V=$PWD/maxsquareloss_b200/lib/variants
for n in 2; do
for v in base; do echo "== $v batch $n"; AB_N=$n MSQ_B200_LIB=$V/libmsq_$v.so timeout 300 python scripts/ab_queue.py 2>&1 | grep "iters  4000 unthrottled"; done
done
AB_STEPS=3000 MSQ_B200_LIB=$V/libmsq_trace.so timeout 300 python scripts/trace_step.py 2>&1 | tee gpurun_out/r02_trace_two_pretouch.txt | tail -32
timeout 1200 python -m pytest tests -m gpu -x -q --timeout 300 2>&1 | tail -3
