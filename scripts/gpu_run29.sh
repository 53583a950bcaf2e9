#!/bin/bash
mkdir -p gpurun_out
V=$PWD/maxsquareloss_b200/lib/variants
for n in 2 4 1; do for v in base tw256; do echo "== $v batch $n"; MSQ_B200_LIB=$V/libmsq_$v.so AB_QUICK=1 AB_N=$n timeout 120 python scripts/ab_queue.py 2>&1 | tail -3; done; done
