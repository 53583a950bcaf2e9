#!/bin/bash
mkdir -p gpurun_out
V=$PWD/maxsquareloss_b200/lib/variants
for n in 2 4 1; do echo "== batch $n"; AB_QUICK=1 AB_N=$n timeout 120 python scripts/ab_queue.py 2>&1 | tail -3; done
AB_N=2 timeout 120 python scripts/ab_fused.py 2>&1 | tail -1
timeout 900 python -m pytest tests -m gpu -x -q --timeout 120 -k "fused or multi or source or entropy or step or loss or hard or edge" 2>&1 | tail -3
