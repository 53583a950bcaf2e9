#!/bin/bash
# round-2 GPU call: source-level ncu of the fused forward / backward at the bench shape (per-instruction executed counts)
mkdir -p gpurun_out
timeout 300 python scripts/ab_fused.py 2>&1 | tail -2
timeout 400 ncu --set full --import-source on --clock-control none -k regex:fused_fwd_kernel -s 40 -c 1 -f -o gpurun_out/r02_src_fwd python scripts/ab_fused.py > gpurun_out/r02_src_fwd.log 2>&1; echo "ncu fwd rc $?"
timeout 400 ncu --set full --import-source on --clock-control none -k regex:fused_bwd_kernel -s 40 -c 1 -f -o gpurun_out/r02_src_bwd python scripts/ab_fused.py > gpurun_out/r02_src_bwd.log 2>&1; echo "ncu bwd rc $?"
ls -la gpurun_out/*.ncu-rep
