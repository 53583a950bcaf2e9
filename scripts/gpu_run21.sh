#!/bin/bash
mkdir -p gpurun_out
for m in 15 14 13 11 10 8; do AB_PDL=$m AB_N=2 timeout 300 python scripts/ab_fused.py 2>&1 | tail -1; done
