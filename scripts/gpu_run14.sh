#!/bin/bash
# round-2 GPU call: A/B of the forward SHF mask, the backward L1 prefetch and the software-pipelined forward; parity tests on the last one
mkdir -p gpurun_out
for n in 2 1 4; do echo "batch $n"; AB_N=$n timeout 400 python scripts/ab_variants.py run 2>&1 | grep "^libmsq"; done
timeout 900 python scripts/ab_variants.py test shf_pf3_pipe 2>&1 | tail -3
