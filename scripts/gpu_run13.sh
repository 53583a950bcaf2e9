#!/bin/bash
# round-2 GPU call: forward with CTA-level class accumulators (shared atomics) -- parity tests + A/B timing (base vs SHF mask)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --timeout 300 -k "fused or multi or source or entropy or step or loss or hard or edge" 2>&1 | tail -4
for n in 1 2 4; do AB_N=$n timeout 300 python scripts/ab_variants.py run 2>&1 | grep -v "^$" | tail -2; done
