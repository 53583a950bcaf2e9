#!/bin/bash
# round-2 GPU call 1: the three A/Bs round 1 left unrun, the host-cost profile of the module path, a sanitizer pass
mkdir -p gpurun_out
L=maxsquareloss_b200/lib
{
for n in 1 2 4; do AB_N=$n python scripts/ab_fused.py; done
for n in 1 2 4; do AB_N=$n MSQ_B200_LIB=$PWD/$L/libmsq_tw64.so python scripts/ab_fused.py; done
for n in 2; do AB_N=$n MSQ_B200_LIB=$PWD/$L/libmsq_ring4.so python scripts/ab_fused.py; done
} > gpurun_out/r02_ab_fused.log 2>&1
python scripts/ab_conf.py > gpurun_out/r02_ab_conf.log 2>&1
python scripts/prof_host.py > gpurun_out/r02_prof_host.log 2>&1
timeout 600 compute-sanitizer --tool memcheck python scripts/sanitize_small.py > gpurun_out/r02_sanitizer_memcheck.log 2>&1
tail -5 gpurun_out/r02_sanitizer_memcheck.log
cat gpurun_out/r02_ab_fused.log gpurun_out/r02_ab_conf.log
head -30 gpurun_out/r02_prof_host.log
