#!/bin/bash
# round-2 GPU call (2 GPUs): all GPU tests incl. the 2-rank ones, the launch list and the ncu --set full capture of the step,
# the bench at N=1 (record) and N=2, the copy-only ceiling at 1 and 2 ranks, the reference arm
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 900 python -m pytest tests -m gpu -x -q --timeout 300 > gpurun_out/r02_pytest_gpu_f.log 2>&1
tail -5 gpurun_out/r02_pytest_gpu_f.log
timeout 900 python bench.py > gpurun_out/r02_bench_n1_b.json 2> gpurun_out/r02_bench_n1_b.err; echo "bench1 rc $?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_reference.json 2>/dev/null; echo "ref rc $?"
timeout 900 $TR --nproc-per-node 2 --master-port 29541 bench.py --gpus 2 > gpurun_out/r02_bench_n2_b.json 2> gpurun_out/r02_bench_n2_b.err; echo "bench2 rc $?"
{ python scripts/ab_copy_ranks.py; timeout 300 $TR --nproc-per-node 2 --master-port 29542 scripts/ab_copy_ranks.py 2>/dev/null | grep copy-only; } > gpurun_out/r02_copy_ranks_12.log 2>&1
cat gpurun_out/r02_copy_ranks_12.log
MSQ_BENCH_MIN_WARM_S=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 20 --warmup 3 --skip-secondary --skip-cpu > gpurun_out/r02_ncu_launches.log 2>&1; echo "ncu launches rc $?"
MSQ_BENCH_MIN_WARM_S=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"fused_|finalize" -c 12 -o /tmp/r02_fused python bench.py --steps 3 --warmup 3 --skip-secondary --skip-cpu > gpurun_out/r02_ncu_fused.log 2>&1; echo "ncu full rc $?"
ncu -i /tmp/r02_fused.ncu-rep --page raw --csv > gpurun_out/r02_fused_raw.csv 2>/dev/null
python scripts/ncu_summary.py /tmp/r02_fused.ncu-rep "ncu --set full --clock-control none --import-source on, bench.py --steps 3 --warmup 3 (MSQ_BENCH_MIN_WARM_S=0): the kernels of the fused step" > gpurun_out/r02_ncu_fused_summary.txt 2>&1
python - <<'PY'
import json
for f in ("r02_bench_n1_b", "r02_bench_n2_b"):
    try:
        d = json.load(open(f"gpurun_out/{f}.json"))
        print(f, {k: d[k] for k in ("value", "ms_per_step", "gpu_launches")}, "e2e", round(d["e2e"]["value"], 2), "sync", round(d["e2e"]["sync_every_step"]["value"], 2), "pipe", round(d["e2e"]["c_abi_pipeline"]["value"], 2), "floor_ms", round(d["e2e"]["torch_floor"]["ms_per_step"], 4))
        if "marginal_image" in d: print("  marginal", d["marginal_image"])
        if "stats_check" in d: print("  stats ok", d["stats_check"]["ok"])
    except Exception as e:
        print(f, "parse failed", e)
PY
