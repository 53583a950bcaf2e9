#!/bin/bash
# round-2 GPU call (8 GPUs): bench at N=8 and N=4, the copy-only ceiling at 1/2/4/8 ranks, GPU tests of the new row
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 300 python -m pytest tests/test_gpu_hard.py tests/test_gpu_multi.py -m gpu -x -q --timeout 120 2>&1 | tail -3
timeout 900 $TR --nproc-per-node 8 --master-port 29521 bench.py --gpus 8 > gpurun_out/r02_bench_n8_a.json 2> gpurun_out/r02_bench_n8_a.err
echo "bench8 rc $?"; tail -3 gpurun_out/r02_bench_n8_a.err
{
python scripts/ab_copy_ranks.py
for n in 2 4 8; do timeout 300 $TR --nproc-per-node $n --master-port 2953$n scripts/ab_copy_ranks.py 2>/dev/null | grep copy-only; done
} > gpurun_out/r02_copy_ranks.log 2>&1
cat gpurun_out/r02_copy_ranks.log
timeout 900 $TR --nproc-per-node 4 --master-port 29524 bench.py --gpus 4 > gpurun_out/r02_bench_n4_a.json 2> gpurun_out/r02_bench_n4_a.err
echo "bench4 rc $?"; tail -3 gpurun_out/r02_bench_n4_a.err
python - <<'PY'
import json
for n in (8, 4):
    try:
        d = json.load(open(f"gpurun_out/r02_bench_n{n}_a.json"))
        print(n, {k: d[k] for k in ("value", "ms_per_step", "gpu_launches")}, "e2e", round(d["e2e"]["value"], 2), "sync", round(d["e2e"]["sync_every_step"]["value"], 2), "pipe", round(d["e2e"]["c_abi_pipeline"]["value"], 2), "floor_ms", d["e2e"]["torch_floor"]["ms_per_step"])
        print("  stats_check ok", d["stats_check"]["ok"], d["stats_check"]["exchange"], "spread", d["rank_spread"])
        print("  cfg3", d["cfg3_multi_level"]["us_per_step"], d["cfg3_multi_level"]["value"], d["cfg3_multi_level"].get("check", {}).get("ok"))
        print("  cfg5", {k: v for k, v in d["cfg5_crosscity"].items() if k != "what"})
        ch = d["confusion_hist"]
        print("  conf", {k: (round(v["value"], 1), round(v["frac_of_hbm_aggregate"], 3)) for k, v in ch.items() if isinstance(v, dict)}, ch["miou_16_13"], ch["matrix_total"])
    except Exception as e:
        print(n, "parse failed", e)
PY
