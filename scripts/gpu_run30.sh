#!/bin/bash
# round-2 GPU call (2 GPUs): the 2-rank tests and the N=2 bench with the two-kernel step (finalisation + mailbox exchange in extra CTAs of the backward)
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 600 python -m pytest tests/test_gpu_dist.py tests/test_gpu_loss.py -m gpu -x -q --timeout 300 2>&1 | tail -3
timeout 900 $TR --nproc-per-node 2 --master-port 29552 bench.py --gpus 2 > gpurun_out/r02_bench_n2_d.json 2> gpurun_out/r02_bench_n2_d.err; echo "bench2 rc $?"; tail -2 gpurun_out/r02_bench_n2_d.err
python - <<'PY'
import json
for n in (2,):
    try:
        d = json.load(open(f"gpurun_out/r02_bench_n{n}_d.json"))
        print(n, {k: d[k] for k in ("value", "ms_per_step", "gpu_launches")}, "e2e", round(d["e2e"]["value"], 2), "pipe", round(d["e2e"]["c_abi_pipeline"]["value"], 2))
        c3 = d["cfg3_multi_level"]; print("  cfg3", round(c3["us_per_step"], 1), round(c3["value"], 1), c3.get("check", {}).get("ok"))
        print("  stats", d["stats_check"], "spread", d["rank_spread"])
        print("  cfg5", d["cfg5_crosscity"]["fused_ms"], d["cfg5_crosscity"].get("check", {}).get("ok"))
    except Exception as e:
        print(n, "parse failed", e)
PY
