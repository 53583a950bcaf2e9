#!/bin/bash
# round-2 GPU call (8 GPUs): 2-rank tests, bench at N=8 (cfg-3 with the same-step mailbox exchange), N=2; deferred Eval at N=1
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 600 python -m pytest tests/test_gpu_dist.py tests/test_gpu_perimage.py tests/test_gpu_eval.py -m gpu -x -q --timeout 300 2>&1 | tail -3
timeout 900 $TR --nproc-per-node 8 --master-port 29551 bench.py --gpus 8 > gpurun_out/r02_bench_n8_c.json 2> gpurun_out/r02_bench_n8_c.err; echo "bench8 rc $?"; tail -2 gpurun_out/r02_bench_n8_c.err
timeout 900 $TR --nproc-per-node 4 --master-port 29554 bench.py --gpus 4 > gpurun_out/r02_bench_n4_c.json 2> gpurun_out/r02_bench_n4_c.err; echo "bench4 rc $?"
timeout 900 $TR --nproc-per-node 2 --master-port 29552 bench.py --gpus 2 > gpurun_out/r02_bench_n2_c.json 2> gpurun_out/r02_bench_n2_c.err; echo "bench2 rc $?"
timeout 900 python bench.py --skip-cpu > gpurun_out/r02_bench_n1_c.json 2> gpurun_out/r02_bench_n1_c.err; echo "bench1 rc $?"
python - <<'PY'
import json
for n in (8, 4, 2, 1):
    try:
        d = json.load(open(f"gpurun_out/r02_bench_n{n}_c.json"))
        print(n, {k: d[k] for k in ("value", "ms_per_step", "gpu_launches")}, "e2e", round(d["e2e"]["value"], 2), "pipe", round(d["e2e"]["c_abi_pipeline"]["value"], 2))
        c3 = d["cfg3_multi_level"]; print("  cfg3", round(c3["us_per_step"], 1), round(c3["value"], 1), c3.get("check", {}).get("ok"), c3.get("exchange"))
        ch = d["confusion_hist"]
        print("  conf", {k: (round(v["value"], 1), round(v["frac_of_hbm_aggregate"], 3), round(v["us_per_image_per_rank"], 2)) for k, v in ch.items() if isinstance(v, dict)})
        if "stats_check" in d: print("  stats ok", d["stats_check"]["ok"], "cfg5", d["cfg5_crosscity"]["fused_ms"], d["cfg5_crosscity"].get("check", {}).get("ok"))
    except Exception as e:
        print(n, "parse failed", e)
PY
