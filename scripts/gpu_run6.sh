#!/bin/bash
# round-2 GPU call 6 (2 GPUs): the 2-process NCCL/mailbox tests, the bench at N=2
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_dist.py -m gpu -x -q --timeout 300 > gpurun_out/r02_pytest_dist.log 2>&1
tail -5 gpurun_out/r02_pytest_dist.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 > gpurun_out/r02_bench_n2_a.json 2> gpurun_out/r02_bench_n2_a.err
echo "bench rc $?"; tail -15 gpurun_out/r02_bench_n2_a.err
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/r02_bench_n2_a.json"))
    print({k: d[k] for k in ("value", "ms_per_step", "warmup", "gpu_launches")})
    print("e2e", {k: (v if not isinstance(v, dict) else {kk: vv for kk, vv in v.items() if kk != "how"}) for k, v in d["e2e"].items() if k != "how"})
    print("stats_check", d["stats_check"])
    print("rank_spread", d["rank_spread"])
    print("cfg3", {k: v for k, v in d["cfg3_multi_level"].items() if k not in ("what", "exchange")})
    print("cfg5", {k: v for k, v in d["cfg5_crosscity"].items() if k != "what"})
    ch = d["confusion_hist"]
    print("conf", {k: (round(v["value"], 1), round(v["frac_of_hbm_aggregate"], 3)) for k, v in ch.items() if isinstance(v, dict)}, ch["miou_16_13"], ch["matrix_total"])
except Exception as e:
    print("parse failed", e)
PY
