#!/bin/bash
# round-2 GPU call (N GPUs given as $1): the bench at N ranks with the two-kernel step
N=$1
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 900 $TR --nproc-per-node $N --master-port 2956$N bench.py --gpus $N > gpurun_out/r02_bench_n${N}_d.json 2> gpurun_out/r02_bench_n${N}_d.err; echo "bench$N rc $?"; tail -2 gpurun_out/r02_bench_n${N}_d.err
python - <<PY
import json
d = json.load(open("gpurun_out/r02_bench_n${N}_d.json"))
print($N, {k: d[k] for k in ("value", "ms_per_step", "gpu_launches")}, "e2e", round(d["e2e"]["value"], 2), "pipe", round(d["e2e"]["c_abi_pipeline"]["value"], 2), "floor", d["e2e"]["torch_floor"]["ms_per_step"])
c3 = d["cfg3_multi_level"]; print("  cfg3", round(c3["us_per_step"], 1), round(c3["value"], 1), c3.get("check", {}).get("ok"))
print("  stats ok", d["stats_check"]["ok"], d["stats_check"]["exchange"], "spread", d["rank_spread"])
print("  cfg5", d["cfg5_crosscity"]["fused_ms"], d["cfg5_crosscity"]["images_per_s"], d["cfg5_crosscity"].get("check", {}).get("ok"))
ch = d["confusion_hist"]
print("  conf", {k: (round(v["value"], 1), round(v["frac_of_hbm_aggregate"], 3)) for k, v in ch.items() if isinstance(v, dict) and "value" in v}, ch["miou_16_13"], ch["matrix_total"])
PY
