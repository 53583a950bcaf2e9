#!/bin/bash
# round-2 closing run (1 GPU): full GPU test suite, smoke, bench record + reference arm with the final library
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --timeout 300 > gpurun_out/r02_pytest_gpu_final.log 2>&1; tail -3 gpurun_out/r02_pytest_gpu_final.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/r02_bench_n1_f.json 2> gpurun_out/r02_bench_n1_f.err; echo "bench1 rc $?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_reference_f.json 2>/dev/null; echo "ref rc $?"
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02_bench_n1_f.json"))
print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches", "warmup")}, "e2e", round(d["e2e"]["value"], 2), d["e2e"]["ms_per_step"], "pipe", round(d["e2e"]["c_abi_pipeline"]["value"], 2), "floor", d["e2e"]["torch_floor"]["ms_per_step"])
print(d["roofline"]["frac"], d["issue_roofline"]["frac"], d["issue_roofline"]["warp_instructions_per_launch"], d.get("marginal_image"))
r = json.load(open("gpurun_out/r02_bench_reference_f.json")); print("reference", r["value"], r.get("cpu_baseline"))
PY
