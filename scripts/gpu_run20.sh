#!/bin/bash
mkdir -p gpurun_out
for n in 2 4; do echo "batch $n"; AB_N=$n timeout 600 python scripts/ab_variants.py run 2>&1 | grep "^libmsq" | grep -v trace; done
timeout 900 python bench.py --skip-cpu > gpurun_out/r02_bench_n1_d.json 2> gpurun_out/r02_bench_n1_d.err; echo "bench rc $?"
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02_bench_n1_d.json"))
print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches")}, "e2e", round(d["e2e"]["value"], 2), d["e2e"]["ms_per_step"], "pipe", round(d["e2e"]["c_abi_pipeline"]["value"], 2))
print(d.get("marginal_image"))
print([ (k["kernel"][:28], round(k["ms"]*1e3,1)) for k in d["kernels"]])
PY
