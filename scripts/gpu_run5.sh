#!/bin/bash
# round-2 GPU call 5: full GPU tests, smoke, the rewritten bench at N=1 (default settings), reference arm
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q --timeout 120 > gpurun_out/r02_pytest_gpu_d.log 2>&1
tail -6 gpurun_out/r02_pytest_gpu_d.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.log 2>&1; tail -2 gpurun_out/r02_smoke.log
timeout 900 python bench.py > gpurun_out/r02_bench_a.json 2> gpurun_out/r02_bench_a.err
echo "bench rc $?"; tail -5 gpurun_out/r02_bench_a.err
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/r02_bench_a.json"))
    print({k: d[k] for k in ("value", "ms_per_step", "warmup", "gpu_launches")})
    print("e2e", {k: (v if not isinstance(v, dict) else {kk: vv for kk, vv in v.items() if kk != "how"}) for k, v in d["e2e"].items() if k != "how"})
    print("roofline", {k: d["roofline"][k] for k in ("achieved", "frac", "traffic")}, "issue", d["issue_roofline"]["frac"])
    print("cfg3", {k: v for k, v in d["cfg3_multi_level"].items() if k not in ("what", "exchange")})
    print("cfg5", {k: v for k, v in d["cfg5_crosscity"].items() if k != "what"})
    ch = d["confusion_hist"]
    print("conf", {k: (round(v["value"], 1), round(v["frac_of_hbm_aggregate"], 3)) for k, v in ch.items() if isinstance(v, dict)})
    print("cpu", d.get("cpu_baseline"))
    for k in d["kernels"]:
        print(f'{k["kernel"]:70s} {k["ms"]*1e3:8.1f} us {k["frac_of_hbm"]:.3f}')
except Exception as e:
    print("parse failed", e)
PY
