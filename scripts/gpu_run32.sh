#!/bin/bash
# round-2 evidence run (1 GPU): full GPU test suite, smoke, bench record, ncu launch list + ncu --set full of the step, timeline
mkdir -p gpurun_out
V=$PWD/maxsquareloss_b200/lib/variants
timeout 900 python -m pytest tests -m gpu -x -q --timeout 300 > gpurun_out/r02_pytest_gpu_final.log 2>&1; tail -3 gpurun_out/r02_pytest_gpu_final.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/r02_bench_n1_e.json 2> gpurun_out/r02_bench_n1_e.err; echo "bench1 rc $?"
MSQ_BENCH_MIN_WARM_S=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_b.csv python bench.py --steps 20 --warmup 3 --skip-secondary --skip-cpu > gpurun_out/r02_ncu_launches_b.log 2>&1; echo "ncu launches rc $?"
MSQ_BENCH_MIN_WARM_S=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"fused_|finalize" -c 12 -f -o /tmp/r02_fused_b python bench.py --steps 3 --warmup 3 --skip-secondary --skip-cpu > gpurun_out/r02_ncu_fused_b.log 2>&1; echo "ncu full rc $?"
ncu -i /tmp/r02_fused_b.ncu-rep --page raw --csv > gpurun_out/r02_fused_raw_b.csv 2>/dev/null
python scripts/ncu_summary.py /tmp/r02_fused_b.ncu-rep "ncu --set full --clock-control none --import-source on, bench.py --steps 3 --warmup 3 (MSQ_BENCH_MIN_WARM_S=0): the two kernels of the one-call fused step (the backward carries the finalisation in an extra CTA)" > gpurun_out/r02_ncu_fused_summary_b.txt 2>&1
AB_STEPS=3000 MSQ_B200_LIB=$V/libmsq_trace.so timeout 120 python scripts/trace_step.py > gpurun_out/r02_trace_two_final.txt 2>&1
AB_STEPS=3000 AB_LATE=0 MSQ_B200_LIB=$V/libmsq_trace.so timeout 120 python scripts/trace_step.py > gpurun_out/r02_trace_two_final_3kernel.txt 2>&1
for n in 2 1 4; do echo "== batch $n"; AB_QUICK=1 AB_N=$n timeout 120 python scripts/ab_queue.py 2>&1 | tail -3; done > gpurun_out/r02_ab_queue_final.txt 2>&1
cat gpurun_out/r02_ab_queue_final.txt
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02_bench_n1_e.json"))
print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches", "warmup")}, "e2e", round(d["e2e"]["value"], 2), d["e2e"]["ms_per_step"], "pipe", round(d["e2e"]["c_abi_pipeline"]["value"], 2), "floor", d["e2e"]["torch_floor"]["ms_per_step"])
print(d["roofline"]["frac"], d["issue_roofline"]["frac"] if d["issue_roofline"] else None, d.get("marginal_image"))
print([(k["kernel"][:30], round(k["ms"]*1e3, 1), round(k.get("frac_of_hbm", 0), 3)) for k in d["kernels"]])
print({k: (round(v["value"], 1), round(v["frac_of_hbm_aggregate"], 3)) for k, v in d["confusion_hist"].items() if isinstance(v, dict) and "value" in v})
print(d["cpu_baseline"]["value"], d["cfg3_multi_level"]["us_per_step"], d["maxsquare"], d["next_rows"])
PY
