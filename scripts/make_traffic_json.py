"""profiles/rNN_traffic.json from an ncu raw CSV (``ncu -i X.ncu-rep --page raw --csv``) of the bench's step kernels:
per kernel (median over the captured launches) DRAM bytes, duration, warp instructions, issue-slot and pipe utilisation --
the figures bench.py quotes next to its live timings (``roofline.traffic``, ``issue_roofline``).
    python scripts/make_traffic_json.py gpurun_out/r02_fused_raw.csv profiles/r02_traffic.json "source text" """
import csv
import json
import statistics
import sys

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3}


def main(path, out, source):
    rows = list(csv.reader(open(path)))
    head, units, body = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(head)}

    def val(r, name):
        return float(r[col[name]].replace(",", "")) * UNIT.get(units[col[name]], 1.0)
    groups = {"fused_fwd": [], "finalize": [], "fused_bwd": []}
    for r in body:
        name = r[col["Kernel Name"]]
        for key, pat in (("fused_fwd", "fused_fwd_kernel"), ("finalize", "finalize_kernel"), ("fused_bwd", "fused_bwd_kernel")):
            if pat in name:
                groups[key].append(r)
    res = {"source": source, "kernels": {}}
    for key, rs in groups.items():
        if not rs:
            continue
        med = lambda n: statistics.median(val(r, n) for r in rs)          # noqa: E731
        k = {"launches_captured": len(rs), "dram_read_bytes": med("dram__bytes_read.sum"), "dram_write_bytes": med("dram__bytes_write.sum"),
             "duration_us": med("gpu__time_duration.sum")}
        if key != "finalize":
            k.update({"warp_instructions": med("smsp__inst_executed.sum"),
                      "issue_active_pct": med("smsp__issue_active.avg.pct_of_peak_sustained_active"),
                      "xu_pipe_pct": med("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"),
                      "fma_pipe_pct": med("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"),
                      "alu_pipe_pct": med("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
                      "registers": med("launch__registers_per_thread"),
                      "warps_active_pct": med("sm__warps_active.avg.pct_of_peak_sustained_active"),
                      "sm_cycles_active_avg": med("sm__cycles_active.avg"), "sm_cycles_elapsed_max": med("sm__cycles_elapsed.max")})
        res["kernels"][key] = k
    json.dump(res, open(out, "w"), indent=1)
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else sys.argv[1])
