#!/usr/bin/env python
"""Executed FP64 thread instructions per step from `ncu --metrics ... --csv` logs (profiles/r02_y_ncu_fp64.sh): sums the
kernels of ONE step (fused launch: the solve kernel; split launch: guess + solve + post) and writes the JSON that
bench.py reads for `roofline.executed_fp64`.

    python profiles/fp64_ops.py profiles/r02_y_fp64_ops.json config2=gpurun_out/a.csv:4096 config2_large=...:65536
"""
import csv
import io
import json
import sys


def parse(path):
    rows = [l for l in open(path) if not l.startswith("==")]
    rd = csv.DictReader(io.StringIO("".join(rows)))
    per_kernel = {}
    for r in rd:
        k = per_kernel.setdefault(r["ID"], {"name": r["Kernel Name"]})
        k[r["Metric Name"]] = float(r["Metric Value"].replace(",", ""))
    return [per_kernel[i] for i in sorted(per_kernel, key=int)]


def main():
    out_path = sys.argv[1]
    out = {}
    for spec in sys.argv[2:]:
        name, rest = spec.split("=")
        path, batch = rest.rsplit(":", 1)
        ks = parse(path)
        g = lambda key: sum(k.get(key, 0.0) for k in ks)
        dfma = g("smsp__sass_thread_inst_executed_op_dfma_pred_on.sum")
        dmul = g("smsp__sass_thread_inst_executed_op_dmul_pred_on.sum")
        dadd = g("smsp__sass_thread_inst_executed_op_dadd_pred_on.sum")
        solve = max(ks, key=lambda k: k.get("gpu__time_duration.sum", 0.0))
        out[name] = {
            "batch": int(batch),
            "kernels": [k["name"].split("(")[0] for k in ks],
            "ns_under_ncu": g("gpu__time_duration.sum"),
            "fp64_pipe_active_pct_solve_kernel": solve.get("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"),
            "warp_instructions": g("smsp__inst_executed.sum"),
            "dadd": dadd, "dfma": dfma, "dmul": dmul,
            "executed_fp64_flops_per_solve": (2.0 * dfma + dmul + dadd) / int(batch),
        }
    with open(out_path, "w") as fh:
        json.dump(out, fh, indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
