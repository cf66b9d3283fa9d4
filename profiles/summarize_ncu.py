"""Summarises an `ncu --set full --import-source on` report of the solve kernel into a small JSON + text file that
can be committed under profiles/ (the .ncu-rep itself stays in gpurun_out/).

    python profiles/summarize_ncu.py gpurun_out/prof.ncu-rep profiles/r01_xxx
"""
from __future__ import annotations

import collections
import csv
import io
import json
import re
import subprocess
import sys

RAW_KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
    "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed_op_shared_ld.sum",
    "smsp__inst_executed_op_shared_st.sum", "smsp__sass_inst_executed_op_local_ld.sum",
    "smsp__sass_inst_executed_op_local_st.sum",
    "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum", "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum",
    "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum", "sm__cycles_elapsed.max",
]


def ncu_csv(rep, page):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep, prefix = sys.argv[1], sys.argv[2]
    raw = ncu_csv(rep, "raw")
    hdr, units, vals = raw[0], raw[1], raw[2]
    summary = {"report": rep, "kernel": vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else None, "raw": {}}
    for k in RAW_KEYS:
        if k in hdr:
            i = hdr.index(k)
            summary["raw"][k] = {"value": vals[i], "unit": units[i]}
    for i, h in enumerate(hdr):
        if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
            try:
                v = float(vals[i])
            except ValueError:
                continue
            if v >= 0.01:
                summary.setdefault("stalls_per_issue", {})[h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]] = v

    src = ncu_csv(rep, "source")
    if len(src) > 2:
        h2 = src[1]
        ix = {h: i for i, h in enumerate(h2)}
        data = src[2:]
        g = lambda r, k: int(r[ix[k]] or 0)
        execd = [g(r, "Instructions Executed") for r in data]
        samples = [g(r, "# Samples") for r in data]
        tot_e, tot_s = sum(execd), sum(samples)
        summary["sass_instructions"] = len(data)
        summary["sass_bytes"] = 16 * len(data)
        summary["warp_instructions_executed"] = tot_e
        order = sorted(range(len(data)), key=lambda i: -execd[i])
        hot = {}
        for frac in (0.5, 0.9, 0.99):
            c = n = 0
            for i in order:
                c += execd[i]
                n += 1
                if c >= frac * tot_e:
                    break
            hot[str(frac)] = {"static_instructions": n, "kib": round(n * 16 / 1024, 1)}
        summary["hot_code_footprint"] = hot
        stall = {}
        for h in h2:
            if h.startswith("stall_") and "(" not in h:
                v = sum(g(r, h) for r in data)
                if v >= 0.01 * tot_s:
                    stall[h] = round(100.0 * v / tot_s, 1)
        summary["stall_samples_pct"] = stall

        def opc(s):
            s = re.sub(r"^@!?U?P\d+\s+", "", s.strip())
            return s.split()[0].split(".")[0] if s else "?"

        mix = collections.Counter()
        for e, r in zip(execd, data):
            mix[opc(r[ix["Source"]])] += e
        summary["opcode_mix_pct"] = {k: round(100.0 * v / tot_e, 1) for k, v in mix.most_common(14)}
    with open(prefix + ".json", "w") as fh:
        json.dump(summary, fh, indent=1)
    print(json.dumps(summary, indent=1))


if __name__ == "__main__":
    main()
