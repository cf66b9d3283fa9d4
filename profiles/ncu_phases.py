#!/usr/bin/env python
"""Attributes the stall samples / executed instructions of an ncu report of cpmpc_solve_kernel to solver phases.

    python profiles/ncu_phases.py gpurun_out/prof.ncu-rep cart_pole_mpc_b200/libcpmpc.so cpmpc_solve_kernelILi4E

Instructions are walked in address order; inlined helper code (dynamics.cuh, fastmath.cuh) is charged to the phase of
the closest preceding solver_kernel.cuh line.  Phases are found from marker comments `// PHASE: name` is not needed:
the function / section boundaries are located by searching the source for their first lines.
"""
import collections
import os
import re
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import ncu_lines as nl  # noqa: E402

MARKS = [  # (regex that matches the first line of a region, phase name) in source order
    (r"void control_cost_partial", "ctrl_cost"),
    (r"void boundary_terms", "boundary"),
    (r"double rollout_defects", "merit_rollout"),
    (r"void evaluate_errors", "merit_eval"),
    (r"int cbase\(", "qp_helpers"),
    (r"bool riccati_qp", "riccati"),
    (r"bool condensed_qp", "condensed"),
    (r"^struct SolveResult", "solve_setup"),
    (r"---- linearise", "linearise"),
    (r"gradient of the control costs", "gradient"),
    (r"---- QP step", "qp_dir"),
    (r"---- line search", "linesearch"),
    (r"---- lambda update", "update"),
    (r"^__global__ void __launch_bounds__\(256, 1\) cpmpc_solve_kernel", "kernel_setup"),
    (r"---- BuildProblem", "init_fill"),
    (r"SolveResult& res = kc.res", "solve_call"),
    (r"---- outputs", "outputs_pred"),
    (r"^__global__ void cpmpc_sim_kernel", "sim_kernel"),
]


def phase_table(src_path):
    lines = open(src_path).read().split("\n")
    table = []
    for rx, name in MARKS:
        for i, ln in enumerate(lines):
            if re.search(rx, ln):
                table.append((i + 1, name))
                break
    table.sort()
    return table


def main():
    rep, lib, kern = sys.argv[1:4]
    src = os.path.join(os.path.dirname(os.path.abspath(lib)), "csrc", "solver_kernel.cuh")
    table = phase_table(src)

    def ph(line):
        cur = "prologue"
        for lo, name in table:
            if line >= lo:
                cur = name
        return cur

    kname, ix, rows = nl.sass_rows(rep)
    lmap = nl.line_map(lib, kern)
    n = min(len(lmap), len(rows))
    if len(lmap) != len(rows):
        print(f"warning: {len(rows)} profiled vs {len(lmap)} disassembled instructions (library rebuilt?)")
    S, E, F = collections.Counter(), collections.Counter(), collections.Counter()
    reasons = [h for h in ix if h.startswith("stall_") and "Not Issued" not in h]
    R = collections.defaultdict(collections.Counter)  # phase -> stall reason -> samples
    ts = te = 0
    cur = "prologue"
    for i in range(n):
        f, l = lmap[i][2]
        if f == "solver_kernel.cuh":
            cur = ph(l)
        s = int(rows[i][ix["# Samples"]] or 0)
        e = int(rows[i][ix["Instructions Executed"]] or 0)
        S[cur] += s
        E[cur] += e
        for h in reasons:
            v = rows[i][ix[h]]
            if v:
                R[cur][h] += int(v)
        ts += s
        te += e
        txt = lmap[i][1]
        op = txt.split()[1] if txt.startswith("@") else txt.split()[0]
        if op.startswith(("DFMA", "DMUL", "DADD")):
            F[cur] += e
    print(f"{kname}: {ts} samples, {te} warp instructions")
    print(f"{'phase':14s} {'time%':>6s} {'insts%':>7s} {'fp64/insts':>10s} {'rel cycles/inst':>15s}")
    for k, v in S.most_common():
        if v == 0 and E[k] == 0:
            continue
        print(f"{k:14s} {100 * v / ts:6.1f} {100 * E[k] / te:7.1f} {100 * F[k] / max(1, E[k]):10.1f} "
              f"{(v / ts) / (E[k] / te + 1e-12):15.2f}")
    if "--stalls" in sys.argv:
        print("\nstall samples per phase (% of the phase's samples), top reasons")
        for k, v in S.most_common():
            if v < 0.01 * ts:
                continue
            top = ", ".join(f"{h[6:]} {100 * c / max(1, sum(R[k].values())):.0f}" for h, c in R[k].most_common(5))
            print(f"{k:14s} {top}")


if __name__ == "__main__":
    main()
