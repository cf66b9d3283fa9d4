#!/usr/bin/env python
"""Attributes the warp-stall samples of an `ncu --set full --import-source on` report to source lines.

    python profiles/ncu_lines.py gpurun_out/prof.ncu-rep cart_pole_mpc_b200/libcpmpc.so 'cpmpc_solve_kernelILi4E' [--top 40]

ncu's SASS page gives samples / executed counts per instruction; nvdisasm -g on the cubin embedded in the library gives
the (innermost, after inlining) source line of every instruction in the same order.  The two are zipped and summed per
line and per source range, which shows where the time of the one big solve kernel goes (merit rollouts, linearisation,
QP, ...), see DESIGN.md section 4.
"""
from __future__ import annotations

import argparse
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile


def sass_rows(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    return rows[0][1], ix, rows[2:]


def line_map(lib, mangled_sub):
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, capture_output=True)
    txt = ""
    for cubin in sorted(f for f in os.listdir(tmp) if f.endswith(".cubin")):
        # the library holds one cubin per translation unit: take the one that defines the kernel
        cand = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
        if any(l.startswith("//---") and ".text." in l and mangled_sub in l for l in cand.splitlines()):
            txt = cand
            break
    lines = txt.splitlines()
    out = []
    inside = False
    cur = ("?", 0)
    for ln in lines:
        if ln.startswith("//---") and ".text." in ln:
            inside = mangled_sub in ln
            continue
        if not inside:
            continue
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s*/\*([0-9a-f]+)\*/\s+(.*?);", ln)
        if m:
            out.append((int(m.group(1), 16), m.group(2).strip(), cur))
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("rep")
    ap.add_argument("lib")
    ap.add_argument("kernel", help="substring of the MANGLED kernel name, e.g. cpmpc_solve_kernelILi4E")
    ap.add_argument("--top", type=int, default=40)
    ap.add_argument("--ranges", default=None,
                    help="comma list file:lo-hi=name to sum over source ranges, e.g. solver_kernel.cuh:150-200=merit")
    args = ap.parse_args()
    kname, ix, rows = sass_rows(args.rep)
    lmap = line_map(args.lib, args.kernel)
    if len(lmap) != len(rows):
        print(f"warning: {len(rows)} profiled instructions vs {len(lmap)} disassembled (library rebuilt since?)")
    n = min(len(lmap), len(rows))
    per_line = collections.Counter()
    per_line_inst = collections.Counter()
    tot_s = tot_e = 0
    for i in range(n):
        s = int(rows[i][ix["# Samples"]] or 0)
        e = int(rows[i][ix["Instructions Executed"]] or 0)
        per_line[lmap[i][2]] += s
        per_line_inst[lmap[i][2]] += e
        tot_s += s
        tot_e += e
    print(f"kernel {kname}: {tot_s} samples, {tot_e} warp instructions")
    print("-- top lines by samples")
    for (f, l), s in per_line.most_common(args.top):
        print(f"  {f}:{l:5d}  samples {100.0 * s / tot_s:5.1f}%  warp-insts {100.0 * per_line_inst[(f, l)] / tot_e:5.1f}%")
    per_file = collections.Counter()
    for (f, l), s in per_line.items():
        per_file[f] += s
    print("-- by file")
    for f, s in per_file.most_common():
        print(f"  {f:24s} {100.0 * s / tot_s:5.1f}%")
    if args.ranges:
        print("-- by range (innermost inlined line)")
        for spec in args.ranges.split(","):
            loc, name = spec.split("=")
            f, r = loc.split(":")
            lo, hi = (int(v) for v in r.split("-"))
            s = sum(v for (ff, l), v in per_line.items() if ff == f and lo <= l <= hi)
            e = sum(v for (ff, l), v in per_line_inst.items() if ff == f and lo <= l <= hi)
            print(f"  {name:28s} samples {100.0 * s / tot_s:5.1f}%  warp-insts {100.0 * e / tot_e:5.1f}%")


if __name__ == "__main__":
    sys.exit(main())
