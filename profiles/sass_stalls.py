#!/usr/bin/env python
"""Static issue-time estimate of SASS basic blocks from the scheduler control bits (no GPU needed).

    python profiles/sass_stalls.py cart_pole_mpc_b200/libcpmpc.so 'cpmpc_solve_kernel<4>' [--min-insts 60]

For every basic block of the named kernel it prints: instruction count, FP64 instruction count, the sum of the
compiler-scheduled stall counts (bits 105..108 of each 128-bit instruction: the minimum number of cycles ONE warp
needs to issue the block when no scoreboard wait fires), and the opcode mix.  Used to compare code variants of the
dynamics evaluation / RK4 stage loop before spending GPU time (a single resident warp per scheduler is
issue/latency bound).
"""
from __future__ import annotations

import argparse
import collections
import re
import subprocess
import sys


def demangled_functions(lib):
    out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    funcs = collections.OrderedDict()
    name = None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            name = m.group(1)
            funcs[name] = []
            continue
        if name is not None:
            funcs[name].append(line)
    return funcs


def demangle(names):
    p = subprocess.run(["c++filt"] + list(names), capture_output=True, text=True)
    return dict(zip(names, p.stdout.splitlines()))


INST = re.compile(r"^\s*/\*([0-9a-f]+)\*/\s+(.*?);\s*/\* 0x([0-9a-f]+) \*/")
HI = re.compile(r"^\s*/\* 0x([0-9a-f]+) \*/")


def parse(lines):
    insts = []
    for i, line in enumerate(lines):
        m = INST.match(line)
        if not m:
            continue
        addr = int(m.group(1), 16)
        text = m.group(2).strip()
        hi = 0
        if i + 1 < len(lines):
            mh = HI.match(lines[i + 1])
            if mh:
                hi = int(mh.group(1), 16)
        stall = (hi >> 41) & 0xF
        wait_mask = (hi >> 52) & 0x3F
        wbar = (hi >> 46) & 0x7
        insts.append((addr, text, stall, wait_mask, wbar))
    return insts


def opcode(text):
    t = text
    if t.startswith("@"):
        t = t.split(None, 1)[1]
    return t.split()[0].split(".")[0]


def blocks(insts):
    targets = set()
    for addr, text, *_ in insts:
        m = re.search(r"\b(BRA|BSSY|CALL)\b.*?0x([0-9a-f]+)", text)
        if m and m.group(1) == "BRA":
            targets.add(int(m.group(2), 16))
    out, cur = [], []
    for ins in insts:
        if ins[0] in targets and cur:
            out.append(cur)
            cur = []
        cur.append(ins)
        if opcode(ins[1]) in ("BRA", "EXIT", "RET", "BRX"):
            out.append(cur)
            cur = []
    if cur:
        out.append(cur)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("lib")
    ap.add_argument("kernel", help="substring of the demangled kernel name")
    ap.add_argument("--min-insts", type=int, default=60)
    ap.add_argument("--dump", type=int, default=None, help="print the instructions of block #N")
    args = ap.parse_args()
    funcs = demangled_functions(args.lib)
    names = demangle(list(funcs))
    for mangled, lines in funcs.items():
        if args.kernel not in names.get(mangled, mangled):
            continue
        insts = parse(lines)
        print(f"== {names[mangled]}: {len(insts)} instructions")
        for bi, b in enumerate(blocks(insts)):
            if args.dump is not None and bi == args.dump:
                for addr, text, stall, wm, wb in b:
                    print(f"  {addr:06x} st={stall:2d} wait={wm:02x} wbar={wb} {text}")
            if len(b) < args.min_insts:
                continue
            ops = collections.Counter(opcode(t) for _, t, *_ in b)
            fp64 = sum(v for k, v in ops.items() if k in ("DFMA", "DMUL", "DADD", "DSETP", "DMNMX"))
            mufu = ops.get("MUFU", 0)
            stall = sum(s for _, _, s, *_ in b)
            nwait = sum(1 for _, _, _, wm, _ in b if wm)
            top = ", ".join(f"{k}:{v}" for k, v in ops.most_common(8))
            print(f"  block {bi:3d} @{b[0][0]:06x}: {len(b):4d} insts, fp64 {fp64:4d}, mufu {mufu}, "
                  f"sum(stall) {stall:5d}, scoreboard waits {nwait:3d} | {top}")


if __name__ == "__main__":
    sys.exit(main())
