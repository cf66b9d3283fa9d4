"""Agreement between CPU implementations of the SAME algorithm on the configurations where the GPU path sits below the
95 % bar (closed loop, W = 160): the "round-off floor" a third implementation can be expected to reach.

Test infrastructure (imports oracle/).  Three CPU arms, all following the reference line by line:
  ref      oracle/_ref/libref.so       the reference's own translation units (no FMA contraction)
  port     oracle/liboracle.so         the restatement (no FMA contraction)
  port_fma oracle/_floor/liboracle_fma.so  the restatement compiled with -mfma -ffp-contract=fast (what a GPU does)
For every arm pair the script prints the fraction of solves within 1e-8 relative final merit, fed IDENTICAL inputs:
  * config 4: cold W = 160 solves of the sweep batch;
  * config 3: teacher-forced closed loop — the compiled reference runs optimization_test.cc:38-61 and records, per
    tick, the plant state and the warm start every Step received; the other arms re-solve each tick from those inputs.

Usage: python tests/parity_floor.py [--long 512] [--plants 64] [--ticks 100] [--json out.json]
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from cart_pole_mpc_b200 import synthetic  # noqa: E402
from oracle import oracle_py  # noqa: E402


def build_fma_port() -> str:
    """The restatement with fused multiply-adds: same sources as liboracle.so, contraction on."""
    out_dir = os.path.join(ROOT, "oracle", "_floor")
    os.makedirs(out_dir, exist_ok=True)
    path = os.path.join(out_dir, "liboracle_fma.so")
    src = ["sqp.cc", "cartpole.cc", "double_cartpole.cc", "oracle_capi.cc"]
    newest = max(os.path.getmtime(os.path.join(ROOT, "oracle", f)) for f in src)
    if not os.path.exists(path) or os.path.getmtime(path) < newest:
        subprocess.run(["g++", "-O3", "-std=c++17", "-fPIC", "-fopenmp", "-mfma", "-ffp-contract=fast", "-shared",
                        "-o", path] + src, cwd=os.path.join(ROOT, "oracle"), check=True)
    return path


class FmaPort:
    """batch_step of the FMA build through the same ctypes prototypes."""

    def __init__(self):
        self.lib = oracle_py._bind(C.CDLL(build_fma_port()), full=True)

    def batch_step(self, *a, **k):
        saved = oracle_py._lib
        oracle_py._lib = self.lib
        try:
            return oracle_py.batch_step(*a, **k)
        finally:
            oracle_py._lib = saved


def merit(stats):
    return stats[:, 3] + stats[:, 5] * stats[:, 4]


def agree(a, b, tol=1e-8):
    ma, mb = merit(a), merit(b)
    rel = np.abs(ma - mb) / np.maximum(np.abs(mb), 1e-300)
    both_nan = ~np.isfinite(ma) & ~np.isfinite(mb)
    ok = (rel <= tol) | both_nan
    return {"within_1e-8": float(ok.mean()), "iteration_match": float((a[:, 1] == b[:, 1]).mean()),
            "termination_match": float((a[:, 0] == b[:, 0]).mean()), "n": int(ok.size)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--long", type=int, default=512, help="config-4 instances (W = 160)")
    ap.add_argument("--plants", type=int, default=64)
    ap.add_argument("--ticks", type=int, default=100)
    ap.add_argument("--json", default=None)
    args = ap.parse_args()
    fma = FmaPort()
    report = {}

    if args.long > 0:
        st, dy, sp, w = synthetic.sweep_batch(args.long)
        opt = oracle_py.opt_array(window_length=160)
        with oracle_py.using_ref():
            ref = oracle_py.batch_step(opt, st, dy, sp, weights=w, want_iter_log=False)
        port = oracle_py.batch_step(opt, st, dy, sp, weights=w, want_iter_log=False)
        pf = fma.batch_step(opt, st, dy, sp, weights=w, want_iter_log=False)
        report["config4_w160"] = {"port_vs_ref": agree(port.stats, ref.stats), "port_fma_vs_ref": agree(pf.stats, ref.stats),
                                  "port_fma_vs_port": agree(pf.stats, port.stats)}
        print("config 4 (W=160, cold, sweep batch):", json.dumps(report["config4_w160"]))

    for name, over in (("config3_defaults", dict()), ("config3_testvariant", dict(state_spacing=5, max_iterations=10))):
        if args.plants <= 0:
            break
        B, T = args.plants, args.ticks
        st, dy, sp = synthetic.swingup_batch(B, seed=synthetic.SEED_CONFIG2 + 1)
        opt = oracle_py.opt_array(**over)
        tr = oracle_py.closed_loop_trace(opt, st, dy, sp, T)
        N = tr.prev.shape[2]
        states = tr.states.reshape(B * T, 4)
        prev = tr.prev.reshape(B * T, N)
        has = np.tile((np.arange(T) > 0).astype(np.uint8), B)
        dyn_t = np.repeat(dy, T, axis=0)
        sp_t = np.repeat(sp, T)
        ref_stats = tr.stats.reshape(B * T, -1)
        port = oracle_py.batch_step(opt, states, dyn_t, sp_t, prev=prev, has_prev=has, want_iter_log=False)
        pf = fma.batch_step(opt, states, dyn_t, sp_t, prev=prev, has_prev=has, want_iter_log=False)
        report[name] = {"port_vs_ref": agree(port.stats, ref_stats), "port_fma_vs_ref": agree(pf.stats, ref_stats),
                        "ticks": T, "plants": B}
        print(f"{name} teacher-forced ({B} plants x {T} ticks):", json.dumps(report[name]))
    if args.json:
        with open(args.json, "w") as f:
            json.dump(report, f, indent=1)


if __name__ == "__main__":
    main()
