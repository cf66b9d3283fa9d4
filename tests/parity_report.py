"""Parity report: CUDA path (through the C ABI) vs the CPU oracle on a seeded synthetic batch.

Test infrastructure (may import oracle/).  Usage:
    python tests/parity_report.py [--batch 4096] [--config default|test|sweep] [--json out.json]
Prints, per SURVEY.md section 8(d): fraction of instances within 1e-8 relative final merit, max |du|,
iteration-count and termination-state match rates.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from cart_pole_mpc_b200 import BatchedOptimization, synthetic  # noqa: E402
from oracle import oracle_py  # noqa: E402


def compare(gpu_out, ora, rel_tol=1e-8):
    """Returns a dict of parity metrics between a StepOutputs and an oracle BatchResult."""
    g_f, g_eq, g_pen = gpu_out.stats["f"], gpu_out.stats["equality"], gpu_out.stats["penalty"]
    o_f, o_eq, o_pen = ora.stats[:, 3], ora.stats[:, 4], ora.stats[:, 5]
    g_merit = g_f + g_pen * g_eq
    o_merit = o_f + o_pen * o_eq
    denom = np.maximum(np.abs(o_merit), 1e-300)
    rel = np.abs(g_merit - o_merit) / denom
    # cost alone (penalty-independent): f and equality compared separately, relative to the oracle's merit scale
    rel_f = np.abs(g_f - o_f) / np.maximum(np.abs(o_f), 1e-300)
    du = np.abs(gpu_out.u - ora.u).max(axis=1)
    u_scale = np.maximum(np.abs(ora.u).max(axis=1), 1.0)
    it_match = gpu_out.stats["n_iterations"] == ora.n_iterations
    ls_match = gpu_out.stats["n_line_search_steps"] == ora.n_line_search
    term_match = gpu_out.stats["termination_state"] == ora.termination
    ok = rel <= rel_tol
    return {
        "instances": int(rel.size),
        "frac_within_rel_merit_1e-8": float(ok.mean()),
        "frac_within_rel_f_1e-8": float((rel_f <= rel_tol).mean()),
        "median_rel_merit": float(np.median(rel)),
        "p99_rel_merit": float(np.percentile(rel, 99)),
        "max_abs_du": float(du.max()),
        "median_abs_du": float(np.median(du)),
        "frac_du_rel_1e-6": float((du / u_scale <= 1e-6).mean()),
        "iteration_match": float(it_match.mean()),
        "line_search_match": float(ls_match.mean()),
        "termination_match": float(term_match.mean()),
        "all_three_match": float((ok & it_match & term_match).mean()),
    }


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=4096)
    ap.add_argument("--config", default="default", choices=["default", "test", "sweep", "long"])
    ap.add_argument("--json", default=None)
    ap.add_argument("--show-bad", type=int, default=3)
    args = ap.parse_args()

    B = args.batch
    over = {}
    weights = None
    if args.config == "test":
        over = dict(state_spacing=5, max_iterations=10)
    if args.config in ("sweep", "long"):
        states, dyn, sp, weights = synthetic.sweep_batch(B)
        if args.config == "long":
            over = dict(window_length=160)
    else:
        states, dyn, sp = synthetic.swingup_batch(B)
    opt = oracle_py.opt_array(**over)
    t0 = time.time()
    ora = oracle_py.batch_step(opt, states, dyn, sp, weights=weights)
    t_cpu = time.time() - t0

    solver = BatchedOptimization(B, **over)
    solver.enable_iteration_log(True)
    t0 = time.time()
    out = solver.step(states, dyn, sp, weights=weights)
    t_gpu = time.time() - t0
    rep = compare(out, ora)
    rep["config"] = args.config
    rep["cpu_seconds"] = t_cpu
    rep["cpu_threads"] = ora.threads
    rep["gpu_seconds_first_call"] = t_gpu
    rep["oracle_termination_hist"] = np.bincount(ora.termination, minlength=7).tolist()
    rep["gpu_termination_hist"] = np.bincount(np.maximum(out.stats["termination_state"], 0), minlength=7).tolist()
    rep["oracle_mean_iterations"] = float(ora.n_iterations.mean())
    rep["oracle_mean_line_search"] = float(ora.n_line_search.mean())
    rep["gpu_mean_iterations"] = float(out.stats["n_iterations"].mean())
    rep["gpu_mean_line_search"] = float(out.stats["n_line_search_steps"].mean())
    # predicted states parity
    rep["max_abs_dpred"] = float(np.nanmax(np.abs(out.predicted_states - ora.predicted)))
    print(json.dumps(rep, indent=1))

    if args.show_bad:
        g_merit = out.merit
        o_merit = ora.merit
        rel = np.abs(g_merit - o_merit) / np.maximum(np.abs(o_merit), 1e-300)
        worst = np.argsort(-rel)[: args.show_bad]
        glog = solver.get_iteration_log()
        for i in worst:
            print(f"--- instance {i}: rel merit diff {rel[i]:.3e}; gpu term {out.stats['termination_state'][i]} "
                  f"iters {out.stats['n_iterations'][i]} | oracle term {ora.termination[i]} "
                  f"iters {ora.n_iterations[i]}")
            for it in range(max(int(out.stats['n_iterations'][i]), int(ora.n_iterations[i]))):
                g, o = glog[i, it], ora.iter_log[i, it]
                print(f"  it{it} GPU f0 {g[2]:.12e} eq0 {g[3]:.12e} df {g[4]:.6e} deq {g[5]:.6e} pen {g[6]:.8e} "
                      f"res {int(g[7])} n {int(g[8])} alphas {[round(g[9 + 3 * s], 6) for s in range(int(g[8]))]}")
                print(f"      ORA f0 {o[2]:.12e} eq0 {o[3]:.12e} df {o[4]:.6e} deq {o[5]:.6e} pen {o[6]:.8e} "
                      f"res {int(o[7])} n {int(o[8])} alphas {[round(o[9 + 3 * s], 6) for s in range(int(o[8]))]}")
    if args.json:
        os.makedirs(os.path.dirname(os.path.abspath(args.json)), exist_ok=True)
        with open(args.json, "w") as fh:
            json.dump(rep, fh, indent=1)


if __name__ == "__main__":
    main()
