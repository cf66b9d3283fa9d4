"""Parity report for the plant-generic CUDA path (test infrastructure; may import oracle/).

    python tests/plant_report.py [--batch 1024]
Prints, for the single pendulum through the generic kernels and for the double pendulum, the fraction of instances
within 1e-8 relative final merit of the matching oracle, iteration-count / termination match and max |du|.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cart_pole_mpc_b200 import PlantBatchedOptimization, synthetic  # noqa: E402
from oracle import oracle_py as O  # noqa: E402


def compare(out, ora):
    g = out.merit
    rel = np.abs(g - ora.merit) / np.maximum(np.abs(ora.merit), 1e-300)
    du = np.abs(out.u - ora.u).max(axis=1)
    return {
        "frac_within_1e-8_rel_cost": float((rel <= 1e-8).mean()),
        "median_rel": float(np.median(rel)),
        "iteration_match": float((out.stats["n_iterations"] == ora.n_iterations).mean()),
        "line_search_match": float((out.stats["n_line_search_steps"] == ora.n_line_search).mean()),
        "termination_match": float((out.stats["termination_state"] == ora.termination).mean()),
        "median_abs_du": float(np.median(du)),
        "max_abs_dpred": float(np.nanmax(np.abs(out.predicted_states - ora.predicted))),
        "oracle_term_hist": np.bincount(ora.termination, minlength=7).tolist(),
        "gpu_term_hist": np.bincount(np.maximum(out.stats["termination_state"], 0), minlength=7).tolist(),
        "mean_iterations": float(ora.n_iterations.mean()),
        "mean_line_search": float(ora.n_line_search.mean()),
    }


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=1024)
    args = ap.parse_args()
    B = args.batch
    st, dy, sp = synthetic.swingup_batch(B)
    ora = O.batch_step(O.opt_array(), st, dy, sp, want_iter_log=False)
    t0 = time.time()
    out = PlantBatchedOptimization("single", B).step(st, dy, sp)
    print("single pendulum through the generic kernels:", json.dumps(compare(out, ora)), f"{time.time()-t0:.2f}s")
    for over in (dict(), dict(state_spacing=5, max_iterations=10)):
        st, dy, sp = synthetic.double_swingup_batch(B)
        ora = O.dp_batch_step(O.dp_opt_array(**over), st, dy, sp, want_iter_log=False)
        t0 = time.time()
        out = PlantBatchedOptimization("double", B, **over).step(st, dy, sp)
        print("double pendulum", over, json.dumps(compare(out, ora)), f"{time.time()-t0:.2f}s")
    # warm-started second step
    solver = PlantBatchedOptimization("double", B)
    o1 = solver.step(st, dy, sp)
    prev, has = solver.get_previous_solution()
    o2 = solver.step(st, dy, sp)
    ora2 = O.dp_batch_step(O.dp_opt_array(), st, dy, sp, prev=prev, has_prev=has.astype(np.uint8), want_iter_log=False)
    print("double pendulum warm second step:", json.dumps(compare(o2, ora2)))


if __name__ == "__main__":
    main()
