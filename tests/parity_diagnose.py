"""Where does the CUDA path part from the compiled reference?  (VERDICT r01 "next round" item 1.)

Test infrastructure (imports oracle/; needs a GPU and oracle/_ref/libref.so).  Three studies, all through the C ABI:

  teacher   closed loop, teacher-forced: the compiled reference runs optimization_test.cc:38-61 for P plants x T ticks and
            records what every Step was GIVEN (plant state, warm start); the GPU re-solves every tick from exactly those
            inputs (cpmpc_set_previous_solution + cpmpc_step, all P*T solves as one batch) and each solve is compared
            with the reference's.  This separates solver disagreement from the chaotic divergence of free-running
            trajectories.
  long      config 4 (W = 160) cold solves of the sweep batch, per-iteration logs of both sides: for every instance
            outside 1e-8 the first solver quantity that differs by more than 1e-9 relative is located.
  blocks    cpmpc_debug_first_qp against oracle_linearize_qp: Phi / Gamma / e of the first linearisation and the QP step
            dx element-wise (condensed form and Riccati recursion, lambda = 0 and lambda = 1e-2), W = 40 and W = 160.

Usage: python tests/parity_diagnose.py [--plants 256] [--ticks 100] [--long 512] [--json out.json]
"""
from __future__ import annotations

import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from cart_pole_mpc_b200 import BatchedOptimization, synthetic  # noqa: E402
from oracle import oracle_py  # noqa: E402

LOG_FIELDS = ["optimizer_state", "lambda", "f(0)", "c(0)", "d_f", "d_eq", "penalty", "step_result", "n_steps"]


def _merit_gpu(stats):
    return stats["f"] + stats["penalty"] * stats["equality"]


def _merit_ref(stats):
    return stats[:, 3] + stats[:, 5] * stats[:, 4]


def compare(gpu_stats, ref_stats, tol=1e-8):
    g, r = _merit_gpu(gpu_stats), _merit_ref(ref_stats)
    rel = np.abs(g - r) / np.maximum(np.abs(r), 1e-300)
    ok = rel <= tol
    return ok, rel, {
        "n": int(ok.size), "within_1e-8": float(ok.mean()), "within_1e-6": float((rel <= 1e-6).mean()),
        "median_rel": float(np.median(rel)),
        "iteration_match": float((gpu_stats["n_iterations"] == ref_stats[:, 1].astype(int)).mean()),
        "termination_match": float((gpu_stats["termination_state"] == ref_stats[:, 0].astype(int)).mean()),
        "line_search_match": float((gpu_stats["n_line_search_steps"] == ref_stats[:, 2].astype(int)).mean()),
    }


def first_divergence(glog, rlog, n_iter, tol=1e-9):
    """(iteration, field name, relative difference) of the first logged quantity that differs by more than tol."""
    for it in range(n_iter):
        g, r = glog[it], rlog[it]
        for k in (2, 3, 4, 5, 6):
            d = abs(g[k] - r[k]) / max(abs(r[k]), 1e-300)
            if d > tol:
                return it, LOG_FIELDS[k], d
        if g[7] != r[7] or g[8] != r[8]:
            return it, LOG_FIELDS[7] if g[7] != r[7] else LOG_FIELDS[8], float("inf")
        for s in range(int(r[8])):
            for off, nm in ((9, "alpha"), (10, "step f"), (11, "step c")):
                a, b = g[off + 3 * s], r[off + 3 * s]
                d = abs(a - b) / max(abs(b), 1e-300)
                if d > tol:
                    return it, f"{nm}[{s}]", d
    return -1, "none", 0.0


def teacher_forced(P, T, over, name):
    st, dy, sp = synthetic.swingup_batch(P, seed=synthetic.SEED_CONFIG2 + 1)
    opt = oracle_py.opt_array(**over)
    tr = oracle_py.closed_loop_trace(opt, st, dy, sp, T)
    N = tr.prev.shape[2]
    B = P * T
    states = tr.states.reshape(B, 4)
    prev = tr.prev.reshape(B, N)
    mask = np.tile((np.arange(T) > 0).astype(np.uint8), P)
    s = BatchedOptimization(B, **over)
    s.set_previous_solution(prev, mask=mask)
    out = s.step(states, np.repeat(dy, T, axis=0), np.repeat(sp, T))
    ref_stats = tr.stats.reshape(B, -1)
    ok, rel, rep = compare(out.stats, ref_stats)
    du = np.abs(out.u - tr.u.reshape(B, -1)).max(axis=1) / np.maximum(np.abs(tr.u.reshape(B, -1)).max(axis=1), 1.0)
    rep["median_du_rel_where_ok"] = float(np.median(du[ok]))
    rep["u0_within_1e-6"] = float((np.abs(out.u[:, 0] - tr.u.reshape(B, -1)[:, 0]) <=
                                   1e-6 * np.maximum(np.abs(tr.u.reshape(B, -1)[:, 0]), 1.0)).mean())
    okm = ok.reshape(P, T)
    rep["cold_tick_within_1e-8"] = float(okm[:, 0].mean())
    rep["warm_ticks_within_1e-8"] = float(okm[:, 1:].mean())
    rep["plants"], rep["ticks"] = P, T
    rep["mean_ref_iterations"] = float(ref_stats[:, 1].mean())
    s.close()
    print(f"[teacher] {name}: {json.dumps(rep)}", flush=True)
    return rep


def long_horizon(B):
    over = dict(window_length=160)
    st, dy, sp, w = synthetic.sweep_batch(B)
    opt = oracle_py.opt_array(**over)
    with oracle_py.using_ref():
        ref = oracle_py.batch_step(opt, st, dy, sp, weights=w, want_iter_log=True)
    s = BatchedOptimization(B, **over)
    s.enable_iteration_log(True)
    out = s.step(st, dy, sp, weights=w)
    glog = s.get_iteration_log()
    ok, rel, rep = compare(out.stats, ref.stats)
    table = []
    hist = {}
    for i in np.where(~ok)[0]:
        n = int(min(ref.stats[i, 1], out.stats["n_iterations"][i]))
        it, field, d = first_divergence(glog[i], ref.iter_log[i], n)
        # agreement of iteration 0 (identical variables on both sides): the round-off level before amplification
        d0 = max(abs(glog[i, 0, k] - ref.iter_log[i, 0, k]) / max(abs(ref.iter_log[i, 0, k]), 1e-300) for k in (2, 3, 4, 5))
        table.append({"instance": int(i), "final_rel": float(rel[i]), "first_iteration": it, "first_field": field,
                      "first_rel_diff": (d if np.isfinite(d) else "different branch"), "iteration0_max_rel": float(d0),
                      "ref_iterations": int(ref.stats[i, 1]), "ref_line_search_steps": int(ref.stats[i, 2])})
        hist[f"iter{it}:{field.split('[')[0]}"] = hist.get(f"iter{it}:{field.split('[')[0]}", 0) + 1
    rep["first_divergence_histogram"] = dict(sorted(hist.items()))
    rep["failing"] = table
    # the same classification over the passing instances: how large is the iteration-0 difference there?
    d0_ok = [max(abs(glog[i, 0, k] - ref.iter_log[i, 0, k]) / max(abs(ref.iter_log[i, 0, k]), 1e-300) for k in (2, 3, 4, 5))
             for i in np.where(ok)[0][:256]]
    rep["iteration0_max_rel_median_passing"] = float(np.median(d0_ok)) if d0_ok else None
    rep["iteration0_max_rel_median_failing"] = float(np.median([t["iteration0_max_rel"] for t in table])) if table else None
    s.close()
    short = {k: v for k, v in rep.items() if k != "failing"}
    print(f"[long] W=160 x {B}: {json.dumps(short)}", flush=True)
    return rep


def _rel(a, b):
    scale = max(np.abs(b).max(), 1e-300)
    return float(np.abs(a - b).max() / scale)


def blocks(B, over, name, weights=False, force_condensed_layout=False):
    if weights:
        st, dy, sp, w = synthetic.sweep_batch(B)
    else:
        st, dy, sp = synthetic.swingup_batch(B, seed=31)
        w = None
    opt = oracle_py.opt_array(**over)
    W, S = int(opt[1]), int(opt[2])
    nseg = W // S
    NS = nseg + 1
    if force_condensed_layout:
        # long horizons get no condensed scratch (the form is not used there); CPMPC_QP=condensed at creation is the
        # experiment switch that lays it out anyway, to SHOW why it is not used (open-loop sensitivity products)
        os.environ["CPMPC_QP"] = "condensed"
    try:
        s = BatchedOptimization(B, **over)
    finally:
        os.environ.pop("CPMPC_QP", None)
    rep = {}
    for lam in (0.0, 1e-2):
        for method, mname in ((1, "riccati"), (2, "condensed")):
            if method == 2 and lam > 0.0:
                continue  # the condensed form has no damped variant: Step() never uses it with lambda > 0
            if method == 2 and int(opt[1]) * opt[0] > 0.6 and not force_condensed_layout:
                continue
            g = s.debug_first_qp(st, dy, sp, weights=w, lam=lam, qp_method=method)
            worst = {"Phi": 0.0, "Gam": 0.0, "e": 0.0, "dx": 0.0, "dx_vs_kkt": 0.0}
            n_ok = 0
            for i in range(B):
                o_i = opt.copy()
                if w is not None:
                    o_i[8:14] = w[i]
                q = oracle_py.linearize_qp(o_i, st[i], dy[i], sp[i], lam=lam)
                if q.status != 0 or not g["qp_ok"][i]:
                    continue
                n_ok += 1
                for sgm in range(nseg):
                    rows = q.A[4 * sgm:4 * sgm + 4]
                    worst["Phi"] = max(worst["Phi"], _rel(g["Phi"][i, sgm], rows[:, 4 * sgm:4 * sgm + 4]))
                    worst["Gam"] = max(worst["Gam"], _rel(g["Gam"][i][:, sgm * S:(sgm + 1) * S],
                                                          rows[:, 4 * NS + sgm * S:4 * NS + (sgm + 1) * S]))
                    worst["e"] = max(worst["e"], float(np.abs(g["e"][i, sgm] - q.b[4 * sgm:4 * sgm + 4]).max()))
                dx_g = np.concatenate([g["dxs"][i].ravel(), g["du"][i]])
                worst["dx"] = max(worst["dx"], _rel(dx_g, q.dx))
                # independent of both solvers: residual of the KKT conditions of the reference's QP at the GPU's step
                Gs = np.tril(q.G) + np.tril(q.G, -1).T
                r_eq = np.abs(q.A @ dx_g + q.b).max()
                # projected gradient: Z^T (G dx + c) = 0  <=>  G dx + c in range(A^T)
                grad = Gs @ dx_g + q.c
                lam_ls = np.linalg.lstsq(q.A.T, grad, rcond=None)[0]
                r_dual = np.abs(grad - q.A.T @ lam_ls).max() / max(np.abs(grad).max(), 1e-300)
                worst["dx_vs_kkt"] = max(worst["dx_vs_kkt"], float(max(r_eq, r_dual)))
            worst["instances"] = n_ok
            rep[f"{mname}_lambda{lam:g}"] = worst
    s.close()
    print(f"[blocks] {name}: {json.dumps(rep)}", flush=True)
    return rep


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--plants", type=int, default=256)
    ap.add_argument("--ticks", type=int, default=100)
    ap.add_argument("--long", type=int, default=512)
    ap.add_argument("--blocks", type=int, default=48)
    ap.add_argument("--json", default=None)
    args = ap.parse_args()
    if not oracle_py.ref_available():
        raise SystemExit("oracle/_ref/libref.so is required")
    report = {}
    if args.blocks > 0:
        report["blocks_w40"] = blocks(args.blocks, dict(), "W=40 defaults")
        report["blocks_w40_s5"] = blocks(args.blocks, dict(state_spacing=5, max_iterations=10), "W=40 S=5")
        report["blocks_w160"] = blocks(max(8, args.blocks // 3), dict(window_length=160), "W=160 sweep", weights=True)
        report["blocks_w160_condensed_forced"] = blocks(max(8, args.blocks // 3), dict(window_length=160),
                                                        "W=160 sweep, condensed scratch forced", weights=True,
                                                        force_condensed_layout=True)
    if args.plants > 0:
        report["teacher_defaults"] = teacher_forced(args.plants, args.ticks, dict(), "defaults S=10")
        report["teacher_testvariant"] = teacher_forced(args.plants, args.ticks, dict(state_spacing=5, max_iterations=10),
                                                       "optimization_test.cc S=5")
    if args.long > 0:
        report["long_w160"] = long_horizon(args.long)
    if args.json:
        with open(args.json, "w") as f:
            json.dump(report, f, indent=1)


if __name__ == "__main__":
    main()
