#!/usr/bin/env python
"""Step time of every 4096-instance shard of the 8-GPU weak-scaling run, measured one after the other on ONE GPU.
Every rank of `bench.py --gpus 8` owns instances [rank*4096, (rank+1)*4096) of the same seeded stream and one launch is
one wave of warps, so the job's step time is the slowest shard's slowest warp; this reproduces that per-shard time
without an 8-GPU box and names the straggler (slowest / median warp cycles, hardest instance).

    python profiles/shard_times.py [n_shards]
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from cart_pole_mpc_b200 import BatchedOptimization, _capi, synthetic  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 8
    B = 4096
    dev = torch.device("cuda:0")
    solver = BatchedOptimization(B, device=0)
    W = solver.W
    u = torch.empty((B, W), dtype=torch.float64, device=dev)
    pred = torch.empty((B, W, 4), dtype=torch.float64, device=dev)
    stats = torch.empty(B * np.dtype(_capi.STATS_DTYPE).itemsize, dtype=torch.uint8, device=dev)
    rows = []
    for shard in range(n):
        st, dy, sp = synthetic.swingup_batch(B, first=shard * B)
        std, dyd, spd = (torch.from_numpy(a).to(dev) for a in (st, dy, sp))
        ts = []
        for i in range(13):
            solver.reset_device()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            solver.step_device(std, dyd, spd, u, pred_out=pred, stats_out=stats)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        s = np.frombuffer(stats.cpu().numpy().tobytes(), dtype=np.dtype(_capi.STATS_DTYPE))
        if "--why" in sys.argv:
            # one more (untimed) launch through the general kernel, which times its phases (1.5 % slower)
            solver.enable_phase_clocks(True)
            solver.reset_device()
            solver.step_device(std, dyd, spd, u, pred_out=pred, stats_out=stats)
            torch.cuda.synchronize()
            solver.enable_phase_clocks(False)
        wci = solver.warp_cycles(with_instances=True)
        wci = wci[wci[:, 0] > 0]
        wc = wci[:, 0].astype(np.float64)
        worst = int(wci[np.argmax(wc), 1])
        if "--why" in sys.argv:
            med = np.median(wci[:, 2:].astype(np.float64), axis=0)
            top = wci[np.argmax(wc), 2:].astype(np.float64)
            print(json.dumps({"shard": shard, "phase_kcycles": {nm: [round(m / 1e3, 1), round(t / 1e3, 1)] for nm, m, t in
                                                                 zip(BatchedOptimization.PHASE_NAMES, med, top)},
                              "columns": "[median warp, slowest warp]"}))
        if "--why" in sys.argv and wc.max() > 1.2 * np.median(wc):
            solver.enable_iteration_log(True)
            solver.reset_device()
            solver.step_device(std, dyd, spd, u, pred_out=pred, stats_out=stats)
            torch.cuda.synchronize()
            log = solver.get_iteration_log()
            for i in range(worst, min(worst + 8, B)):
                it = int(s["n_iterations"][i])
                print(json.dumps({"instance": i, "iterations": it, "merit_evals": [int(v) for v in log[i, :it, 8]],
                                  "step_results": [int(v) for v in log[i, :it, 7]],
                                  "lambda": [float(v) for v in log[i, :it, 1]],
                                  "state": [round(float(v), 3) for v in st[i]], "m_b,m_1": [round(float(v), 3) for v in dy[i, :2]]}))
            solver.enable_iteration_log(False)
        rows.append({"shard": shard, "ms": round(float(np.median(ts[3:])), 4),
                     "slowest_warp_kcycles": round(float(wc.max()) / 1e3, 1),
                     "median_warp_kcycles": round(float(np.median(wc)) / 1e3, 1),
                     "max_failed_line_searches": int(s["n_failed_line_searches"].max()),
                     "instances_with_failed_line_search": int((s["n_failed_line_searches"] > 0).sum()),
                     "max_merit_evaluations": int(s["n_line_search_steps"].max())})
        print(json.dumps(rows[-1]), flush=True)
    ms = [r["ms"] for r in rows]
    print(json.dumps({"max_over_shards_ms": max(ms), "min_over_shards_ms": min(ms),
                      "weak_scaling_efficiency_if_run_on_8_gpus": round(ms[0] / max(ms), 3)}))


if __name__ == "__main__":
    main()
