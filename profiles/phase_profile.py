#!/usr/bin/env python
"""Where the solve kernel's cycles go, per solver phase, from the kernel's own per-warp clocks (cpmpc_get_warp_cycles;
no profiler attached, so the launch runs at full speed).  Prints the share of warp-cycles per phase for
    config2        4096 cold swing-ups (one wave of warps)
    config2-large  65536 cold swing-ups
    config4        W = 160 sweep batch (segment-wise Riccati every iteration)
and the launch time.

    python profiles/phase_profile.py [config2 config2-large config4 ...]
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from cart_pole_mpc_b200 import BatchedOptimization, synthetic  # noqa: E402


def run(name):
    over, weights = {}, None
    if name == "config2":
        B = 4096
        st, dy, sp = synthetic.swingup_batch(B)
    elif name == "config2-large":
        B = 65536
        st, dy, sp = synthetic.swingup_batch(B)
    elif name == "config4":
        B = 32768
        st, dy, sp, weights = synthetic.sweep_batch(B)
        over = dict(window_length=160)
    else:
        raise SystemExit(f"unknown workload {name}")
    dev = torch.device("cuda:0")
    to = lambda a: torch.from_numpy(a).to(dev)
    std, dyd, spd = to(st), to(dy), to(sp)
    wd = to(weights) if weights is not None else None
    s = BatchedOptimization(B, **over)
    s.enable_phase_clocks()  # the general kernel instantiation: 1.5 % slower than the lean one the library launches by default
    u = torch.empty((B, s.W), dtype=torch.float64, device=dev)
    pred = torch.empty((B, s.W, 4), dtype=torch.float64, device=dev)
    ts = []
    for i in range(5):
        s.reset_device()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        s.step_device(std, dyd, spd, u, weights=wd, pred_out=pred)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    rec = s.warp_cycles(with_instances=True).astype(np.float64)
    rec = rec[rec[:, 0] > 0]
    ph = rec[:, 2:].sum(axis=0)
    out = {"workload": name, "batch": B, "ms": round(float(np.median(ts[1:])), 4), "warps": int(rec.shape[0]),
           "mean_warp_kcycles": round(float(rec[:, 0].mean()) / 1e3, 1),
           "phase_share_percent": {n: round(100.0 * v / ph.sum(), 1) for n, v in zip(BatchedOptimization.PHASE_NAMES, ph)}}
    print(json.dumps(out), flush=True)
    s.close()


if __name__ == "__main__":
    for nm in (sys.argv[1:] or ["config2", "config2-large", "config4"]):
        run(nm)
