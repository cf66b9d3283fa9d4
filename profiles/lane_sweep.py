#!/usr/bin/env python
"""Device-timed sweep of batch sizes / mapping overrides for the solve kernel (CUDA events on the launch stream,
inputs resident, cold start every step).  Mapping overrides come from the environment (CPMPC_LANES,
CPMPC_WARPS_PER_BLOCK, CPMPC_CHUNKS, CPMPC_QP), see csrc/cpmpc_capi.cu.

    CPMPC_LANES=8 python profiles/lane_sweep.py 8 32 4096 65536
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from cart_pole_mpc_b200 import BatchedOptimization, _capi, synthetic  # noqa: E402

if os.environ.get("SWEEP_LIB"):  # A/B of two builds of the library on the same box (a tool switch, not the product's)
    _capi.LIB_PATH = os.path.abspath(os.environ["SWEEP_LIB"])


def run(B, reps=20, first=0, **kw):
    states, dyn, sp = synthetic.swingup_batch(B, first=first)
    dev = torch.device("cuda:0")
    st = torch.from_numpy(states).to(dev)
    dy = torch.from_numpy(dyn).to(dev)
    spd = torch.from_numpy(sp).to(dev)
    solver = BatchedOptimization(B, device=0, **kw)
    W = solver.W
    u = torch.empty((B, W), dtype=torch.float64, device=dev)
    pred = torch.empty((B, W, 4), dtype=torch.float64, device=dev)
    stats = torch.empty(B * np.dtype(_capi.STATS_DTYPE).itemsize, dtype=torch.uint8, device=dev)
    ts = []
    for i in range(reps + 3):
        solver.reset_device()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        solver.step_device(st, dy, spd, u, pred_out=pred, stats_out=stats)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    sts = np.frombuffer(stats.cpu().numpy().tobytes(), dtype=np.dtype(_capi.STATS_DTYPE))
    import hashlib
    digest = hashlib.sha1(u.cpu().numpy().tobytes() + pred.cpu().numpy().tobytes()).hexdigest()[:12]
    solver.close()
    ts = ts[3:]
    return float(np.median(ts)), float(sts["n_line_search_steps"].mean()), float(sts["n_iterations"].mean()), digest


if __name__ == "__main__":
    env = {k: v for k, v in os.environ.items() if k.startswith("CPMPC")}
    for B in (int(x) for x in sys.argv[1:]):
        ms, nls, nit, digest = run(B)
        print(json.dumps({"B": B, "env": env, "ms_kernel_median": round(ms, 4), "solves_per_s": round(B / ms * 1e3),
                          "mean_ls": round(nls, 2), "mean_iter": round(nit, 2), "sha1_u_pred": digest}))
