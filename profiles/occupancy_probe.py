#!/usr/bin/env python
"""Does a second warp per scheduler pay for THIS kernel?  The default problem (W=40, S=10) needs 5.4 KB of shared memory
per instance, which caps a block at 5 warps; the same code on a half-size problem (W=20, S=5: same 4 segments, same
lane mapping, same phases) needs 3.4 KB and fits 8 warps, so sweeping CPMPC_WARPS_PER_BLOCK there measures what extra
resident warps are worth before any shared-memory diet is attempted on the real problem.

    python profiles/occupancy_probe.py            # prints one JSON line per warps-per-block setting
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import json, os, sys
sys.path.insert(0, %r)
import numpy as np, torch
from cart_pole_mpc_b200 import BatchedOptimization, synthetic
B, W, S = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
st, dy, sp = synthetic.swingup_batch(B)
dev = torch.device("cuda:0")
to = lambda a: torch.from_numpy(a).to(dev)
std, dyd, spd = to(st), to(dy), to(sp)
s = BatchedOptimization(B, window_length=W, state_spacing=S)
u = torch.empty((B, W), dtype=torch.float64, device=dev)
pred = torch.empty((B, W, 4), dtype=torch.float64, device=dev)
ts = []
for i in range(8):
    s.reset_device()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); s.step_device(std, dyd, spd, u, pred_out=pred); e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
print(json.dumps({"B": B, "W": W, "S": S, "warps_per_block": os.environ.get("CPMPC_WARPS_PER_BLOCK"),
                  "ms": round(float(np.median(ts[2:])), 4), "solves_per_s": round(B / float(np.median(ts[2:])) * 1e3)}))
''' % ROOT

if __name__ == "__main__":
    for (W, S) in ((20, 5), (40, 10)):
        for wpb in (2, 3, 4, 5, 6, 7, 8):
            env = dict(os.environ, CPMPC_WARPS_PER_BLOCK=str(wpb))
            out = subprocess.run([sys.executable, "-c", CHILD, "65536", str(W), str(S)], env=env, capture_output=True, text=True)
            print(out.stdout.strip() or out.stderr.strip()[-300:], flush=True)
