#!/usr/bin/env python
"""Small runs of every launch shape for compute-sanitizer (memcheck): fused + roomy, fused + compact, split + compact,
W = 160 split, closed loop split, plant-generic split.

    compute-sanitizer --tool memcheck --error-exitcode 3 python profiles/memcheck_target.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

from cart_pole_mpc_b200 import BatchedOptimization, synthetic  # noqa: E402
from cart_pole_mpc_b200.plant import PlantBatchedOptimization  # noqa: E402


def run(name, solver, st, dy, sp, **kw):
    out = solver.step(st, dy, sp, **kw)
    out2 = solver.step(st + 0.01, dy, sp, **kw)
    print(name, "ok", float(np.nanmean(out.stats["n_iterations"])), float(np.nanmean(out2.stats["n_iterations"])), flush=True)
    solver.close()


for B in (1000, 8192, 12288):
    st, dy, sp = synthetic.swingup_batch(B)
    run(f"config2 B={B}", BatchedOptimization(B), st, dy, sp)
st, dy, sp = synthetic.swingup_batch(1500, seed=13)
run("lm-heavy compact split", BatchedOptimization(1500, split_launch=True, u_cost_weight=0.0), st, dy, sp)
os.environ["CPMPC_ZGLOBAL"] = "1"
run("lm-heavy forced compact", BatchedOptimization(1500, split_launch=True, u_cost_weight=0.0), st, dy, sp)
del os.environ["CPMPC_ZGLOBAL"]
st, dy, sp, wt = synthetic.sweep_batch(2500)
run("config4 W=160 split", BatchedOptimization(2500, window_length=160), st, dy, sp, weights=wt)
st, dy, sp = synthetic.swingup_batch(20000)
s = BatchedOptimization(20000)
fin, ctl, counts = s.closed_loop(3, st, dy, sp)
print("closed loop split ok", counts, flush=True)
s.close()
st, dy, sp = synthetic.double_swingup_batch(4000)
run("double pendulum split", PlantBatchedOptimization("double", 4000), st, dy, sp)
