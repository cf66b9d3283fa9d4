#!/usr/bin/env python
"""Profiling target for config 4 (W=160 long horizon, Riccati QP on every iteration): a few cold-start solves.

    ncu --set full --clock-control none --import-source on -k regex:cpmpc_solve_kernel -s 1 -c 1 \
        -o gpurun_out/prof_long python profiles/prof_target_long.py [batch] [steps]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cart_pole_mpc_b200 import BatchedOptimization, synthetic  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
st, dy, sp, wt = synthetic.sweep_batch(B)
solver = BatchedOptimization(B, device=0, window_length=160)
for _ in range(steps):
    solver.reset()
    out = solver.step(st, dy, sp, weights=wt)
print("n_iter mean", out.stats["n_iterations"].mean(), "n_ls mean", out.stats["n_line_search_steps"].mean())
solver.close()
