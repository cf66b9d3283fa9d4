#!/usr/bin/env python
"""Profiling target: a few cold-start solves of the headline batch through the C ABI (host buffers), nothing else.

    ncu --set full --clock-control none --import-source on -k regex:cpmpc_solve_kernel -s 2 -c 1 \
        -o gpurun_out/prof python profiles/prof_target.py [batch] [steps]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cart_pole_mpc_b200 import BatchedOptimization, synthetic  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
states, dyn, sp = synthetic.swingup_batch(B)
solver = BatchedOptimization(B, device=0)
for _ in range(steps):
    solver.reset()
    out = solver.step(states, dyn, sp)
print("n_iter mean", out.stats["n_iterations"].mean(), "n_ls mean", out.stats["n_line_search_steps"].mean())
solver.close()
