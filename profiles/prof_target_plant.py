#!/usr/bin/env python
"""Profiling target for the plant-generic kernel (config 5, double pendulum on a cart): a few cold-start solves.

    ncu --set full --clock-control none --import-source on -k regex:plant_solve_kernel -s 1 -c 1 \
        -o gpurun_out/prof_plant python profiles/prof_target_plant.py [batch] [steps]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cart_pole_mpc_b200 import synthetic  # noqa: E402
from cart_pole_mpc_b200.plant import PlantBatchedOptimization  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
states, dyn, sp = synthetic.double_swingup_batch(B)
solver = PlantBatchedOptimization("double", B, device=0)
for _ in range(steps):
    solver.reset()
    out = solver.step(states, dyn, sp)
print("n_iter mean", out.stats["n_iterations"].mean(), "n_ls mean", out.stats["n_line_search_steps"].mean())
solver.close()
