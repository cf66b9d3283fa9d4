#!/usr/bin/env python
"""Profiling target for config 3 (fused closed loop, warm-started solves): 16384 plants x 20 control ticks.

    ncu --set full --clock-control none --import-source on -k regex:cpmpc_solve_kernel -c 1 \
        -o gpurun_out/prof_cl python profiles/prof_target_closed_loop.py [plants] [ticks]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cart_pole_mpc_b200 import BatchedOptimization, synthetic  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
T = int(sys.argv[2]) if len(sys.argv) > 2 else 20
st, dy, sp = synthetic.swingup_batch(B, seed=synthetic.SEED_CONFIG2 + 1)
solver = BatchedOptimization(B, device=0)
fin, ctl, counts = solver.closed_loop(T, st, dy, sp)
print("solves", counts[0], "mean iterations", counts[1] / counts[0], "mean merit evaluations", counts[2] / counts[0])
solver.close()
