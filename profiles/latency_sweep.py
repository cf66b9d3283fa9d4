#!/usr/bin/env python
"""Batch-of-one latency through cpmpc_step (host buffers) for lane / chunk configurations (CPMPC_LANES, CPMPC_CHUNKS
are read by cpmpc_create):  for l in 4 8 16 32; do CPMPC_LANES=$l CPMPC_CHUNKS=$((l/4)) python profiles/latency_sweep.py; done"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cart_pole_mpc_b200 import BatchedOptimization, synthetic
st, dy, sp = synthetic.swingup_batch(1)
one = BatchedOptimization(1, device=0, low_latency=True)
ts = []
for i in range(80):
    one.reset()
    t0 = time.perf_counter()
    out = one.step(st, dy, sp)
    ts.append(time.perf_counter() - t0)
print("lanes", os.environ.get("CPMPC_LANES"), "cps", os.environ.get("CPMPC_CHUNKS"), "p50 ms", round(float(np.percentile(ts[10:], 50)) * 1e3, 4),
      "merit", float(out.merit[0]), "iters", int(out.stats["n_iterations"][0]))
