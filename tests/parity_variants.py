"""Prints parity fractions for unusual configurations (test infrastructure; may import oracle/)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cart_pole_mpc_b200 import BatchedOptimization, synthetic
from oracle import oracle_py as O

def frac(over, B=1024, seed=7, weights=None):
    st, dy, sp = synthetic.swingup_batch(B, seed=seed)
    ora = O.batch_step(O.opt_array(**over), st, dy, sp, want_iter_log=False)
    out = BatchedOptimization(B, **over).step(st, dy, sp)
    g = out.stats["f"] + out.stats["penalty"] * out.stats["equality"]
    rel = np.abs(g - ora.merit) / np.maximum(np.abs(ora.merit), 1e-300)
    return (rel <= 1e-8).mean(), (out.stats["n_iterations"] == ora.n_iterations).mean(), (out.stats["termination_state"] == ora.termination).mean(), np.bincount(ora.termination, minlength=7).tolist(), np.bincount(np.maximum(out.stats["termination_state"],0), minlength=7).tolist()

for over in (dict(), dict(state_spacing=5, max_iterations=10), dict(state_spacing=40), dict(state_spacing=20), dict(state_spacing=1, max_iterations=4), dict(window_length=20, state_spacing=5), dict(max_iterations=1),
             dict(th_final_cost_weight=50.0, b_x_dot_final_cost_weight=20.0, th_dot_final_cost_weight=20.0), dict(b_x_final_cost_weight=-1.0), dict(b_x_final_cost_weight=0.0), dict(u_cost_weight=0.0), dict(u_derivative_cost_weight=0.0)):
    print(over, frac(over))
for B in (1, 3, 9, 67):
    print("B", B, frac(dict(), B=B, seed=5))
