"""Generates tests/golden/*.npz from the CPU oracle (oracle/).

The reference's own tests pin no control/cost/iteration vectors for Optimization::Step (the dump in
optimization_test.cc:68-76 is compiled out).  These fixtures are frozen outputs of the oracle RESTATEMENT; the
fixtures produced by the reference's own compiled sources are tests/golden/ref_*.npz (make_golden_ref.py), and
tests/test_reference_pinning.py ties the two together.  The double-pendulum fixture (config 5) has no reference
implementation at all: PARITY UNPINNED for it.  Re-run only when the oracle is intentionally changed:

    python tests/golden/make_golden.py
"""
from __future__ import annotations

import math
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from cart_pole_mpc_b200 import synthetic  # noqa: E402
from oracle import oracle_py  # noqa: E402


def save(name, opt, states, dyn, sp, weights=None):
    res = oracle_py.batch_step(opt, states, dyn, sp, weights=weights)
    np.savez_compressed(
        os.path.join(HERE, name), opt=opt, states=states, dyn=dyn, set_points=sp,
        weights=weights if weights is not None else np.zeros((0, 6)), u=res.u, predicted=res.predicted,
        solution=res.solution, stats=res.stats, iter_log=res.iter_log)
    print(name, "term hist", np.bincount(res.termination, minlength=7).tolist(), "mean iters",
          res.n_iterations.mean())


def main():
    # config 1: the single default solve from the hanging-down state (+ the optimization_test.cc variant)
    x0 = np.array([[0.0, -math.pi / 2, 0.0, 0.0]])
    save("config1_default.npz", oracle_py.opt_array(), x0, np.array([synthetic.DEFAULT_DYN]), np.zeros(1))
    save("config1_test_variant.npz", oracle_py.opt_array(state_spacing=5, max_iterations=10), x0,
         np.array([synthetic.TEST_DYN]), np.zeros(1))
    # model/scratch.py:23-36 parameter set (level start, zero u cost, soft terminal velocities)
    save("scratch_py.npz",
         oracle_py.opt_array(max_iterations=30, absolute_first_derivative_tol=1e-3, u_cost_weight=0.0,
                             b_x_final_cost_weight=5.0, b_x_dot_final_cost_weight=100.0,
                             th_dot_final_cost_weight=100.0),
         np.array([[0.0, 0.0, 0.0, 0.0]]), np.array([synthetic.DEFAULT_DYN]), np.zeros(1))
    # config 2 sample: 64 randomised swing-up instances
    st, dy, sp = synthetic.swingup_batch(64)
    save("config2_sample64.npz", oracle_py.opt_array(), st, dy, sp)
    # config 4 style sweep (default horizon) with per-instance weights
    st, dy, sp, w = synthetic.sweep_batch(32)
    save("sweep_sample32.npz", oracle_py.opt_array(), st, dy, sp, weights=w)

    # config 5: cart-mounted double pendulum (the reference has no implementation: oracle/double_cartpole.hpp)
    st, dy, sp = synthetic.double_swingup_batch(48)
    opt5 = oracle_py.dp_opt_array()
    res = oracle_py.dp_batch_step(opt5, st, dy, sp)
    np.savez_compressed(os.path.join(HERE, "config5_double_sample48.npz"), opt=opt5, states=st, dyn=dy,
                        set_points=sp, u=res.u, predicted=res.predicted, solution=res.solution, stats=res.stats,
                        iter_log=res.iter_log)
    print("config5_double_sample48.npz term hist", np.bincount(res.termination, minlength=7).tolist())

    # closed loop (optimization_test.cc:12-77): 200 steps, controls and final state
    opt = oracle_py.opt_array(state_spacing=5, max_iterations=10)
    states, controls, term, counts, hist, _ = oracle_py.closed_loop(
        opt, x0, np.array([synthetic.TEST_DYN]), np.zeros(1), 200, n_threads=1)
    np.savez_compressed(os.path.join(HERE, "closed_loop_test200.npz"), opt=opt, x0=x0,
                        dyn=np.array([synthetic.TEST_DYN]), controls=controls, final_state=states,
                        last_pred_terminal=term, counts=counts, term_hist=hist)
    print("closed loop counts", counts.tolist(), "hist", hist.tolist(), "final", states)


if __name__ == "__main__":
    main()
