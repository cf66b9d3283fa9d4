"""Generates tests/golden/ref_*.npz from oracle/_ref/libref.so — the REFERENCE's own translation units
(optimization.cc, simulator.cc, mini_opt nonlinear.cc / qp.cc / residual.cc / structs.cc and the generated
single_pendulum_dynamics.hpp) compiled unmodified against oracle/eigen_shim (see that header for what the shim is).

These fixtures are what pins the three graded quantities (optimised controls, final cost, iteration count) to the
reference: tests/test_reference_pinning.py checks the CPU restatement (oracle/) against them on CPU and
tests/test_gpu_parity.py checks the CUDA path against them on the GPU box, where /root/reference does not exist.

Can only be run in the build container (needs /root/reference):

    make -C oracle ref && python tests/golden/make_golden_ref.py
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
    with oracle_py.using_ref():
        res = oracle_py.batch_step(opt, states, dyn, sp, weights=weights)
    np.savez_compressed(
        os.path.join(HERE, name), opt=opt, states=states, dyn=dyn, set_points=sp,
        weights=weights if weights is not None else np.zeros((0, 6)), u=res.u, predicted=res.predicted,
        stats=res.stats, iter_log=res.iter_log)
    print(name, "term hist", np.bincount(res.termination, minlength=7).tolist(), "mean iters",
          res.n_iterations.mean(), "mean ls", res.n_line_search.mean())


def main():
    assert oracle_py.build_ref() is not None, "needs /root/reference (build container only)"
    with oracle_py.using_ref():
        print(oracle_py.lib().oracle_ref_info().decode())

    # config 1: the single default solve from the hanging-down state, the optimization_test.cc:15-25 variant and the
    # model/scratch.py:23-36 parameter set
    x0 = np.array([[0.0, -math.pi / 2, 0.0, 0.0]])
    save("ref_config1_default.npz", oracle_py.opt_array(), x0, np.array([synthetic.DEFAULT_DYN]), np.zeros(1))
    save("ref_config1_test_variant.npz", oracle_py.opt_array(state_spacing=5, max_iterations=10), x0,
         np.array([synthetic.TEST_DYN]), np.zeros(1))
    save("ref_scratch_py.npz",
         oracle_py.opt_array(max_iterations=30, absolute_first_derivative_tol=1e-3, u_cost_weight=0.0,
                             b_x_final_cost_weight=5.0, b_x_dot_final_cost_weight=100.0,
                             th_dot_final_cost_weight=100.0),
         np.array([[0.0, 0.0, 0.0, 0.0]]), np.array([synthetic.DEFAULT_DYN]), np.zeros(1))

    # config 2: the first 256 of the 4096 randomised swing-up instances of the bench
    st, dy, sp = synthetic.swingup_batch(256)
    save("ref_config2_256.npz", oracle_py.opt_array(), st, dy, sp)
    # the optimization_test.cc variant on the same instances
    save("ref_config2_testvariant_128.npz", oracle_py.opt_array(state_spacing=5, max_iterations=10), st[:128],
         dy[:128], sp[:128])
    # per-instance weights + dynamics sweep at the default horizon, and the long horizon of config 4 (W = 160)
    st, dy, sp, w = synthetic.sweep_batch(64)
    save("ref_sweep_64.npz", oracle_py.opt_array(), st, dy, sp, weights=w)
    save("ref_config4_w160_24.npz", oracle_py.opt_array(window_length=160), st[:24], dy[:24], sp[:24], weights=w[:24])

    # closed loop, the reference's own controller test (optimization_test.cc:12-77): 200 warm-started steps
    opt = oracle_py.opt_array(state_spacing=5, max_iterations=10)
    with oracle_py.using_ref():
        states, controls, term, counts, hist, _ = oracle_py.closed_loop(
            opt, x0, np.array([synthetic.TEST_DYN]), np.zeros(1), 200, n_threads=1)
    np.savez_compressed(os.path.join(HERE, "ref_closed_loop_test200.npz"), opt=opt, x0=x0,
                        dyn=np.array([synthetic.TEST_DYN]), controls=controls, final_state=states,
                        last_pred_terminal=term, counts=counts, term_hist=hist)
    print("closed loop counts", counts.tolist(), "hist", hist.tolist(), "final", states)

    # closed loop on 8 randomised plants x 60 steps at the defaults (config 3 in miniature)
    st, dy, sp = synthetic.swingup_batch(8, seed=synthetic.SEED_CONFIG2 + 1)
    with oracle_py.using_ref():
        states, controls, term, counts, hist, _ = oracle_py.closed_loop(oracle_py.opt_array(), st, dy, sp, 60)
    np.savez_compressed(os.path.join(HERE, "ref_closed_loop_8x60.npz"), opt=oracle_py.opt_array(), x0=st, dyn=dy,
                        set_points=sp, controls=controls, final_state=states, last_pred_terminal=term,
                        counts=counts, term_hist=hist)
    print("closed loop 8x60 counts", counts.tolist(), "hist", hist.tolist())


def check_debug():
    """Rebuilds the compiled reference with assertions on and -O1 (make -C oracle ref_debug) and requires it to
    reproduce every committed ref_*.npz bit for bit: no shape assertion of the Eigen-interface shim may fire on the
    path, and the fixtures must not depend on the optimisation level."""
    import subprocess
    subprocess.run(["make", "-C", os.path.join(os.path.dirname(HERE), "..", "oracle"), "ref_debug"], check=True,
                   capture_output=True)
    oracle_py._REF_LIB_PATH = os.path.join(os.path.dirname(os.path.dirname(HERE)), "oracle", "_ref", "libref_debug.so")
    oracle_py._ref_lib = None
    ok = True
    for name in sorted(os.listdir(HERE)):
        if not (name.startswith("ref_") and name.endswith(".npz")):
            continue
        g = np.load(os.path.join(HERE, name))
        with oracle_py.using_ref():
            if "x0" in g.files:
                T = g["controls"].shape[1]
                sp = g["set_points"] if "set_points" in g.files else np.zeros(g["x0"].shape[0])
                nt = 1 if g["x0"].shape[0] == 1 else 0
                _, ctl, _, counts, _, _ = oracle_py.closed_loop(g["opt"], g["x0"], g["dyn"], sp, T, n_threads=nt)
                same = np.array_equal(ctl, g["controls"]) and np.array_equal(counts, g["counts"])
            else:
                w = g["weights"] if g["weights"].size else None
                r = oracle_py.batch_step(g["opt"], g["states"], g["dyn"], g["set_points"], weights=w)
                same = np.array_equal(r.u, g["u"]) and np.array_equal(r.stats, g["stats"])
        print(f"{name}: {'bit-identical' if same else 'DIFFERENT'}")
        ok = ok and same
    return 0 if ok else 1


if __name__ == "__main__":
    if "--check-debug" in sys.argv:
        sys.exit(check_debug())
    main()
