"""GPU parity tests: the CUDA path, called through the C ABI, against the CPU oracle and the frozen golden vectors.

Tolerances (FP64).  A single SQP iteration from identical variables agrees to ~1e-12 relative; over 8 cold-start
iterations a few percent of swing-up instances are chaotic (the oracle compiled with and without FMA contraction
agrees with ITSELF on only ~98% of config-2 instances at 1e-8), so batch tests assert the north-star
bar: >= 95% of instances within 1e-8 relative final cost, plus iteration-count and termination match rates.
"""
import math
import os

import numpy as np
import pytest

from conftest import load_golden

pytestmark = pytest.mark.gpu

REL_COST_TOL = 1e-8  # north_star: relative final cost vs the reference CPU solver


def _solver(batch, **over):
    from cart_pole_mpc_b200 import BatchedOptimization
    return BatchedOptimization(batch, **over)


def _overrides_from_opt(opt):
    from oracle.oracle_py import OPT_FIELDS
    ints = ("window_length", "state_spacing", "max_iterations")
    return {k: (int(v) if k in ints else float(v)) for k, v in zip(OPT_FIELDS, opt)}


def _rel_cost(out, ora_stats):
    g = out.stats["f"] + out.stats["penalty"] * out.stats["equality"]
    o = ora_stats[:, 3] + ora_stats[:, 5] * ora_stats[:, 4]
    return np.abs(g - o) / np.maximum(np.abs(o), 1e-300)


def _check_batch(out, ora_stats, ora_u, min_frac=0.95):
    rel = _rel_cost(out, ora_stats)
    ok = rel <= REL_COST_TOL
    it = out.stats["n_iterations"] == ora_stats[:, 1].astype(int)
    term = out.stats["termination_state"] == ora_stats[:, 0].astype(int)
    assert ok.mean() >= min_frac, f"only {ok.mean():.4f} within {REL_COST_TOL} relative cost"
    assert it.mean() >= 0.98 and term.mean() >= 0.98
    # where the cost agrees the controls agree too
    du = np.abs(out.u - ora_u).max(axis=1) / np.maximum(np.abs(ora_u).max(axis=1), 1.0)
    assert np.median(du[ok]) < 1e-9
    return ok


# ----------------------------------------------------------------------------------------------------------------------
# golden vectors (configs 1, 2 sample, scratch.py parameters, per-instance weights)
# ----------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["config1_default.npz", "config1_test_variant.npz", "scratch_py.npz"])
def test_single_solve_matches_golden(name):
    g = load_golden(name)
    s = _solver(1, **_overrides_from_opt(g["opt"]))
    s.enable_iteration_log(True)
    out = s.step(g["states"], g["dyn"], g["set_points"])
    assert out.stats["termination_state"][0] == int(g["stats"][0, 0])
    assert out.stats["n_iterations"][0] == int(g["stats"][0, 1])
    assert out.stats["n_line_search_steps"][0] == int(g["stats"][0, 2])
    assert _rel_cost(out, g["stats"])[0] <= REL_COST_TOL
    assert np.allclose(out.u, g["u"], rtol=1e-7, atol=1e-7)
    assert np.allclose(out.predicted_states, g["predicted"], rtol=1e-7, atol=1e-7)
    # per-iteration record: first iteration starts from identical variables -> agreement to round-off
    log = s.get_iteration_log()[0]
    assert np.allclose(log[0, 2:7], g["iter_log"][0, 0, 2:7], rtol=1e-11, atol=1e-11)
    n = int(g["stats"][0, 1])
    assert np.array_equal(log[:n, 7:9], g["iter_log"][0, :n, 7:9])  # step results and merit evaluations/iteration
    assert np.allclose(log[:n, 9::3], g["iter_log"][0, :n, 9::3], rtol=1e-6, atol=1e-9)  # alphas
    sol, has = s.get_previous_solution()
    assert has[0] and np.allclose(sol, g["solution"], rtol=1e-7, atol=1e-7)


def test_config2_sample_matches_golden():
    g = load_golden("config2_sample64.npz")
    s = _solver(64)
    out = s.step(g["states"], g["dyn"], g["set_points"])
    _check_batch(out, g["stats"], g["u"], min_frac=0.9)


def test_per_instance_weights_match_golden():
    g = load_golden("sweep_sample32.npz")
    s = _solver(32)
    out = s.step(g["states"], g["dyn"], g["set_points"], weights=g["weights"])
    _check_batch(out, g["stats"], g["u"], min_frac=0.9)


# ----------------------------------------------------------------------------------------------------------------------
# fixtures produced by the COMPILED REFERENCE (oracle/_ref/libref.so: the reference's own translation units on the
# Eigen-interface shim; tests/golden/make_golden_ref.py).  These pin the CUDA path to the reference's sources, not
# only to the restatement.
# ----------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["ref_config1_default.npz", "ref_config1_test_variant.npz", "ref_scratch_py.npz"])
def test_single_solve_matches_compiled_reference(name):
    g = load_golden(name)
    s = _solver(1, **_overrides_from_opt(g["opt"]))
    s.enable_iteration_log(True)
    out = s.step(g["states"], g["dyn"], g["set_points"])
    assert out.stats["termination_state"][0] == int(g["stats"][0, 0])
    assert out.stats["n_iterations"][0] == int(g["stats"][0, 1])
    assert out.stats["n_line_search_steps"][0] == int(g["stats"][0, 2])
    assert _rel_cost(out, g["stats"])[0] <= REL_COST_TOL
    assert np.allclose(out.u, g["u"], rtol=1e-7, atol=1e-7)
    assert np.allclose(out.predicted_states, g["predicted"], rtol=1e-7, atol=1e-7)
    log = s.get_iteration_log()[0]
    assert np.allclose(log[0, 2:7], g["iter_log"][0, 0, 2:7], rtol=1e-11, atol=1e-11)
    n = int(g["stats"][0, 1])
    assert np.array_equal(log[:n, 7:9], g["iter_log"][0, :n, 7:9])


@pytest.mark.parametrize("name,min_frac", [("ref_config2_256.npz", 0.95), ("ref_config2_testvariant_128.npz", 0.95),
                                           ("ref_sweep_64.npz", 0.93), ("ref_config4_w160_24.npz", 0.85)])
def test_batches_match_compiled_reference(name, min_frac):
    g = load_golden(name)
    B = g["states"].shape[0]
    w = g["weights"] if g["weights"].size else None
    out = _solver(B, **_overrides_from_opt(g["opt"])).step(g["states"], g["dyn"], g["set_points"], weights=w)
    _check_batch(out, g["stats"], g["u"], min_frac=min_frac)


def test_closed_loop_matches_compiled_reference():
    """8 plants x 60 warm-started control ticks at the defaults, run by the compiled reference's Optimization +
    Simulator, against the device-resident closed loop."""
    g = load_golden("ref_closed_loop_8x60.npz")
    B, T = g["x0"].shape[0], g["controls"].shape[1]
    s = _solver(B)
    fin, ctl, counts = s.closed_loop(T, g["x0"], g["dyn"], g["set_points"])
    assert counts[0] == B * T == g["counts"][0]
    close = np.abs(ctl - g["controls"]).max(axis=1) <= 1e-5 * np.maximum(np.abs(g["controls"]).max(axis=1), 1.0)
    assert close.mean() >= 0.75, close.mean()
    assert np.allclose(fin[close], g["final_state"][close], rtol=1e-5, atol=1e-6)
    # the first control tick is a cold solve from identical inputs: all but (at most) one chaotic plant agree
    first = np.abs(ctl[:, 0] - g["controls"][:, 0]) <= 1e-6 * np.maximum(np.abs(g["controls"][:, 0]), 1.0)
    assert first.sum() >= B - 1, first
    # total SQP iterations of the free-running loops: a plant whose trajectory has left the reference's (rounding,
    # amplified by an unstable plant) also iterates differently; the per-tick comparison that is not subject to this is
    # the teacher-forced one below
    assert abs(int(counts[1]) - int(g["counts"][1])) <= 0.10 * g["counts"][1]


# ----------------------------------------------------------------------------------------------------------------------
# seeded batches against the oracle
# ----------------------------------------------------------------------------------------------------------------------
# (OptimizationParams overrides, minimum fraction within 1e-8 relative cost).  The reference's own configurations
# (defaults; optimization_test.cc's S=5) carry the north-star bar of 95%; single shooting (S = W) is more chaotic —
# the oracle disagrees with its own FMA build more often there too — and gets a 90% bar.
SWINGUP_CONFIGS = [
    (dict(), 0.95), (dict(state_spacing=5, max_iterations=10), 0.95), (dict(state_spacing=40), 0.90),
    (dict(state_spacing=20), 0.93), (dict(state_spacing=1, max_iterations=4), 0.95),
    (dict(window_length=20, state_spacing=5), 0.95), (dict(max_iterations=1), 0.99),
]


@pytest.mark.parametrize("over,min_frac", SWINGUP_CONFIGS)
def test_swingup_batch_against_oracle(oracle, over, min_frac):
    from cart_pole_mpc_b200 import synthetic
    B = 1024
    st, dy, sp = synthetic.swingup_batch(B, seed=7)
    ora = oracle.batch_step(oracle.opt_array(**over), st, dy, sp, want_iter_log=False)
    out = _solver(B, **over).step(st, dy, sp)
    _check_batch(out, ora.stats, ora.u, min_frac=min_frac)
    ok = _rel_cost(out, ora.stats) <= REL_COST_TOL
    # predicted states are an open-loop rollout of an unstable plant: tiny control differences are amplified, so the
    # bound is on the typical instance, with a loose cap on the worst one
    dpred = np.abs(out.predicted_states[ok] - ora.predicted[ok]).reshape(int(ok.sum()), -1).max(axis=1)
    assert np.median(dpred) < 1e-9 and np.max(dpred) < 1e-2


def test_first_iteration_agrees_to_roundoff(oracle):
    """One SQP iteration from identical variables: no chaotic amplification yet, every instance must agree."""
    from cart_pole_mpc_b200 import synthetic
    B = 2048
    st, dy, sp = synthetic.swingup_batch(B, seed=11)
    ora = oracle.batch_step(oracle.opt_array(max_iterations=1), st, dy, sp)
    s = _solver(B, max_iterations=1)
    s.enable_iteration_log(True)
    out = s.step(st, dy, sp)
    log = s.get_iteration_log()
    finite = np.isfinite(ora.stats[:, 3]) & np.isfinite(ora.stats[:, 4])
    same_steps = log[:, 0, 8] == ora.iter_log[:, 0, 8]
    assert same_steps.mean() > 0.999
    sel = finite & same_steps
    # f(0), c(0), df/dalpha, dc/dalpha, penalty of iteration 0
    assert np.allclose(log[sel, 0, 2:7], ora.iter_log[sel, 0, 2:7], rtol=1e-9, atol=1e-9)
    rel = _rel_cost(out, ora.stats)
    assert (rel[sel] <= 1e-9).mean() > 0.995
    assert np.median(np.abs(out.u[sel] - ora.u[sel]).max(axis=1)) < 1e-11


def test_ragged_and_tiny_batches(oracle):
    """Batch sizes that do not fill a warp of groups, including a batch of one: every instance must come out
    bit-identical to the same instance solved inside a larger batch, and the larger batch must match the oracle."""
    from cart_pole_mpc_b200 import synthetic
    st, dy, sp = synthetic.swingup_batch(67, seed=5)
    ora = oracle.batch_step(oracle.opt_array(), st, dy, sp, want_iter_log=False)
    big = _solver(67).step(st, dy, sp)
    assert (_rel_cost(big, ora.stats) <= REL_COST_TOL).mean() >= 0.9
    assert np.array_equal(big.stats["n_iterations"], ora.n_iterations)
    for B in (1, 3, 9, 33):
        out = _solver(B).step(st[:B], dy[:B], sp[:B])
        assert np.array_equal(out.u, big.u[:B]) and np.array_equal(out.stats, big.stats[:B])
        assert np.array_equal(out.predicted_states, big.predicted_states[:B])


def test_terminal_weight_variants(oracle):
    """All-cost and all-equality terminal handling (optimization.cc:236-267), and a zero weight (still a cost)."""
    from cart_pole_mpc_b200 import synthetic
    B = 256
    st, dy, sp = synthetic.swingup_batch(B, seed=13)
    # u_cost_weight = 0 leaves only the control-rate Hessian (condition number ~ (2W/pi)^2) and drives some instances
    # into Levenberg-Marquardt mode / MAX_LAMBDA: the most chaotic variant, hence the lower bar.
    for over, min_frac in ((dict(th_final_cost_weight=50.0, b_x_dot_final_cost_weight=20.0,
                                 th_dot_final_cost_weight=20.0), 0.93),
                           (dict(b_x_final_cost_weight=-1.0), 0.95),
                           (dict(b_x_final_cost_weight=0.0), 0.95),
                           (dict(u_cost_weight=0.0), 0.80), (dict(u_derivative_cost_weight=0.0), 0.93)):
        ora = oracle.batch_step(oracle.opt_array(**over), st, dy, sp, want_iter_log=False)
        out = _solver(B, **over).step(st, dy, sp)
        _check_batch(out, ora.stats, ora.u, min_frac=min_frac)


def test_non_finite_line_search_is_reported_like_the_reference(oracle):
    """Instances whose full step blows the plant up to NaN: FAILURE_NON_FINITE_COST repeats deterministically until
    max_iterations and leaves the variables unchanged (nonlinear.cc:395-397)."""
    from cart_pole_mpc_b200 import synthetic
    B = 4096
    st, dy, sp = synthetic.swingup_batch(B)
    ora = oracle.batch_step(oracle.opt_array(), st, dy, sp)
    bad = np.where(ora.stats[:, 7] >= 3)[0]
    assert bad.size > 0
    s = _solver(B)
    out = s.step(st, dy, sp)
    agree = (out.stats["n_failed_line_searches"][bad] == ora.stats[bad, 7].astype(int)).mean()
    assert agree >= 0.8
    assert np.all(np.isfinite(out.u)) and np.all(np.abs(out.u) <= 300.0)
    assert np.all(np.isfinite(out.stats["f"])) and np.all(np.isfinite(out.stats["equality"]))


# ----------------------------------------------------------------------------------------------------------------------
# live comparison with the COMPILED REFERENCE on the box (oracle/_ref/libref.so travels with the snapshot): the same
# parameter space as above, now against the reference's own sources rather than the restatement.
# ----------------------------------------------------------------------------------------------------------------------
REF_VARIANTS = [
    (dict(), 0.95), (dict(state_spacing=5, max_iterations=10), 0.95), (dict(state_spacing=40), 0.90),
    (dict(state_spacing=1, max_iterations=4), 0.95), (dict(window_length=20, state_spacing=5), 0.95),
    (dict(window_length=80, state_spacing=10), 0.90), (dict(control_dt=0.02), 0.90),
    (dict(u_guess_sinusoid_amplitude=0.0), 0.95), (dict(equality_penalty_initial=10.0), 0.95),
    (dict(th_final_cost_weight=50.0, b_x_dot_final_cost_weight=20.0, th_dot_final_cost_weight=20.0), 0.93),
    (dict(b_x_final_cost_weight=-1.0), 0.95), (dict(b_x_final_cost_weight=0.0), 0.95),
    (dict(u_derivative_cost_weight=0.0), 0.93), (dict(max_iterations=1), 0.99),
]


@pytest.mark.parametrize("over,min_frac", REF_VARIANTS)
def test_variants_against_compiled_reference_live(oracle, over, min_frac):
    if not oracle.ref_available():
        pytest.skip("oracle/_ref/libref.so not present")
    from cart_pole_mpc_b200 import synthetic
    # 1024 instances: the sampling noise of a pass fraction near 0.95 is 0.7 % (1.4 % on the 256 instances this test
    # used before, which put the single-shooting variant within a standard deviation of its bar: profiles/parity_ab.py
    # prints the fractions of a build on this sample)
    B = 1024
    st, dy, sp = synthetic.swingup_batch(B, seed=17)
    sp = sp + np.linspace(-0.3, 0.3, B)  # non-zero set points as well
    with oracle.using_ref():
        ref = oracle.batch_step(oracle.opt_array(**over), st, dy, sp, want_iter_log=False)
    out = _solver(B, **over).step(st, dy, sp)
    _check_batch(out, ref.stats, ref.u, min_frac=min_frac)
    ok = _rel_cost(out, ref.stats) <= REL_COST_TOL
    dpred = np.abs(out.predicted_states[ok] - ref.predicted[ok]).reshape(int(ok.sum()), -1).max(axis=1)
    assert np.median(dpred) < 1e-9


@pytest.mark.parametrize("over,min_frac", [
    (dict(window_length=400, state_spacing=10), 0.6),                      # 40 segments > 32 lanes, 4 s horizon
    (dict(window_length=330, state_spacing=10, max_iterations=4), 0.9),    # 33 segments: ragged over the lanes
    (dict(window_length=64, state_spacing=1, max_iterations=3), 0.9),      # one knot per segment, 64 segments
    (dict(window_length=100, state_spacing=100, max_iterations=3), 0.9),   # single shooting with a 100-knot segment
])
def test_maximum_sizes_against_compiled_reference_live(oracle, over, min_frac):
    """Problem shapes at the edges of the lane mapping (more segments than lanes, one-knot segments, one huge
    segment): must agree with the reference's own code, not just run."""
    if not oracle.ref_available():
        pytest.skip("oracle/_ref/libref.so not present")
    from cart_pole_mpc_b200 import synthetic
    B = 16
    st, dy, sp = synthetic.swingup_batch(B, seed=5)
    with oracle.using_ref():
        ref = oracle.batch_step(oracle.opt_array(**over), st, dy, sp, want_iter_log=False)
    out = _solver(B, **over).step(st, dy, sp)
    rel = _rel_cost(out, ref.stats)
    assert (rel <= REL_COST_TOL).mean() >= min_frac
    assert (out.stats["n_iterations"] == ref.n_iterations).mean() >= 0.9
    assert (out.stats["termination_state"] == ref.termination).mean() >= 0.9


def test_warm_started_step_against_compiled_reference_live(oracle):
    """Step 2 of a receding horizon: both sides are handed the same previous solution (SetPreviousSolution,
    optimization.hpp:86-89) and the same advanced plant state; covers the shift of the controls and u_prev
    (optimization.cc:44-57,289-291) against the reference's own code."""
    if not oracle.ref_available():
        pytest.skip("oracle/_ref/libref.so not present")
    from cart_pole_mpc_b200 import synthetic
    B = 256
    over = dict(state_spacing=5, max_iterations=10)
    opt = oracle.opt_array(**over)
    st, dy, sp = synthetic.swingup_batch(B, seed=23)
    first = oracle.batch_step(opt, st, dy, sp, want_iter_log=False)  # restatement: it exposes the solution vector
    with oracle.using_ref():
        st2 = np.stack([oracle.sim_step(dy[i], st[i], 0.01, first.u[i, 0]) for i in range(B)])
        ref2 = oracle.batch_step(opt, st2, dy, sp, prev=first.solution, has_prev=np.ones(B, np.uint8),
                                 want_iter_log=False)
    s = _solver(B, **over)
    s.set_previous_solution(first.solution)
    out2 = s.step(st2, dy, sp)
    _check_batch(out2, ref2.stats, ref2.u, min_frac=0.93)
    assert ref2.n_iterations.min() < 10  # some warm starts exit early: the exit tests are exercised too


# ----------------------------------------------------------------------------------------------------------------------
# warm start / Reset / SetPreviousSolution semantics (optimization.cc:44-57,84-85,289-291; optimization.hpp:83-89)
# ----------------------------------------------------------------------------------------------------------------------
def test_warm_started_second_step_matches_oracle(oracle):
    from cart_pole_mpc_b200 import synthetic, sim_step
    B = 256
    over = dict(state_spacing=5, max_iterations=10)
    opt = oracle.opt_array(**over)
    st, dy, sp = synthetic.swingup_batch(B, seed=3)
    s = _solver(B, **over)
    out1 = s.step(st, dy, sp)
    ora1 = oracle.batch_step(opt, st, dy, sp, want_iter_log=False)
    ok1 = _rel_cost(out1, ora1.stats) <= REL_COST_TOL
    # advance the plants with the ORACLE's control so both sides see identical inputs on step 2
    st2 = np.stack([oracle.sim_step(dy[i], st[i], 0.01, ora1.u[i, 0]) for i in range(B)])
    s.set_previous_solution(ora1.solution)
    out2 = s.step(st2, dy, sp)
    ora2 = oracle.batch_step(opt, st2, dy, sp, prev=ora1.solution, has_prev=np.ones(B, np.uint8), want_iter_log=False)
    _check_batch(out2, ora2.stats, ora2.u, min_frac=0.93)
    # batched GPU simulator == oracle simulator
    st2_gpu = sim_step(st, dy, ora1.u[:, 0], 0.01)
    assert np.allclose(st2_gpu, st2, rtol=1e-12, atol=1e-13)
    assert ok1.mean() > 0.9


def test_reset_and_masks():
    from cart_pole_mpc_b200 import synthetic
    B = 16
    st, dy, sp = synthetic.swingup_batch(B, seed=21)
    s = _solver(B)
    _, has = s.get_previous_solution()
    assert not has.any()
    cold = s.step(st, dy, sp)
    sol, has = s.get_previous_solution()
    assert has.all() and np.array_equal(sol[:, -40:], cold.u)
    mask = np.zeros(B, np.uint8)
    mask[::2] = 1
    s.reset(mask)
    _, has = s.get_previous_solution()
    assert np.array_equal(has, mask == 0)
    s.reset()
    again = s.step(st, dy, sp)
    assert np.array_equal(again.u, cold.u), "a reset controller must reproduce the cold solve bit for bit"
    assert np.array_equal(again.stats, cold.stats)


def test_device_pointer_path_equals_host_path():
    import torch
    from cart_pole_mpc_b200 import synthetic, _capi
    B = 128
    st, dy, sp = synthetic.swingup_batch(B, seed=17)
    s = _solver(B)
    host = s.step(st, dy, sp)
    s.reset()
    dev = torch.device("cuda", 0)
    u = torch.empty((B, 40), dtype=torch.float64, device=dev)
    pred = torch.empty((B, 40, 4), dtype=torch.float64, device=dev)
    stats = torch.empty((B, 56), dtype=torch.uint8, device=dev)
    s.step_device(torch.from_numpy(st).to(dev), torch.from_numpy(dy).to(dev), torch.from_numpy(sp).to(dev), u,
                  pred_out=pred, stats_out=stats)
    torch.cuda.synchronize()
    assert np.array_equal(u.cpu().numpy(), host.u)
    assert np.array_equal(pred.cpu().numpy(), host.predicted_states)
    assert np.array_equal(np.frombuffer(stats.cpu().numpy().tobytes(), dtype=_capi.STATS_DTYPE), host.stats)


# ----------------------------------------------------------------------------------------------------------------------
# closed loop (config 3 in miniature, and the reference's own end-to-end test)
# ----------------------------------------------------------------------------------------------------------------------
def test_reference_closed_loop_test_through_pypendulum_mirror():
    """optimization_test.cc:12-77 run through the pypendulum-compatible API, with the reference's assertions."""
    from cart_pole_mpc_b200 import pypendulum as pp
    op = pp.OptimizationParams()
    op.control_dt, op.window_length, op.state_spacing, op.max_iterations = 0.01, 40, 5, 10
    dyn = pp.SingleCartPoleParams(1.0, 0.1, 0.25, 9.81, 0.03, 0.1, 0.13, 0.8, 100.0)
    sim = pp.Simulator()
    sim.set_state(pp.SingleCartPoleState(0.0, -math.pi / 2, 0.0, 0.0))
    opt = pp.Optimization(op)
    g = load_golden("closed_loop_test200.npz")
    controls = []
    for t in range(200):
        out = opt.step(sim.get_state(), dyn, 0.0)
        assert out.stats["termination_state"] not in (4, 5)  # MAX_LAMBDA, QP_INDEFINITE
        terminal = out.predicted_states[-1]
        if t > 20:
            assert abs(terminal.b_x_dot) < 1e-4 and abs(terminal.th_1_dot) < 1e-4
            assert abs(terminal.th_1 - math.pi / 2) < 1e-4
        controls.append(out.u[0])
        sim.step(dyn, op.control_dt, out.u[0], pp.Vector2(0, 0), pp.Vector2(0, 0))
    fin = sim.get_state()
    assert abs(fin.b_x_dot) < 1e-4 and abs(fin.th_1_dot) < 1e-3 and abs(fin.th_1 - math.pi / 2) < 1e-4
    # and the whole control sequence tracks the oracle's closed loop
    assert np.allclose(controls, g["controls"][0], rtol=1e-5, atol=1e-5)
    assert "termination =" in out.solver_summary()


def test_device_resident_closed_loop_matches_oracle(oracle):
    from cart_pole_mpc_b200 import synthetic
    B, T = 64, 40
    over = dict(state_spacing=5, max_iterations=10)
    st, dy, sp = synthetic.swingup_batch(B, seed=20241102)
    fin_o, ctl_o, _, counts_o, hist_o, _ = oracle.closed_loop(oracle.opt_array(**over), st, dy, sp, T)
    s = _solver(B, **over)
    fin_g, ctl_g, counts_g = s.closed_loop(T, st, dy, sp)
    assert counts_g[0] == B * T == counts_o[0]
    # trajectories: the overwhelming majority of plants follow the oracle's closed loop closely
    close = np.abs(ctl_g - ctl_o).max(axis=1) <= 1e-5 * np.maximum(np.abs(ctl_o).max(axis=1), 1.0)
    assert close.mean() >= 0.85, close.mean()
    assert np.allclose(fin_g[close], fin_o[close], rtol=1e-5, atol=1e-6)
    assert abs(int(counts_g[1]) - int(counts_o[1])) <= 0.02 * counts_o[1]
    # stepping the same handle again continues from the device-resident warm start
    _, has = s.get_previous_solution()
    assert has.all()


# ----------------------------------------------------------------------------------------------------------------------
# size-independent properties at BASELINE.json's full size (config 2: 4096 instances)
# ----------------------------------------------------------------------------------------------------------------------
def test_full_size_properties():
    from cart_pole_mpc_b200 import synthetic
    B = 4096
    st, dy, sp = synthetic.swingup_batch(B)
    s = _solver(B)
    out = s.step(st, dy, sp)
    assert np.all(np.isfinite(out.u)) and np.all(np.abs(out.u) <= 300.0)          # retraction clamp
    assert np.all((out.stats["n_iterations"] >= 1) & (out.stats["n_iterations"] <= 8))
    ls, it = out.stats["n_line_search_steps"], out.stats["n_iterations"]
    assert np.all(ls >= it) and np.all(ls <= 6 * it)                                  # 1..6 merit evaluations each
    assert np.all(out.stats["penalty"] >= 1.0)                                        # penalty never decreases
    th = out.predicted_states[:, :, 1]
    assert np.all((th > -math.pi - 1e-12) & (th <= math.pi + 1e-12))                   # mod_pi after every knot
    sol, _ = s.get_previous_solution()
    assert np.array_equal(sol[:, -40:], out.u)
    assert np.all(np.abs(sol[:, 0:20:4]) <= 5.0)                                      # b_x clamp on shooting states
    # determinism: the same inputs give the same bits
    s.reset()
    out2 = s.step(st, dy, sp)
    assert np.array_equal(out.u, out2.u) and np.array_equal(out.stats, out2.stats)
    # batch-composition independence: instance i alone == instance i inside the batch
    idx = [0, 1, 777, 4095]
    single = _solver(len(idx)).step(st[idx], dy[idx], sp[idx])
    assert np.array_equal(single.u, out.u[idx])


def test_low_latency_mapping_matches_oracle(oracle):
    """CPMPC_FLAG_LOW_LATENCY (two lanes per shooting segment, chunked linearisation) against the oracle: same parity
    bar as the default mapping; it is NOT required to be bitwise equal to the default mapping."""
    from cart_pole_mpc_b200 import BatchedOptimization, synthetic
    for over in (dict(), dict(state_spacing=5, max_iterations=10)):
        B = 512
        st, dy, sp = synthetic.swingup_batch(B, seed=11)
        ora = oracle.batch_step(oracle.opt_array(**over), st, dy, sp, want_iter_log=False)
        out = BatchedOptimization(B, low_latency=True, **over).step(st, dy, sp)
        rel = np.abs(out.merit - ora.merit) / np.maximum(np.abs(ora.merit), 1e-300)
        assert (rel <= REL_COST_TOL).mean() >= 0.95
        assert (out.stats["n_iterations"] == ora.n_iterations).mean() >= 0.98


@pytest.mark.parametrize("over", [dict(), dict(state_spacing=5, max_iterations=10), dict(state_spacing=40),
                                  dict(window_length=80, state_spacing=10), dict(window_length=20, state_spacing=2)])
def test_min_latency_mapping_matches_oracle(oracle, over):
    """CPMPC_FLAG_MIN_LATENCY (one warp per instance, up to 32 / segments linearisation chunks per segment, ragged
    chunk lengths included) against the oracle, same parity bar as the default mapping; and a batch of one equals the
    same instance inside a batch (the mapping, not the batch composition, decides the rounding)."""
    from cart_pole_mpc_b200 import BatchedOptimization, synthetic
    B = 256
    st, dy, sp = synthetic.swingup_batch(B, seed=11)
    ora = oracle.batch_step(oracle.opt_array(**over), st, dy, sp, want_iter_log=False)
    out = BatchedOptimization(B, min_latency=True, **over).step(st, dy, sp)
    rel = np.abs(out.merit - ora.merit) / np.maximum(np.abs(ora.merit), 1e-300)
    assert (rel <= REL_COST_TOL).mean() >= (0.90 if over.get("state_spacing") == 40 else 0.95)
    assert (out.stats["n_iterations"] == ora.n_iterations).mean() >= 0.98
    one = BatchedOptimization(1, min_latency=True, **over).step(st[3:4], dy[3:4], sp[3:4])
    assert np.array_equal(one.u[0], out.u[3]) and np.array_equal(one.stats[0], out.stats[3])


def test_sharding_two_handles_equal_one(oracle):
    """Multi-GPU data path in miniature: two handles over the two halves == one handle over the whole batch."""
    from cart_pole_mpc_b200 import synthetic
    from cart_pole_mpc_b200.sharding import shard_range
    B = 300
    st, dy, sp = synthetic.swingup_batch(B, seed=9)
    whole = _solver(B).step(st, dy, sp)
    parts = []
    for r in range(2):
        first, count = shard_range(B, r, 2)
        s_, d_, p_ = synthetic.swingup_batch(count, seed=9, first=first)
        parts.append(_solver(count).step(s_, d_, p_))
    assert np.array_equal(np.concatenate([p.u for p in parts]), whole.u)
    assert np.array_equal(np.concatenate([p.stats for p in parts]), whole.stats)


def test_fp64_peak_probe_is_plausible():
    from cart_pole_mpc_b200 import measure_fp64_peak
    t = measure_fp64_peak(0)
    assert 10.0 < t < 60.0, t   # B200 FP64 FMA pipe: ~37-40 TFLOP/s nominal


def test_cpp_facade_runs_the_reference_closed_loop_test(tmp_path):
    """optimization_test.cc:12-77 compiled against host/optimization.hpp (the C++ drop-in) and run on the GPU."""
    import subprocess
    from test_host_logic import _build_facade_test
    exe = _build_facade_test(tmp_path)
    proc = subprocess.run([str(exe)], capture_output=True, text=True, timeout=600)
    assert proc.returncode == 0, proc.stdout + proc.stderr
    assert "FACADE_CLOSED_LOOP_OK" in proc.stdout
    # NLSSolverOutputs::ToString of the C++ facade == solver_summary() of the Python mirror for the same first solve
    # (same kernel, batch of one): the two host layers format the reference's summary identically
    cpp_summary = proc.stdout.split("SUMMARY_BEGIN\n")[1].split("SUMMARY_END")[0]
    from cart_pole_mpc_b200 import pypendulum as pp
    op = pp.OptimizationParams()
    op.control_dt, op.window_length, op.state_spacing, op.max_iterations = 0.01, 40, 5, 10
    dyn = pp.SingleCartPoleParams(1.0, 0.1, 0.25, 9.81, 0.03, 0.1, 0.13, 0.8, 100.0)
    out = pp.Optimization(op).step(pp.SingleCartPoleState(0.0, -math.pi / 2, 0.0, 0.0), dyn, 0.0)
    assert out.solver_summary() == cpp_summary


# ----------------------------------------------------------------------------------------------------------------------
# round 2: solve-by-solve and element-wise comparisons with the COMPILED REFERENCE (oracle/ref_glue_debug.cc)
# ----------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("over,min_frac", [(dict(), 0.98), (dict(state_spacing=5, max_iterations=10), 0.98)])
def test_teacher_forced_closed_loop_against_compiled_reference(oracle, over, min_frac):
    """Config 3, teacher-forced: the compiled reference runs optimization_test.cc:38-61 for 256 plants x 100 ticks and
    records what every Step was given; the GPU re-solves each of the 25600 ticks from the reference's plant state and
    warm start (cpmpc_set_previous_solution + cpmpc_step) and is compared solve by solve.  Free-running trajectories
    (test_closed_loop_matches_compiled_reference) mix solver disagreement with the chaotic divergence of an unstable
    plant; this does not.  Two CPU implementations agree on 99.8 % here (tests/parity_floor.py); measured on B200:
    99.5 % (defaults), 99.1 % (S=5)."""
    if not oracle.ref_available():
        pytest.skip("oracle/_ref/libref.so not present")
    from cart_pole_mpc_b200 import synthetic
    P, T = 256, 100
    st, dy, sp = synthetic.swingup_batch(P, seed=synthetic.SEED_CONFIG2 + 1)
    tr = oracle.closed_loop_trace(oracle.opt_array(**over), st, dy, sp, T)
    B = P * T
    s = _solver(B, **over)
    s.set_previous_solution(tr.prev.reshape(B, -1), mask=np.tile((np.arange(T) > 0).astype(np.uint8), P))
    out = s.step(tr.states.reshape(B, 4), np.repeat(dy, T, axis=0), np.repeat(sp, T))
    ref_stats = tr.stats.reshape(B, -1)
    ok = _rel_cost(out, ref_stats) <= REL_COST_TOL
    assert ok.mean() >= min_frac, ok.mean()
    assert ok.reshape(P, T)[:, 1:].mean() >= min_frac, "warm-started ticks"
    assert (out.stats["n_iterations"] == ref_stats[:, 1].astype(int)).mean() >= 0.995
    assert (out.stats["termination_state"] == ref_stats[:, 0].astype(int)).mean() >= 0.995
    ru = tr.u.reshape(B, -1)
    du = np.abs(out.u - ru).max(axis=1) / np.maximum(np.abs(ru).max(axis=1), 1.0)
    assert np.median(du[ok]) < 1e-10
    # the control that is actually applied to the plant
    assert (np.abs(out.u[:, 0] - ru[:, 0]) <= 1e-6 * np.maximum(np.abs(ru[:, 0]), 1.0)).mean() >= 0.99
    assert ref_stats[:, 1].min() < ref_stats[:, 1].max()  # early exits and full-length solves both occur


def _check_blocks(oracle, s, st, dy, sp, w, opt, lam, method, tol_dx):
    W, S = int(opt[1]), int(opt[2])
    nseg, NS = W // S, W // S + 1
    g = s.debug_first_qp(st, dy, sp, weights=w, lam=lam, qp_method=method)
    checked = 0
    for i in range(st.shape[0]):
        o_i = opt.copy()
        if w is not None:
            o_i[8:14] = w[i]
        q = oracle.linearize_qp(o_i, st[i], dy[i], sp[i], lam=lam)
        if q.status != 0:
            continue
        assert g["qp_ok"][i]
        checked += 1
        assert np.isclose(g["f0"][i], q.f, rtol=1e-13) and np.isclose(g["eq0"][i], q.equality, rtol=1e-12)
        for k in range(nseg):
            rows = q.A[4 * k:4 * k + 4]
            phi, gam = rows[:, 4 * k:4 * k + 4], rows[:, 4 * NS + k * S:4 * NS + (k + 1) * S]
            assert np.abs(g["Phi"][i, k] - phi).max() <= 1e-12 * max(1.0, np.abs(phi).max())
            assert np.abs(g["Gam"][i][:, k * S:(k + 1) * S] - gam).max() <= 1e-12 * max(1.0, np.abs(gam).max())
            assert np.abs(g["e"][i, k] - q.b[4 * k:4 * k + 4]).max() <= 1e-12
        dx = np.concatenate([g["dxs"][i].ravel(), g["du"][i]])
        assert np.abs(dx - q.dx).max() <= tol_dx * max(1.0, np.abs(q.dx).max()), (i, method, lam)
    assert checked >= st.shape[0] - 1


def test_first_qp_blocks_match_compiled_reference(oracle):
    """a3 and a6 element-wise: Phi / Gamma / e of the first linearisation against the block rows of the reference's
    A_eq / b_eq (CreateDynamicalConstraint, optimization.cc:99-160) at 1e-12, and the QP step of both device methods
    (condensed form, segment-wise Riccati recursion; undamped and Levenberg-Marquardt damped) against
    QPNullSpaceSolver::Solve (qp.cc:679-729) on the reference's own QP at 1e-10.  Measured on B200: Phi/Gamma 2e-14,
    dx 2e-13 .. 1e-12 (W=40), 6e-14 (W=160)."""
    if not oracle.ref_available():
        pytest.skip("oracle/_ref/libref.so not present")
    from cart_pole_mpc_b200 import synthetic
    st, dy, sp = synthetic.swingup_batch(32, seed=31)
    for over in (dict(), dict(state_spacing=5, max_iterations=10)):
        opt = oracle.opt_array(**over)
        s = _solver(32, **over)
        for lam, method in ((0.0, 1), (0.0, 2), (1e-2, 1), (0.5, 1)):
            _check_blocks(oracle, s, st, dy, sp, None, opt, lam, method, 1e-10)
        # the debug view leaves the warm start alone, and a later Step() is the ordinary cold solve
        _, has = s.get_previous_solution()
        assert not has.any()
    st, dy, sp, w = synthetic.sweep_batch(12)
    opt = oracle.opt_array(window_length=160)
    s = _solver(12, window_length=160)
    for lam in (0.0, 1e-2):
        _check_blocks(oracle, s, st, dy, sp, w, opt, lam, 1, 1e-10)
    with pytest.raises(Exception):
        s.debug_first_qp(st, dy, sp, weights=w, qp_method=2)  # no condensed scratch on long horizons


def test_long_horizon_512_against_compiled_reference_and_cpu_floor(oracle):
    """Config 4 (W=160, per-instance dynamics / weights / set points), 512 cold solves.  Eight SQP iterations on a
    1.6 s horizon amplify round-off: two CPU implementations of the reference algorithm agree with each other on
    96.1 % of these instances at 1e-8 and the FMA-contracted CPU port on 94.9 % (tests/parity_floor.py,
    profiles/r02_parity_floor.json); the first iteration agrees to 1e-14 on every instance and every failing instance
    first differs in a LATER iteration (profiles/r02_b_parity_diagnose.json).  The bar is therefore the floor itself:
    the GPU must be no worse than an FMA-contracted CPU port of the same algorithm, minus the binomial noise of 512
    samples, and never below 90 %; iteration counts and termination states must match."""
    if not oracle.ref_available():
        pytest.skip("oracle/_ref/libref.so not present")
    import importlib.util
    from cart_pole_mpc_b200 import synthetic
    spec = importlib.util.spec_from_file_location("parity_floor", os.path.join(os.path.dirname(__file__), "parity_floor.py"))
    floor = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(floor)
    B = 512
    over = dict(window_length=160)
    st, dy, sp, w = synthetic.sweep_batch(B)
    opt = oracle.opt_array(**over)
    with oracle.using_ref():
        ref = oracle.batch_step(opt, st, dy, sp, weights=w, want_iter_log=True)
    fma = floor.FmaPort().batch_step(opt, st, dy, sp, weights=w, want_iter_log=False)
    floor_frac = floor.agree(fma.stats, ref.stats)["within_1e-8"]
    s = _solver(B, **over)
    s.enable_iteration_log(True)
    out = s.step(st, dy, sp, weights=w)
    log = s.get_iteration_log()
    rel = _rel_cost(out, ref.stats)
    frac = (rel <= REL_COST_TOL).mean()
    assert frac >= max(0.90, floor_frac - 0.03), (frac, floor_frac)
    assert (rel <= 1e-6).mean() >= 0.95
    assert (out.stats["n_iterations"] == ref.n_iterations).mean() >= 0.995
    assert (out.stats["termination_state"] == ref.termination).mean() >= 0.995
    # iteration 0 starts from identical variables: f(0), c(0), d_f, d_eq agree to round-off on EVERY instance (d_f is a
    # sum with cancellation: 1e-10 relative on the worst of 512, 1e-14 typically)
    d0 = np.abs(log[:, 0, 2:6] - ref.iter_log[:, 0, 2:6]) / np.maximum(np.abs(ref.iter_log[:, 0, 2:6]), 1e-300)
    assert d0.max() < 1e-8 and np.median(d0.max(axis=1)) < 1e-12, (d0.max(), np.median(d0.max(axis=1)))


def test_two_live_handles_with_different_block_shapes():
    """the dynamic shared-memory attribute belongs to the kernel instantiation, not to a handle.
    A large handle (5 warps per block) must still launch after a batch-of-one handle of the same instantiation was
    created."""
    from cart_pole_mpc_b200 import synthetic
    B = 8192
    st, dy, sp = synthetic.swingup_batch(B, seed=2)
    big = _solver(B)
    first = big.step(st, dy, sp)
    small = _solver(1)
    one = small.step(st[:1], dy[:1], sp[:1])
    big.reset()
    again = big.step(st, dy, sp)      # failed with cudaErrorInvalidValue before the fix
    assert np.array_equal(first.u, again.u)
    assert np.array_equal(one.u[0], first.u[0])
    small.reset()
    assert np.array_equal(small.step(st[:1], dy[:1], sp[:1]).u, one.u)


def test_rejected_instance_keeps_its_warm_start():
    """the reference throws inside Solve() for a non-finite cost at the linearisation point
    (nonlinear.cc:356-357) BEFORE previous_solution_ is assigned (optimization.cc:84-85): the rejected instance must
    keep the warm start it had, and the Python mirror must raise like the C++ facade does."""
    from cart_pole_mpc_b200 import pypendulum as pp, synthetic
    B = 8
    st, dy, sp = synthetic.swingup_batch(B, seed=4)
    s = _solver(B)
    s.step(st, dy, sp)
    sol0, has0 = s.get_previous_solution()
    bad = st.copy()
    bad[3, 2] = np.nan
    out = s.step(bad, dy, sp)
    assert out.stats["termination_state"][3] == -1
    assert (out.stats["termination_state"][np.arange(B) != 3] >= 0).all()
    sol1, has1 = s.get_previous_solution()
    assert has1.all() and np.array_equal(sol1[3], sol0[3]), "warm start of the rejected instance must be untouched"
    assert not np.array_equal(sol1[2], sol0[2])
    # a never-solved instance that is rejected stays cold
    s2 = _solver(1)
    s2.step(bad[3:4], dy[3:4], sp[3:4])
    assert not s2.get_previous_solution()[1][0]
    opt = pp.Optimization(pp.OptimizationParams())
    dyn = pp.SingleCartPoleParams(*dy[0])
    with pytest.raises(RuntimeError):
        opt.step(pp.SingleCartPoleState(0.0, -1.5, float("nan"), 0.0), dyn, 0.0)
    ok = opt.step(pp.SingleCartPoleState(0.0, -1.5, 0.0, 0.0), dyn, 0.0)   # and the controller is still usable, cold
    assert ok.previous_solution == []


def test_native_pypendulum_runs_scratch_py(oracle):
    """The pybind11 module `pypendulum` (host/pypendulum_ext.cc: the binding list of wrapper/wrapper.cc:40-102 over the
    C++ facade) driven by model/scratch.py:14-40's call sequence, and by the closed loop of optimization_test.cc with
    Simulator; its solver_summary() must be byte-equal to the reference's own ToString of the same log, and equal to
    the Python mirror's."""
    import sys
    from cart_pole_mpc_b200 import build, pypendulum as mirror
    build.build_pypendulum()
    sys.path.insert(0, build.NATIVE_DIR)
    try:
        import pypendulum as pp
    finally:
        sys.path.pop(0)
    # model/scratch.py:22-38
    params = pp.SingleCartPoleParams(1.0, 0.1, 0.25, 9.81, 0.05, 0.1, 0.02, 0.8, 100.0)
    x0 = pp.SingleCartPoleState(0.0, 0.0, 0.0, 0.0)
    op = pp.OptimizationParams()
    op.max_iterations = 30
    op.state_spacing = 10
    op.window_length = 40
    op.absolute_first_derivative_tol = 1.0e-3
    op.u_guess_sinusoid_amplitude = 10.0
    op.u_cost_weight = 0.0
    op.b_x_final_cost_weight = 5.0
    op.th_final_cost_weight = -1.0
    op.b_x_dot_final_cost_weight = 100.0
    op.th_dot_final_cost_weight = 100.0
    opt = pp.Optimization(op)
    outputs = opt.step(x0, params, 0.0)
    text = outputs.solver_summary()
    assert text.startswith("Iteration # 0, state = NOMINAL, lambda = 0\n") and "  termination = " in text
    assert len(outputs.u) == 40 and len(outputs.predicted_states) == 40
    assert all(math.isfinite(s.th_1) and math.isfinite(s.b_x_dot) for s in outputs.predicted_states)
    # the same call sequence through the Python mirror: same kernel, same log, same text
    mo = mirror.OptimizationParams()
    for k in ("max_iterations", "state_spacing", "window_length", "absolute_first_derivative_tol",
              "u_guess_sinusoid_amplitude", "u_cost_weight", "b_x_final_cost_weight", "th_final_cost_weight",
              "b_x_dot_final_cost_weight", "th_dot_final_cost_weight"):
        setattr(mo, k, getattr(op, k))
    m_out = mirror.Optimization(mo).step(mirror.SingleCartPoleState(0.0, 0.0, 0.0, 0.0),
                                         mirror.SingleCartPoleParams(1.0, 0.1, 0.25, 9.81, 0.05, 0.1, 0.02, 0.8, 100.0), 0.0)
    assert m_out.solver_summary() == text
    assert list(outputs.u) == list(m_out.u)
    if oracle.ref_available():
        n = int(m_out.stats["n_iterations"])
        assert oracle.solver_summary(int(m_out.stats["termination_state"]), m_out._iter_log[:n]) == text
    # Simulator + Optimization closed loop (optimization_test.cc:24-61), 30 ticks
    sim = pp.Simulator()
    op2 = pp.OptimizationParams()
    op2.state_spacing, op2.max_iterations = 5, 10
    dyn = pp.SingleCartPoleParams(1.0, 0.1, 0.25, 9.81, 0.03, 0.1, 0.13, 0.8, 100.0)
    ctl = pp.Optimization(op2)
    g = load_golden("closed_loop_test200.npz")
    for t in range(30):
        o = ctl.step(sim.get_state(), dyn, 0.0)
        assert abs(o.u[0] - g["controls"][0][t]) <= 1e-5 * max(1.0, abs(g["controls"][0][t]))
        sim.step(dyn, op2.control_dt, o.u[0], pp.Vector2(0, 0), pp.Vector2(0, 0))
    with pytest.raises(Exception):
        bad = pp.OptimizationParams()
        bad.state_spacing = 7
        pp.Optimization(bad)


def test_pipelined_steps_equal_synchronous_steps():
    """cpmpc_step_async / cpmpc_wait (two steps in flight, D2H of step k overlapping the kernel of step k + 1): the
    same bits as cpmpc_step, for a chain of warm-started steps, for cold steps via CPMPC_STEP_RESET, and when
    synchronous calls are mixed in (they retire the steps in flight first)."""
    import torch
    from cart_pole_mpc_b200 import synthetic, _capi
    B, W, T = 512, 40, 5
    st, dy, sp = synthetic.swingup_batch(B, seed=6)
    states = [st + 0.01 * t for t in range(T)]
    ref = _solver(B)
    want = [ref.step(states[t], dy, sp) for t in range(T)]           # warm-started chain, one step at a time
    cold = _solver(B).step(states[2], dy, sp)

    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()
    s = _solver(B)
    ins = [pin(x) for x in states]
    dyp, spp = pin(dy), pin(sp)
    u = [pin(np.zeros((B, W))) for _ in range(T)]
    pred = [pin(np.zeros((B, W, 4))) for _ in range(T)]
    stats = [pin(np.zeros(B, dtype=_capi.STATS_DTYPE).view(np.uint8)) for _ in range(T)]
    for t in range(T):
        s.step_async(ins[t], dyp, spp, u[t], pred_out=pred[t], stats_out=stats[t])
        if t > 0:
            s.wait()
    s.wait()
    s.wait()  # nothing in flight: no-op
    for t in range(T):
        # equal_nan: an instance that ends a step with QP_INDEFINITE warm-starts the next one from clamped controls, its
        # rollout overflows and the step is rejected with non-finite predicted states — in both chains alike
        assert np.array_equal(u[t], want[t].u, equal_nan=True), t
        assert np.array_equal(pred[t], want[t].predicted_states, equal_nan=True), t
        assert stats[t].tobytes() == want[t].stats.tobytes(), t
    # cold steps through the flag, without waiting in between (the third call retires the first by itself)
    for t in range(3):
        s.step_async(ins[2], dyp, spp, u[t], reset=True)
    sol, has = s.get_previous_solution()       # a synchronous call: retires everything in flight first
    assert has.all()
    for t in range(3):
        assert np.array_equal(u[t], cold.u), t
    # and a synchronous step after pipelined ones continues the same warm-start chain
    nxt = s.step(states[3], dy, sp)
    ref2 = _solver(B)
    ref2.step(states[2], dy, sp)
    assert np.array_equal(nxt.u, ref2.step(states[3], dy, sp).u)


@pytest.mark.parametrize("over", [dict(), dict(window_length=160), dict(state_spacing=5, max_iterations=10)],
                         ids=["defaults", "w160", "s5"])
def test_split_launch_equals_fused_launch(over):
    """CPMPC_FLAG_SPLIT_LAUNCH (guess / solve / predicted-state kernels, the serial rollouts of
    optimization.cc:333-371 one thread per instance) against CPMPC_FLAG_FUSED_LAUNCH: the same bits for a cold step, a
    warm-started step with per-instance weights, a step without predicted states, and the closed loop (which asks for
    no u_out: the library's own control scratch), on a ragged batch."""
    from cart_pole_mpc_b200 import BatchedOptimization, synthetic
    B = 203
    st, dy, sp, wt = synthetic.sweep_batch(B)
    outs = []
    for split in (False, True):
        s = BatchedOptimization(B, device=0, split_launch=split, **over)
        cold = s.step(st, dy, sp)
        warm = s.step(st + 0.01, dy, sp, weights=wt)
        nopred = s.step(st + 0.02, dy, sp, want_predicted=False)
        prev, has = s.get_previous_solution()
        s.reset()
        fin, ctl, counts = s.closed_loop(4, st, dy, sp)
        outs.append((cold, warm, nopred, prev, has, fin, ctl, counts))
        s.close()
    a, b = outs
    for k in range(3):
        assert np.array_equal(a[k].u, b[k].u, equal_nan=True), k
        assert a[k].stats.tobytes() == b[k].stats.tobytes(), k
        if k < 2:
            assert np.array_equal(a[k].predicted_states, b[k].predicted_states, equal_nan=True), k
    for k in range(3, 8):
        assert np.array_equal(a[k], b[k], equal_nan=True), k


@pytest.mark.parametrize("over", [dict(), dict(state_spacing=5, max_iterations=10), dict(u_cost_weight=0.0)],
                         ids=["defaults", "s5", "lm_heavy"])
def test_compact_shared_memory_layout_equals_roomy_layout(over, monkeypatch):
    """Batches beyond one wave use a compact shared-memory layout (csrc/cpmpc_capi.cu: the condensed QP's Z rows in the
    L2-resident global scratch, the Riccati forward sweep staged a few knots at a time) so that eight warps fit on an
    SM.  Forced here on a small batch through the experiment switches: the same bits as the roomy layout, including
    instances that take Levenberg-Marquardt iterations through the Riccati recursion (u_cost_weight = 0)."""
    from cart_pole_mpc_b200 import BatchedOptimization, synthetic
    B = 333
    st, dy, sp = synthetic.swingup_batch(B, seed=13)
    outs = []
    for env in (dict(CPMPC_ZGLOBAL="0"), dict(CPMPC_ZGLOBAL="1"), dict(CPMPC_ZGLOBAL="1", CPMPC_STAGE_KNOTS="3")):
        monkeypatch.delenv("CPMPC_STAGE_KNOTS", raising=False)
        for key, val in env.items():
            monkeypatch.setenv(key, val)
        s = BatchedOptimization(B, device=0, **over)
        cold = s.step(st, dy, sp)
        warm = s.step(st + 0.01, dy, sp)
        outs.append((cold, warm))
        s.close()
    for other in outs[1:]:
        for k in range(2):
            assert np.array_equal(outs[0][k].u, other[k].u, equal_nan=True), k
            assert np.array_equal(outs[0][k].predicted_states, other[k].predicted_states, equal_nan=True), k
            assert outs[0][k].stats.tobytes() == other[k].stats.tobytes(), k
    if over.get("u_cost_weight") == 0.0:
        assert (outs[0][0].stats["lambda"] > 0).any() or (outs[0][0].stats["n_failed_line_searches"] > 0).any()


def test_pipelined_steps_with_split_launch_equal_fused_synchronous_steps():
    """cpmpc_step_async / cpmpc_wait on a handle that uses the split launch (the three kernels of a step and the shared
    guess buffer are chained behind the previous step's kernels): the same bits as synchronous fused steps."""
    import torch
    from cart_pole_mpc_b200 import BatchedOptimization, synthetic, _capi
    B, W, T = 300, 40, 4
    st, dy, sp = synthetic.swingup_batch(B, seed=21)
    states = [st + 0.01 * t for t in range(T)]
    ref = BatchedOptimization(B, split_launch=False)
    want = [ref.step(states[t], dy, sp) for t in range(T)]
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()
    s = BatchedOptimization(B, split_launch=True)
    ins = [pin(x) for x in states]
    dyp, spp = pin(dy), pin(sp)
    u = [pin(np.zeros((B, W))) for _ in range(T)]
    pred = [pin(np.zeros((B, W, 4))) for _ in range(T)]
    stats = [pin(np.zeros(B, dtype=_capi.STATS_DTYPE).view(np.uint8)) for _ in range(T)]
    for t in range(T):
        s.step_async(ins[t], dyp, spp, u[t], pred_out=pred[t], stats_out=stats[t])
    s.wait()
    s.wait()
    for t in range(T):
        assert np.array_equal(u[t], want[t].u, equal_nan=True), t
        assert np.array_equal(pred[t], want[t].predicted_states, equal_nan=True), t
        assert stats[t].tobytes() == want[t].stats.tobytes(), t


def test_large_batch_launch_shape_properties_and_parity(oracle):
    """65536 cold swing-ups (BASELINE.json's supplementary size): the library picks the split launch and the compact
    shared-memory layout by itself.  Size-independent properties, determinism, batch-composition independence against
    the single-wave mapping (instance i of the big batch == instance i solved in a batch of a few hundred, bit for bit),
    and parity of a sample with the compiled reference at the headline tolerance."""
    from cart_pole_mpc_b200 import launch_count, synthetic
    B = 65536
    st, dy, sp = synthetic.swingup_batch(B)
    s = _solver(B)
    n0 = launch_count()
    out = s.step(st, dy, sp)
    assert launch_count() - n0 == 3, "guess, solve and predicted-state kernels"
    assert np.all(np.isfinite(out.u)) and np.all(np.abs(out.u) <= 300.0)
    ls, it = out.stats["n_line_search_steps"], out.stats["n_iterations"]
    assert np.all((it >= 1) & (it <= 8)) and np.all(ls >= it) and np.all(ls <= 6 * it)
    th = out.predicted_states[:, :, 1]
    fin = np.isfinite(th)
    assert fin.mean() > 0.999 and np.all((th[fin] > -math.pi - 1e-12) & (th[fin] <= math.pi + 1e-12))
    sol, has = s.get_previous_solution()
    assert has.all() and np.array_equal(sol[:, -40:], out.u)
    s.reset()
    out2 = s.step(st, dy, sp)
    assert np.array_equal(out.u, out2.u) and out.stats.tobytes() == out2.stats.tobytes()
    s.close()
    idx = np.r_[0:64, 4090:4100, 30000:30128, 65472:65536]
    small = _solver(len(idx)).step(st[idx], dy[idx], sp[idx])
    assert np.array_equal(small.u, out.u[idx])
    assert np.array_equal(small.predicted_states, out.predicted_states[idx], equal_nan=True)
    assert small.stats.tobytes() == out.stats[idx].tobytes()
    if oracle.ref_available():
        with oracle.using_ref():
            ref = oracle.batch_step(oracle.opt_array(), st[idx], dy[idx], sp[idx], want_iter_log=False)
        _check_batch(small, ref.stats, ref.u, min_frac=0.93)


def test_far_out_pole_angles_are_wrapped_on_the_way_in(oracle):
    """A caller may pass any finite pole angle; the kernel's branch-free sincos is valid for |x| < 2^20 only, so states
    are wrapped on the way into every kernel (csrc/dynamics.cuh wrap_far_angle).  The solve from th = 2^21 + x equals
    the solve from the wrapped angle bit for bit, and agrees with the reference (libm sin / cos on the raw angle) to
    the rounding of 2 pi in the reduction."""
    from cart_pole_mpc_b200 import synthetic, sim_step
    B = 64
    st, dy, sp = synthetic.swingup_batch(B, seed=5)
    far = st.copy()
    far[:, 1] = 2.0 ** 21 + np.linspace(-3.0, 3.0, B)
    two_pi, pi = 6.283185307179586, 3.141592653589793
    w = np.fmod(far[:, 1], two_pi)
    w = np.where(w < 0.0, w + two_pi, w)
    w = np.where(w > pi, w - two_pi, w)
    wrapped = far.copy()
    wrapped[:, 1] = w
    a = _solver(B).step(far, dy, sp)
    b = _solver(B).step(wrapped, dy, sp)
    assert np.array_equal(a.u, b.u) and np.array_equal(a.predicted_states, b.predicted_states)
    assert a.stats.tobytes() == b.stats.tobytes()
    assert np.array_equal(sim_step(far, dy, a.u[:, 0], 0.01), sim_step(wrapped, dy, a.u[:, 0], 0.01))
    ref = oracle.batch_step(oracle.opt_array(), far, dy, sp, want_iter_log=False)
    rel = _rel_cost(a, ref.stats)
    assert np.median(rel) < 1e-6 and (rel <= 1e-4).mean() >= 0.9, rel


@pytest.mark.parametrize("over", [dict(), dict(u_cost_weight=0.0), dict(window_length=160)],
                         ids=["defaults", "lm_heavy", "w160"])
def test_phase_clocks_are_opt_in_and_change_nothing(over):
    """cpmpc_enable_phase_clocks: the default (lean) kernels record only a warp's total cycles and its last batch; with
    the clocks enabled the general kernel fills the per-phase entries, which then add up to the total, and the results
    are the same bits - also where the Riccati recursion does the work (damped iterations, W = 160), whose sums the
    compiler contracts as it sees fit in each instantiation."""
    from cart_pole_mpc_b200 import synthetic
    B = 512
    st, dy, sp = synthetic.swingup_batch(B, seed=13)
    s = _solver(B, **over)
    a = s.step(st, dy, sp)
    rec = s.warp_cycles(with_instances=True).astype(np.int64)
    rec = rec[rec[:, 0] > 0]
    assert rec.shape[0] == B // (8 if s.W == 40 else 4) and np.all(rec[:, 2:] == 0)
    s.reset()
    s.enable_phase_clocks(True)
    b = s.step(st, dy, sp)
    rec = s.warp_cycles(with_instances=True).astype(np.int64)
    rec = rec[rec[:, 0] > 0]
    assert np.all(rec[:, 2:].sum(axis=1) > 0.95 * rec[:, 0]) and np.all(rec[:, 2:].sum(axis=1) <= rec[:, 0])
    assert np.array_equal(a.u, b.u, equal_nan=True) and np.array_equal(a.predicted_states, b.predicted_states, equal_nan=True)
    assert a.stats.tobytes() == b.stats.tobytes()
