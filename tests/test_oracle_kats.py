"""CPU tests that pin the oracle: the reference's own known-answer tests restated against it (oracle/kat_main.cc)
plus the frozen golden vectors.  No GPU needed."""
import math
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, load_golden

KAT_BIN = os.path.join(ROOT, "oracle", "oracle_kat")

# name -> reference test it restates
KAT_CASES = {
    "null_space_solver_1": "dependencies/mini_opt/test/qp_test.cc:576-630",
    "null_space_solver_2": "dependencies/mini_opt/test/qp_test.cc:633-707",
    "qp_cost_derivative": "dependencies/mini_opt/test/nonlinear_test.cc:109-183",
    "quadratic_approx_minimum": "dependencies/mini_opt/test/nonlinear_test.cc:185-207",
    "cubic_approx_coeffs": "dependencies/mini_opt/test/nonlinear_test.cc:210-250",
    "sphere_nonlinear_equality": "dependencies/mini_opt/test/nonlinear_test.cc:745-825",
    "integration_derivatives": "optimization/integration_test.cc:45-80",
    "integration_derivatives_all_terms": "(extension of integration_test.cc:45-80 with every dynamics term active)",
    "friction_dissipation": "optimization/integration_test.cc:82-102",
    "drag_dissipation": "optimization/integration_test.cc:104-125",
    "external_force_symmetry": "optimization/integration_test.cc:127-175",
    "cartpole_closed_loop": "optimization/optimization_test.cc:12-77",
}


@pytest.mark.parametrize("case", sorted(KAT_CASES))
def test_reference_kat(oracle, case):
    proc = subprocess.run([KAT_BIN, case], capture_output=True, text=True, timeout=120)
    assert proc.returncode == 0, f"{case} ({KAT_CASES[case]}) failed:\n{proc.stdout}\n{proc.stderr}"
    assert f"PASS {case}" in proc.stdout


def test_kat_list_is_complete(oracle):
    proc = subprocess.run([KAT_BIN, "--list"], capture_output=True, text=True, timeout=30)
    assert sorted(proc.stdout.split()) == sorted(KAT_CASES)


def test_null_space_kat_through_c_entry_point(oracle):
    """qp_test.cc:633-707 again, assembled by hand and solved through the C entry point the GPU tests use."""
    G = np.zeros((4, 4))
    c = np.zeros(4)
    for idx, A, b in (((0, 1), [[1.7, -0.2], [2.3, 1.2]], [5.4, -3.4]),
                      ((1, 2), [[-5.0, 3.3], [9.1, 1.9]], [-3.3, 4.4]),
                      ((0, 3), [[0.2, -0.5], [1.1, -3.1]], [0.5, 0.0])):
        A, b = np.array(A), np.array(b)
        H = A.T @ A
        g = A.T @ (-b)
        for a_, ia in enumerate(idx):
            c[ia] += g[a_]
            for b_, ib in enumerate(idx):
                if ib <= ia:
                    G[ia, ib] += H[a_, b_]
    Aeq = np.zeros((2, 4))
    Aeq[:, [1, 3]] = [[1.1, -1.1], [0.3, 0.6]]
    beq = -np.array([1.3, -4.0])
    status, x = oracle.qp_nullspace(G, c, Aeq, beq)
    assert status == 0
    assert abs(x[1] - -3.656565656565657) < 1e-14
    assert abs(x[3] - -4.838383838383838) < 1e-14
    assert abs(x[0] - -0.707724112814252) < 1e-13
    assert abs(x[2] - 0.0247370254266801) < 1e-13


def test_null_space_reports_indefinite(oracle):
    # concave along the null space of the constraint -> NOT_POSITIVE_DEFINITE (qp.cc:713-716)
    G = np.diag([1.0, -1.0])
    status, _ = oracle.qp_nullspace(G, np.zeros(2), np.array([[1.0, 0.0]]), np.zeros(1))
    assert status == 1


def test_mod_pi_range_and_identity(oracle):
    for a in np.linspace(-20, 20, 401):
        m = oracle.mod_pi(a)
        assert -math.pi < m <= math.pi
        assert abs(math.sin(m) - math.sin(a)) < 1e-12 and abs(math.cos(m) - math.cos(a)) < 1e-12
    assert oracle.mod_pi(math.pi) == math.pi


def test_dynamics_jacobian_matches_finite_differences(oracle):
    rng = np.random.default_rng(3)
    dyn = np.array([1.3, 0.2, 0.3, 9.81, 0.05, 0.1, 0.02, 0.8, 100.0])
    for _ in range(20):
        x = rng.uniform([-1.2, -3, -2, -6], [1.2, 3, 2, 6])
        u = rng.uniform(-20, 20)
        _, Jx, Ju = oracle.dynamics(dyn, x, u)
        h = 1e-6
        for j in range(4):
            if j == 0 and min(abs(x[0] - 0.8), abs(x[0] + 0.8)) < 1e-3:
                continue  # kink of the bumper spring
            xp, xm = x.copy(), x.copy()
            xp[j] += h
            xm[j] -= h
            fd = (oracle.dynamics(dyn, xp, u, jacobians=False) - oracle.dynamics(dyn, xm, u, jacobians=False)) / (2 * h)
            assert np.allclose(Jx[:, j], fd, rtol=2e-6, atol=2e-6)
        fd = (oracle.dynamics(dyn, x, u + h, jacobians=False) - oracle.dynamics(dyn, x, u - h, jacobians=False)) / (2 * h)
        assert np.allclose(Ju, fd, rtol=1e-6, atol=1e-7)


def test_simulator_substeps(oracle):
    """Simulator::Step(0.01) == 10 RK4 sub-steps of 1 ms with mod_pi after each (simulator.cc:18-22)."""
    dyn = np.array([1.0, 0.1, 0.25, 9.81, 0.05, 0.1, 0.02, 0.8, 100.0])
    x = np.array([0.1, 3.1, 0.3, 2.0])
    s = oracle.sim_step(dyn, x, 0.01, 4.0)
    y = x.copy()
    dt = 0.01
    while dt > 0.0:
        y = oracle.rk4(dyn, y, 4.0, min(dt, 0.001))
        y[1] = oracle.mod_pi(y[1])
        dt -= 0.001
    assert np.array_equal(s, y)


@pytest.mark.parametrize("name", ["config1_default.npz", "config1_test_variant.npz", "scratch_py.npz",
                                  "config2_sample64.npz", "sweep_sample32.npz"])
def test_oracle_reproduces_golden(oracle, name):
    """The frozen fixtures are what the oracle computes today (guards the oracle against silent edits)."""
    g = load_golden(name)
    w = g["weights"] if g["weights"].shape[0] else None
    res = oracle.batch_step(g["opt"], g["states"], g["dyn"], g["set_points"], weights=w)
    assert np.array_equal(res.stats[:, :3], g["stats"][:, :3])  # termination, iterations, line-search steps
    assert np.allclose(res.u, g["u"], rtol=1e-12, atol=1e-12)
    assert np.allclose(res.stats[:, 3:6], g["stats"][:, 3:6], rtol=1e-12, atol=1e-12)
    assert np.allclose(res.predicted, g["predicted"], rtol=1e-12, atol=1e-12)


def test_golden_config1_sanity():
    """Config 1 from the hanging-down state: 8 SQP iterations, ends at MAX_ITERATIONS (the cold swing-up does not
    converge in 8 iterations), the initial-state constraint holds exactly and the controls respect the clamp."""
    g = load_golden("config1_default.npz")
    assert int(g["stats"][0, 0]) == 0 and int(g["stats"][0, 1]) == 8
    sol = g["solution"][0]
    assert np.allclose(sol[:4], g["states"][0], rtol=0, atol=1e-13)  # dx(0) is QP round-off, not exactly 0
    assert np.array_equal(sol[20:], g["u"][0])
    assert np.all(np.abs(g["u"]) <= 300.0)
    # every merit evaluation is logged: between 1 and 6 per iteration (nonlinear.cc:366)
    n_steps = g["iter_log"][0, :, 8]
    assert np.all((n_steps >= 1) & (n_steps <= 6)) and n_steps.sum() == g["stats"][0, 2]


def test_oracle_closed_loop_golden(oracle):
    g = load_golden("closed_loop_test200.npz")
    states, controls, term, counts, hist, _ = oracle.closed_loop(g["opt"], g["x0"], g["dyn"], np.zeros(1), 200,
                                                                n_threads=1)
    assert np.array_equal(counts, g["counts"]) and np.array_equal(hist, g["term_hist"])
    assert np.allclose(controls, g["controls"], rtol=1e-10, atol=1e-10)
    # the reference's acceptance criteria (optimization_test.cc:63-66)
    assert abs(states[0, 2]) < 1e-4 and abs(states[0, 3]) < 1e-3 and abs(states[0, 1] - math.pi / 2) < 1e-4
    assert counts[3] == 0  # never QP_INDEFINITE / MAX_LAMBDA (optimization_test.cc:44-46)


def test_warm_start_shift_semantics(oracle):
    """Second Step() starts from the previous solution shifted by one knot (optimization.cc:50-57) and u_prev is the
    un-shifted previous u(0) (optimization.cc:289-291): a warm-started solve needs fewer iterations."""
    opt = oracle.opt_array(state_spacing=5, max_iterations=10)
    ctl = oracle.Optimization(opt)
    dyn = np.array([1.0, 0.1, 0.25, 9.81, 0.03, 0.1, 0.13, 0.8, 100.0])
    x0 = np.array([0.0, -math.pi / 2, 0.0, 0.0])
    assert ctl.previous_solution().size == 0
    u1, _, s1, _ = ctl.step(x0, dyn)
    prev = ctl.previous_solution()
    assert prev.size == ctl.N and np.array_equal(prev[-ctl.W:], u1)
    x1 = oracle.sim_step(dyn, x0, 0.01, u1[0])
    u2, _, s2, log2 = ctl.step(x1, dyn)
    assert s2[1] <= s1[1]
    ctl.reset()
    assert ctl.previous_solution().size == 0
