"""CPU tests of the views INSIDE the compiled reference (oracle/ref_glue_debug.cc): the QP of one linearisation, the
per-tick closed-loop trace used for teacher-forcing, and mini_opt's own NLSSolverOutputs::ToString against the two
host-side formatters of the product (no GPU involved: the formatters are pure host code fed the reference's log)."""
import math
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT


@pytest.fixture(scope="module")
def ref(oracle):
    if oracle.build_ref() is None or not oracle.ref_available():
        pytest.skip("oracle/_ref/libref.so cannot be built here (no /root/reference)")
    return oracle


def test_linearize_qp_is_the_qp_the_solver_steps_on(ref):
    """KKT conditions of the returned QP at the returned step; block structure of A_eq (optimization.cc:99-160,
    194-267); and the step reproduces the first logged iteration of Step() (d_f = c . dx, nonlinear.cc:457-470)."""
    from cart_pole_mpc_b200 import synthetic
    st, dy, sp = synthetic.swingup_batch(6, seed=3)
    opt = ref.opt_array()
    N, K = ref.qp_dims(opt)
    assert (N, K) == (60, 23)
    with ref.using_ref():
        full = ref.batch_step(opt, st, dy, sp)
    for i in range(6):
        q = ref.linearize_qp(opt, st[i], dy[i], sp[i])
        assert q.status == 0
        assert np.abs(q.A @ q.dx + q.b).max() < 1e-10
        G = np.tril(q.G) + np.tril(q.G, -1).T
        grad = G @ q.dx + q.c
        lam = np.linalg.lstsq(q.A.T, grad, rcond=None)[0]
        assert np.abs(grad - q.A.T @ lam).max() < 1e-8 * max(1.0, np.abs(grad).max())
        # rows 16..19: initial-state equalities with unit Jacobian; -I block of every shooting row
        assert np.array_equal(q.A[16:20, 0:4], np.eye(4))
        for s in range(4):
            assert np.array_equal(q.A[4 * s:4 * s + 4, 4 * (s + 1):4 * (s + 2)], -np.eye(4))
        assert q.f == full.iter_log[i, 0, 2] and q.equality == full.iter_log[i, 0, 3]
        assert math.isclose(float(q.c @ q.dx), full.iter_log[i, 0, 4], rel_tol=1e-12, abs_tol=1e-12)
    # damping adds lambda to every diagonal entry (nonlinear.cc:187-189)
    q0 = ref.linearize_qp(opt, st[0], dy[0], sp[0])
    q1 = ref.linearize_qp(opt, st[0], dy[0], sp[0], lam=0.25)
    assert np.allclose(np.diag(q1.G) - np.diag(q0.G), 0.25, rtol=0, atol=1e-12)


def test_closed_loop_trace_is_the_closed_loop(ref):
    """The trace re-runs optimization_test.cc:38-61: its first controls equal oracle_closed_loop's, tick 0 is cold,
    and tick t's warm start is tick t-1's solution (whose control part is tick t-1's u)."""
    from cart_pole_mpc_b200 import synthetic
    B, T = 3, 6
    st, dy, sp = synthetic.swingup_batch(B, seed=5)
    opt = ref.opt_array(state_spacing=5, max_iterations=10)
    tr = ref.closed_loop_trace(opt, st, dy, sp, T)
    with ref.using_ref():
        _, ctl, _, counts, _, _ = ref.closed_loop(opt, st, dy, sp, T)
    assert np.array_equal(tr.u[:, :, 0], ctl)
    assert np.array_equal(tr.states[:, 0], st)
    assert not tr.prev[:, 0].any()
    assert np.array_equal(tr.prev[:, 1:, -40:], tr.u[:, :-1])
    assert int(tr.stats[:, :, 1].sum()) == int(counts[1])
    # every tick can be replayed as an independent warm-started solve: same result bit for bit
    with ref.using_ref():
        again = ref.batch_step(opt, tr.states[:, 3], dy, sp, prev=tr.prev[:, 3], has_prev=np.ones(B, np.uint8),
                               want_iter_log=False)
    assert np.array_equal(again.u, tr.u[:, 3])


def _facade_summary_exe(tmp_path):
    src = tmp_path / "summary_main.cc"
    src.write_text(r'''
#include <cstdio>
#include <vector>
#include "cart_pole_mpc_b200/host/optimization.hpp"
// reads: termination n_iter then n_iter*27 doubles (hex floats); prints NLSSolverOutputs::ToString of the facade
int main() {
  int term = 0, n = 0;
  if (std::scanf("%d %d", &term, &n) != 2) return 2;
  mini_opt::NLSSolverOutputs so;
  so.termination_state = static_cast<mini_opt::NLSTerminationState>(term);
  for (int i = 0; i < n; ++i) {
    double r[27];
    for (double& v : r) if (std::scanf("%la", &v) != 1) return 3;
    mini_opt::NLSIteration it;
    it.iteration = i;
    it.optimizer_state = static_cast<mini_opt::OptimizerState>(static_cast<int>(r[0]));
    it.lambda = r[1];
    it.errors_initial = {r[2], r[3]};
    it.directional_derivatives = {r[4], r[5]};
    it.penalty = r[6];
    it.step_result = static_cast<mini_opt::StepSizeSelectionResult>(static_cast<int>(r[7]));
    for (int s = 0; s < static_cast<int>(r[8]) && s < 6; ++s)
      it.line_search_steps.push_back({r[9 + 3 * s], {r[10 + 3 * s], r[11 + 3 * s]}});
    so.iterations.push_back(it);
  }
  std::fputs(so.ToString().c_str(), stdout);
  return so.NumQPIterations() == so.iterations.size() ? 0 : 4;
}
''')
    exe = tmp_path / "summary_main"
    env = dict(os.environ)
    env.pop("CXX", None)
    subprocess.run(["g++", "-std=c++17", "-O1", "-I", ROOT, "-o", str(exe), str(src)], check=True, env=env)
    return exe


def test_solver_summary_is_byte_equal_to_the_references_tostring(ref, tmp_path):
    """mini_opt::NLSSolverOutputs::ToString (structs.cc:159-268, compiled into libref) vs the Python mirror's
    solver_summary() and the C++ facade's ToString(), all three fed the SAME solver logs: cold swing-ups, the
    optimization_test.cc variant (early exits), and a u_cost_weight = 0 batch that enters Levenberg-Marquardt mode
    (state = ATTEMPTING_RESTORE_LM, non-trivial lambda values, failed searches)."""
    from cart_pole_mpc_b200 import pypendulum as pp, synthetic
    exe = _facade_summary_exe(tmp_path)
    st, dy, sp = synthetic.swingup_batch(48, seed=9)
    seen_states, seen_results = set(), set()
    for over in (dict(), dict(state_spacing=5, max_iterations=10), dict(u_cost_weight=0.0, max_iterations=12)):
        opt = ref.opt_array(**over)
        with ref.using_ref():
            r = ref.batch_step(opt, st, dy, sp)
        for i in range(st.shape[0]):
            if r.stats[i, 0] < 0:
                continue
            n = int(r.stats[i, 1])
            log = r.iter_log[i, :n]
            want = ref.solver_summary(int(r.stats[i, 0]), log)
            stats = {"termination_state": int(r.stats[i, 0]), "n_iterations": n}
            got_py = pp.OptimizationOutputs(None, [], [], [], stats, log).solver_summary()
            assert got_py == want, f"python mirror differs for instance {i} of {over}"
            feed = f"{int(r.stats[i, 0])} {n}\n" + "\n".join(" ".join(float(v).hex() for v in row) for row in log) + "\n"
            proc = subprocess.run([str(exe)], input=feed, capture_output=True, text=True)
            assert proc.returncode == 0
            assert proc.stdout == want, f"C++ facade differs for instance {i} of {over}"
            seen_states.update(int(v) for v in log[:, 0])
            seen_results.update(int(v) for v in log[:, 7])
    assert seen_states == {0, 1}, "the LM state must be exercised"
    assert {0, 1} <= seen_results or {0, 3} <= seen_results or len(seen_results) >= 2
