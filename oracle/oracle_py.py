"""ORACLE — TEST INFRASTRUCTURE ONLY.  ctypes binding of oracle/liboracle.so.

Imported only by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
Never imported by the product package (cart_pole_mpc_b200).
"""
from __future__ import annotations

import contextlib
import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")

STATS_STRIDE = 8
ITERLOG_STRIDE = 27
WEIGHTS_STRIDE = 6

OPT_FIELDS = (
    "control_dt", "window_length", "state_spacing", "max_iterations", "relative_exit_tol",
    "absolute_first_derivative_tol", "equality_penalty_initial", "u_guess_sinusoid_amplitude", "u_cost_weight",
    "u_derivative_cost_weight", "b_x_final_cost_weight", "th_final_cost_weight", "b_x_dot_final_cost_weight",
    "th_dot_final_cost_weight",
)
# optimization/optimization.hpp:12-48 defaults
OPT_DEFAULTS = (0.01, 40, 10, 8, 1.0e-5, 1.0e-6, 1.0, 10.0, 0.1, 0.1, 150.0, -1.0, -1.0, -1.0)


def opt_array(**overrides) -> np.ndarray:
    vals = dict(zip(OPT_FIELDS, OPT_DEFAULTS))
    for k, v in overrides.items():
        if k not in vals:
            raise KeyError(k)
        vals[k] = v
    return np.asarray([float(vals[k]) for k in OPT_FIELDS], dtype=np.float64)


def build(force: bool = False) -> str:
    """Compile liboracle.so and oracle_kat with the committed Makefile (g++ only, seconds)."""
    if (force or not os.path.exists(_LIB_PATH) or not os.path.exists(os.path.join(_HERE, "oracle_kat"))
            or not os.path.exists(os.path.join(_HERE, "shim_selftest"))):
        subprocess.run(["make", "-C", _HERE, "-j4"], check=True, capture_output=True)
    return _LIB_PATH


_lib = None
_ref_lib = None
_use_ref = False

# oracle/_ref/libref.so: the reference's own translation units compiled against oracle/eigen_shim (Makefile target
# `ref`).  It can only be BUILT where /root/reference exists (the build container); the built file is git-ignored
# but travels to the GPU box with the repo snapshot.
_REF_LIB_PATH = os.path.join(_HERE, "_ref", "libref.so")
REFERENCE_ROOT = os.environ.get("CPMPC_REFERENCE_ROOT", "/root/reference")


def ref_available() -> bool:
    return os.path.exists(_REF_LIB_PATH)


def build_ref(force: bool = False) -> str | None:
    """Compile oracle/_ref/libref.so from the reference's sources where they lie.  Returns None (and builds nothing)
    when the reference tree is absent, e.g. on the GPU box."""
    if not os.path.isdir(os.path.join(REFERENCE_ROOT, "optimization")):
        return _REF_LIB_PATH if ref_available() else None
    if force or not ref_available():
        subprocess.run(["make", "-C", _HERE, "ref", f"REF={REFERENCE_ROOT}"], check=True, capture_output=True)
    return _REF_LIB_PATH


@contextlib.contextmanager
def using_ref():
    """Route every function of this module that exists in libref.so (dynamics, rk4, sim_step, qp_nullspace,
    Optimization, batch_step, closed_loop) to the compiled reference instead of the restatement."""
    global _use_ref
    if not ref_available():
        raise FileNotFoundError(_REF_LIB_PATH)
    old, _use_ref = _use_ref, True
    try:
        yield
    finally:
        _use_ref = old


def lib() -> C.CDLL:
    global _lib, _ref_lib
    if _use_ref:
        if _ref_lib is None:
            _ref_lib = _bind(C.CDLL(_REF_LIB_PATH), full=False)
        return _ref_lib
    if _lib is None:
        build()
        _lib = _bind(C.CDLL(_LIB_PATH), full=True)
    return _lib


def _bind(L: C.CDLL, full: bool) -> C.CDLL:
    if True:
        dp = C.POINTER(C.c_double)
        L.oracle_mod_pi.restype = C.c_double
        L.oracle_mod_pi.argtypes = [C.c_double]
        L.oracle_dynamics.restype = None
        L.oracle_dynamics.argtypes = [dp, dp, C.c_double, dp, dp, dp, dp, dp]
        L.oracle_rk4.restype = None
        L.oracle_rk4.argtypes = [dp, dp, C.c_double, C.c_double, dp, dp, dp]
        L.oracle_rk4_jac.restype = None
        L.oracle_rk4_jac.argtypes = [dp, dp, C.c_double, C.c_double, dp, dp, dp]
        L.oracle_sim_step.restype = None
        L.oracle_sim_step.argtypes = [dp, dp, C.c_double, C.c_double, dp, dp]
        L.oracle_qp_nullspace.restype = C.c_int
        L.oracle_qp_nullspace.argtypes = [C.c_int, C.c_int, dp, dp, dp, dp, dp]
        if not full:
            L.oracle_ref_info.restype = C.c_char_p
            L.oracle_ref_info.argtypes = []
            # oracle/ref_glue_debug.cc: views inside the reference's solver
            ip = C.POINTER(C.c_int)
            L.oracle_qp_dims.restype = None
            L.oracle_qp_dims.argtypes = [dp, ip, ip]
            L.oracle_linearize_qp.restype = C.c_int
            L.oracle_linearize_qp.argtypes = [dp, dp, dp, C.c_double, dp, C.c_double, dp, dp, dp, dp, dp, dp, dp]
            L.oracle_closed_loop_trace.restype = C.c_int
            L.oracle_closed_loop_trace.argtypes = [dp, C.c_int64, C.c_int64, dp, dp, dp, dp, dp, dp, dp, C.c_int]
            L.oracle_solver_summary.restype = C.c_int64
            L.oracle_solver_summary.argtypes = [C.c_int, C.c_int, dp, C.c_char_p, C.c_int64]
            # oracle/ref_glue_double.cc: the nx = 6 problem solved by the reference's own solver and integrator
            L.oracle_dp_ref_batch_step.restype = C.c_int
            L.oracle_dp_ref_batch_step.argtypes = [dp, C.c_int64, dp, dp, dp, dp, dp, C.POINTER(C.c_uint8), dp, dp, dp,
                                                   dp, dp, C.c_int]
    if full:
        L.oracle_quadratic_approx_minimum.restype = C.c_int
        L.oracle_quadratic_approx_minimum.argtypes = [C.c_double] * 4 + [dp]
        L.oracle_cubic_approx_coeffs.restype = None
        L.oracle_cubic_approx_coeffs.argtypes = [C.c_double] * 6 + [dp]
        L.oracle_cubic_approx_minimum.restype = C.c_int
        L.oracle_cubic_approx_minimum.argtypes = [C.c_double, dp, dp]
    if True:
        L.oracle_opt_create.restype = C.c_void_p
        L.oracle_opt_create.argtypes = [dp]
        L.oracle_opt_destroy.restype = None
        L.oracle_opt_destroy.argtypes = [C.c_void_p]
        L.oracle_opt_reset.restype = None
        L.oracle_opt_reset.argtypes = [C.c_void_p]
        L.oracle_opt_set_previous_solution.restype = None
        L.oracle_opt_set_previous_solution.argtypes = [C.c_void_p, dp, C.c_int64]
        L.oracle_opt_get_previous_solution.restype = C.c_int64
        L.oracle_opt_get_previous_solution.argtypes = [C.c_void_p, dp, C.c_int64]
        L.oracle_opt_step.restype = C.c_int
        L.oracle_opt_step.argtypes = [C.c_void_p, dp, dp, C.c_double, dp, dp, dp, dp]
        L.oracle_batch_step.restype = C.c_int
        L.oracle_batch_step.argtypes = [dp, C.c_int64, dp, dp, dp, dp, dp, C.POINTER(C.c_uint8), dp, dp, dp, dp, dp,
                                        C.c_int]
        L.oracle_closed_loop.restype = C.c_int
        L.oracle_closed_loop.argtypes = [dp, C.c_int64, C.c_int64, dp, dp, dp, dp, dp, C.POINTER(C.c_int64),
                                         C.POINTER(C.c_int64), C.c_int]
    if full:
        L.oracle_dp_dynamics.restype = None
        L.oracle_dp_dynamics.argtypes = [dp, dp, C.c_double, dp, dp, dp]
        L.oracle_dp_rk4.restype = None
        L.oracle_dp_rk4.argtypes = [dp, dp, C.c_double, C.c_double, dp]
        L.oracle_dp_rk4_jac.restype = None
        L.oracle_dp_rk4_jac.argtypes = [dp, dp, C.c_double, C.c_double, dp, dp, dp]
        L.oracle_dp_batch_step.restype = C.c_int
        L.oracle_dp_batch_step.argtypes = L.oracle_batch_step.argtypes
    return L


def _p(a):
    if a is None:
        return None
    assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _f64(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64))


def mod_pi(x: float) -> float:
    return lib().oracle_mod_pi(float(x))


def dynamics(dyn, x, u, f_base=(0.0, 0.0), f_mass=(0.0, 0.0), jacobians=True):
    dyn, x, fb, fm = _f64(dyn), _f64(x), _f64(f_base), _f64(f_mass)
    xdot = np.zeros(4)
    Jx = np.zeros((4, 4)) if jacobians else None
    Ju = np.zeros(4) if jacobians else None
    lib().oracle_dynamics(_p(dyn), _p(x), float(u), _p(fb), _p(fm), _p(xdot), _p(Jx), _p(Ju))
    return (xdot, Jx, Ju) if jacobians else xdot


def rk4(dyn, x, u, h, f_base=(0.0, 0.0), f_mass=(0.0, 0.0)):
    dyn, x, fb, fm = _f64(dyn), _f64(x), _f64(f_base), _f64(f_mass)
    out = np.zeros(4)
    lib().oracle_rk4(_p(dyn), _p(x), float(u), float(h), _p(fb), _p(fm), _p(out))
    return out


def rk4_jac(dyn, x, u, h):
    dyn, x = _f64(dyn), _f64(x)
    out, Dx, Du = np.zeros(4), np.zeros((4, 4)), np.zeros(4)
    lib().oracle_rk4_jac(_p(dyn), _p(x), float(u), float(h), _p(out), _p(Dx), _p(Du))
    return out, Dx, Du


def sim_step(dyn, state, dt, u, f_base=(0.0, 0.0), f_mass=(0.0, 0.0)):
    dyn, fb, fm = _f64(dyn), _f64(f_base), _f64(f_mass)
    s = _f64(state).copy()
    lib().oracle_sim_step(_p(dyn), _p(s), float(dt), float(u), _p(fb), _p(fm))
    return s


def qp_nullspace(G, c, A, b):
    G, c, A, b = _f64(G), _f64(c), _f64(A), _f64(b)
    N, K = G.shape[0], A.shape[0]
    x = np.zeros(N)
    status = lib().oracle_qp_nullspace(N, K, _p(G), _p(c), _p(A), _p(b), _p(x))
    return status, x


@dataclass
class BatchResult:
    u: np.ndarray           # [B, W]
    predicted: np.ndarray   # [B, W, 4]
    solution: np.ndarray    # [B, N]
    stats: np.ndarray       # [B, 8]  term, n_iter, n_ls, f, equality, penalty, last_lambda, n_failed_ls
    iter_log: np.ndarray    # [B, max_iter, 27]
    threads: int

    @property
    def termination(self):
        return self.stats[:, 0].astype(np.int64)

    @property
    def n_iterations(self):
        return self.stats[:, 1].astype(np.int64)

    @property
    def n_line_search(self):
        return self.stats[:, 2].astype(np.int64)

    @property
    def merit(self):
        """f + penalty * equality at the final variables."""
        return self.stats[:, 3] + self.stats[:, 5] * self.stats[:, 4]


def batch_step(opt, states, dyn, set_points, weights=None, prev=None, has_prev=None, n_threads=0,
               want_iter_log=True) -> BatchResult:
    opt = _f64(opt)
    states, dyn, set_points = _f64(states), _f64(dyn), _f64(set_points)
    B = states.shape[0]
    W, S, max_it = int(opt[1]), int(opt[2]), int(opt[3])
    N = (W // S + 1) * 4 + W
    u = np.zeros((B, W))
    pred = np.zeros((B, W, 4))
    sol = np.zeros((B, N))
    stats = np.zeros((B, STATS_STRIDE))
    ilog = np.zeros((B, max_it, ITERLOG_STRIDE)) if want_iter_log else None
    weights = _f64(weights) if weights is not None else None
    prev = _f64(prev) if prev is not None else None
    hp = None
    if has_prev is not None:
        has_prev = np.ascontiguousarray(np.asarray(has_prev, dtype=np.uint8))
        hp = has_prev.ctypes.data_as(C.POINTER(C.c_uint8))
    threads = lib().oracle_batch_step(_p(opt), B, _p(states), _p(dyn), _p(set_points), _p(weights), _p(prev), hp,
                                      _p(u), _p(pred), _p(sol), _p(stats), _p(ilog), int(n_threads))
    return BatchResult(u, pred, sol, stats, ilog, threads)


# ---- double pendulum on a cart (oracle/double_cartpole.hpp) ---------------------------------------------------------
DP_NX = 6
DP_WEIGHTS_STRIDE = 8
DP_OPT_FIELDS = OPT_FIELDS[:10] + ("b_x_final_cost_weight", "th_1_final_cost_weight", "th_2_final_cost_weight",
                                   "b_x_dot_final_cost_weight", "th_1_dot_final_cost_weight",
                                   "th_2_dot_final_cost_weight")
DP_OPT_DEFAULTS = OPT_DEFAULTS[:10] + (150.0, -1.0, -1.0, -1.0, -1.0, -1.0)


def dp_opt_array(**overrides) -> np.ndarray:
    vals = dict(zip(DP_OPT_FIELDS, DP_OPT_DEFAULTS))
    for k, v in overrides.items():
        if k not in vals:
            raise KeyError(k)
        vals[k] = v
    return np.asarray([float(vals[k]) for k in DP_OPT_FIELDS], dtype=np.float64)


def dp_dynamics(dyn, x, u, jacobians=True):
    dyn, x = _f64(dyn), _f64(x)
    xdot = np.zeros(6)
    Jx = np.zeros((6, 6)) if jacobians else None
    Ju = np.zeros(6) if jacobians else None
    lib().oracle_dp_dynamics(_p(dyn), _p(x), float(u), _p(xdot), _p(Jx), _p(Ju))
    return (xdot, Jx, Ju) if jacobians else xdot


def dp_rk4(dyn, x, u, h):
    dyn, x = _f64(dyn), _f64(x)
    out = np.zeros(6)
    lib().oracle_dp_rk4(_p(dyn), _p(x), float(u), float(h), _p(out))
    return out


def dp_rk4_jac(dyn, x, u, h):
    dyn, x = _f64(dyn), _f64(x)
    out, Dx, Du = np.zeros(6), np.zeros((6, 6)), np.zeros(6)
    lib().oracle_dp_rk4_jac(_p(dyn), _p(x), float(u), float(h), _p(out), _p(Dx), _p(Du))
    return out, Dx, Du


def dp_batch_step(opt, states, dyn, set_points, weights=None, prev=None, has_prev=None, n_threads=0,
                  want_iter_log=True) -> BatchResult:
    opt = _f64(opt)
    states, dyn, set_points = _f64(states), _f64(dyn), _f64(set_points)
    B = states.shape[0]
    W, S, max_it = int(opt[1]), int(opt[2]), int(opt[3])
    N = (W // S + 1) * DP_NX + W
    u = np.zeros((B, W))
    pred = np.zeros((B, W, DP_NX))
    sol = np.zeros((B, N))
    stats = np.zeros((B, STATS_STRIDE))
    ilog = np.zeros((B, max_it, ITERLOG_STRIDE)) if want_iter_log else None
    weights = _f64(weights) if weights is not None else None
    prev = _f64(prev) if prev is not None else None
    hp = None
    if has_prev is not None:
        has_prev = np.ascontiguousarray(np.asarray(has_prev, dtype=np.uint8))
        hp = has_prev.ctypes.data_as(C.POINTER(C.c_uint8))
    threads = lib().oracle_dp_batch_step(_p(opt), B, _p(states), _p(dyn), _p(set_points), _p(weights), _p(prev), hp,
                                         _p(u), _p(pred), _p(sol), _p(stats), _p(ilog), int(n_threads))
    return BatchResult(u, pred, sol, stats, ilog, threads)


def dp_ref_batch_step(opt, states, dyn, set_points, weights=None, prev=None, has_prev=None, n_threads=0,
                      want_iter_log=True) -> BatchResult:
    """The double-pendulum controller with the COMPILED REFERENCE's solver (mini_opt SQP + null-space QP + Residual
    scatter) and integrator templates (runge_kutta_4th_order<6>), oracle/ref_glue_double.cc.  Same contract as
    dp_batch_step."""
    opt = _f64(opt)
    states, dyn, set_points = _f64(states), _f64(dyn), _f64(set_points)
    B = states.shape[0]
    W, S, max_it = int(opt[1]), int(opt[2]), int(opt[3])
    N = (W // S + 1) * DP_NX + W
    u, pred, sol = np.zeros((B, W)), np.zeros((B, W, DP_NX)), np.zeros((B, N))
    stats = np.zeros((B, STATS_STRIDE))
    ilog = np.zeros((B, max_it, ITERLOG_STRIDE)) if want_iter_log else None
    weights = _f64(weights) if weights is not None else None
    prev = _f64(prev) if prev is not None else None
    hp = None
    if has_prev is not None:
        has_prev = np.ascontiguousarray(np.asarray(has_prev, dtype=np.uint8))
        hp = has_prev.ctypes.data_as(C.POINTER(C.c_uint8))
    threads = _reflib().oracle_dp_ref_batch_step(_p(opt), B, _p(states), _p(dyn), _p(set_points), _p(weights), _p(prev),
                                                 hp, _p(u), _p(pred), _p(sol), _p(stats), _p(ilog), int(n_threads))
    return BatchResult(u, pred, sol, stats, ilog, threads)


def closed_loop(opt, states, dyn, set_points, n_steps, n_threads=0):
    """Returns (final_states, controls[B,n_steps], last_pred_terminal[B,4], counts[4], term_hist[7], threads)."""
    opt = _f64(opt)
    states = _f64(states).copy()
    dyn, set_points = _f64(dyn), _f64(set_points)
    B = states.shape[0]
    controls = np.zeros((B, n_steps))
    term = np.zeros((B, 4))
    counts = np.zeros(4, dtype=np.int64)
    hist = np.zeros(7, dtype=np.int64)
    threads = lib().oracle_closed_loop(_p(opt), B, int(n_steps), _p(states), _p(dyn), _p(set_points), _p(controls),
                                       _p(term), counts.ctypes.data_as(C.POINTER(C.c_int64)),
                                       hist.ctypes.data_as(C.POINTER(C.c_int64)), int(n_threads))
    return states, controls, term, counts, hist, threads


# ---- views inside the COMPILED REFERENCE (oracle/ref_glue_debug.cc; libref.so only) ----------------------------------
@dataclass
class LinearizedQP:
    status: int           # QPNullSpaceTerminationState (0 = SUCCESS)
    guess: np.ndarray     # [N] variables the linearisation was taken at
    G: np.ndarray         # [N, N] (lower triangle valid)
    c: np.ndarray         # [N]
    A: np.ndarray         # [K, N]
    b: np.ndarray         # [K]
    dx: np.ndarray        # [N]
    f: float
    equality: float


def _reflib():
    with using_ref():
        return lib()


def qp_dims(opt):
    n, k = C.c_int(0), C.c_int(0)
    _reflib().oracle_qp_dims(_p(_f64(opt)), C.byref(n), C.byref(k))
    return n.value, k.value


def linearize_qp(opt, state, dyn, set_point=0.0, prev=None, lam=0.0) -> LinearizedQP:
    """The reference's own QP (LinearizeAndFillQP on BuildProblem's problem) at the initial guess of Step(), and the
    step of its QPNullSpaceSolver.  Compiled reference only."""
    opt, state, dyn = _f64(opt), _f64(state), _f64(dyn)
    N, K = qp_dims(opt)
    prev = _f64(prev) if prev is not None else None
    guess, G, c = np.zeros(N), np.zeros((N, N)), np.zeros(N)
    A, b, dx, err = np.zeros((K, N)), np.zeros(K), np.zeros(N), np.zeros(2)
    rc = _reflib().oracle_linearize_qp(_p(opt), _p(state), _p(dyn), float(set_point), _p(prev), float(lam), _p(guess),
                                        _p(G), _p(c), _p(A), _p(b), _p(dx), _p(err))
    if rc < 0:
        raise RuntimeError("the reference threw while linearising")
    return LinearizedQP(rc, guess, G, c, A, b, dx, float(err[0]), float(err[1]))


@dataclass
class ClosedLoopTrace:
    states: np.ndarray   # [B, T, 4] plant state handed to Step at tick t
    prev: np.ndarray     # [B, T, N] warm start Step used at tick t (zeros at t = 0: cold start)
    u: np.ndarray        # [B, T, W]
    stats: np.ndarray    # [B, T, 8]
    threads: int


def closed_loop_trace(opt, states, dyn, set_points, n_steps, n_threads=0) -> ClosedLoopTrace:
    """optimization_test.cc:38-61 for B plants run by the compiled reference, every tick recorded."""
    opt, states, dyn, set_points = _f64(opt), _f64(states), _f64(dyn), _f64(set_points)
    B, T = states.shape[0], int(n_steps)
    W, S = int(opt[1]), int(opt[2])
    N = (W // S + 1) * 4 + W
    st, pv = np.zeros((B, T, 4)), np.zeros((B, T, N))
    u, stats = np.zeros((B, T, W)), np.zeros((B, T, STATS_STRIDE))
    threads = _reflib().oracle_closed_loop_trace(_p(opt), B, T, _p(states), _p(dyn), _p(set_points), _p(st), _p(pv),
                                                  _p(u), _p(stats), int(n_threads))
    return ClosedLoopTrace(st, pv, u, stats, threads)


def solver_summary(termination_state: int, iter_log) -> str:
    """mini_opt::NLSSolverOutputs::ToString (structs.cc:256-266) of the compiled reference on a log given as the
    [n_iterations, 27] array layout of the iteration logs."""
    log = _f64(iter_log).reshape(-1, ITERLOG_STRIDE)
    n = log.shape[0]
    need = _reflib().oracle_solver_summary(int(termination_state), n, _p(log), None, 0)
    buf = C.create_string_buffer(int(need) + 1)
    _reflib().oracle_solver_summary(int(termination_state), n, _p(log), buf, int(need) + 1)
    return buf.value.decode()


class Optimization:
    """Stateful single controller mirroring pendulum::Optimization (optimization/optimization.hpp:73-108)."""

    def __init__(self, opt):
        self.opt = _f64(opt)
        self._h = lib().oracle_opt_create(_p(self.opt))
        if not self._h:
            raise ValueError("invalid OptimizationParams")
        self.W, self.S, self.max_it = int(self.opt[1]), int(self.opt[2]), int(self.opt[3])
        self.N = (self.W // self.S + 1) * 4 + self.W

    def __del__(self):
        if getattr(self, "_h", None):
            lib().oracle_opt_destroy(self._h)
            self._h = None

    def reset(self):
        lib().oracle_opt_reset(self._h)

    def set_previous_solution(self, sol):
        sol = _f64(sol)
        lib().oracle_opt_set_previous_solution(self._h, _p(sol), sol.size)

    def previous_solution(self):
        out = np.zeros(self.N)
        n = lib().oracle_opt_get_previous_solution(self._h, _p(out), self.N)
        return out[:max(n, 0)]

    def step(self, state, dyn, set_point=0.0):
        state, dyn = _f64(state), _f64(dyn)
        u, pred = np.zeros(self.W), np.zeros((self.W, 4))
        stats, ilog = np.zeros(STATS_STRIDE), np.zeros((self.max_it, ITERLOG_STRIDE))
        rc = lib().oracle_opt_step(self._h, _p(state), _p(dyn), float(set_point), _p(u), _p(pred), _p(stats), _p(ilog))
        if rc != 0:
            raise RuntimeError("oracle solver threw")
        return u, pred, stats, ilog
