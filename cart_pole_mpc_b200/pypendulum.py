"""Drop-in mirror of the reference's nanobind module `pypendulum` (wrapper/wrapper.cc:40-102) on top of the CUDA
library: same class names, constructor signatures, attribute names and method names, so model/scratch.py runs with
only its import line changed.  Every object is a batch of one on the GPU; there is no CPU path.
"""
from __future__ import annotations

from typing import List

import numpy as np

from . import batched as _b


class SingleCartPoleParams:
    """wrapper.cc:41-54 / optimization/structs.hpp:8-41"""
    __slots__ = ("m_b", "m_1", "l_1", "g", "mu_b", "v_mu_b", "c_d_1", "x_s", "k_s")

    def __init__(self, m_b=0.0, m_1=0.0, l_1=0.0, g=0.0, mu_b=0.0, v_mu_b=0.0, c_d_1=0.0, x_s=0.0, k_s=0.0):
        self.m_b, self.m_1, self.l_1, self.g, self.mu_b = float(m_b), float(m_1), float(l_1), float(g), float(mu_b)
        self.v_mu_b, self.c_d_1, self.x_s, self.k_s = float(v_mu_b), float(c_d_1), float(x_s), float(k_s)

    def _array(self) -> np.ndarray:
        return np.array([[getattr(self, k) for k in self.__slots__]], dtype=np.float64)


class SingleCartPoleState:
    """wrapper.cc:74-79 / optimization/structs.hpp:44-64"""
    __slots__ = ("b_x", "th_1", "b_x_dot", "th_1_dot")

    def __init__(self, b_x: float, th_1: float, b_x_dot: float, th_1_dot: float):
        self.b_x, self.th_1, self.b_x_dot, self.th_1_dot = float(b_x), float(th_1), float(b_x_dot), float(th_1_dot)

    def _array(self) -> np.ndarray:
        return np.array([[self.b_x, self.th_1, self.b_x_dot, self.th_1_dot]], dtype=np.float64)

    def __repr__(self):
        return f"SingleCartPoleState({self.b_x}, {self.th_1}, {self.b_x_dot}, {self.th_1_dot})"


class Vector2:
    """wrapper.cc:93 / optimization/structs.hpp:67-70"""
    __slots__ = ("x", "y")

    def __init__(self, x: float, y: float):
        self.x, self.y = float(x), float(y)


class OptimizationParams:
    """wrapper.cc:57-72 / optimization/optimization.hpp:12-53 (same defaults)."""
    _FIELDS = _b.PARAM_FIELDS
    _INT_FIELDS = ("window_length", "state_spacing", "max_iterations")

    def __init__(self):
        d = _b.make_params()
        for k in self._FIELDS:
            setattr(self, k, getattr(d, k))

    def _struct(self):
        kw = {k: (int(getattr(self, k)) if k in self._INT_FIELDS else float(getattr(self, k))) for k in self._FIELDS}
        return _b.make_params(**kw)


class OptimizationOutputs:
    """wrapper.cc:81-85 / optimization/optimization.hpp:55-70"""

    def __init__(self, initial_state, previous_solution, u, predicted_states, stats, iter_log):
        self.initial_state = initial_state
        self.previous_solution: List[float] = previous_solution
        self.u: List[float] = u
        self.predicted_states: List[SingleCartPoleState] = predicted_states
        self.stats = stats
        self._iter_log = iter_log

    @staticmethod
    def _g(x) -> str:
        """fmt's "{}" for a double: shortest round-trip digits, no trailing ".0"."""
        t = repr(float(x))
        return t[:-2] if t.endswith(".0") else t

    def solver_summary(self) -> str:
        """Same line layout as mini_opt::NLSSolverOutputs::ToString (mini_opt/source/structs.cc:159-268)."""
        lines = []
        n_it = int(self.stats["n_iterations"])
        for i in range(n_it):
            r = self._iter_log[i]
            state = "NOMINAL" if int(r[0]) == 0 else "ATTEMPTING_RESTORE_LM"
            pen = r[6]
            lines.append(f"Iteration # {i}, state = {state}, lambda = {self._g(r[1])}")
            lines.append(f"  f(0): {r[2]:.16e}, c(0): {r[3]:.16e}, total: {r[2] + pen * r[3]:.16e}")
            lines.append(f"  penalty = {pen:.16f}")
            lines.append("  QP: SUCCESS")
            lines.append(f"  df/dalpha = {self._g(r[4])}, dc/dalpha = {self._g(r[5])}")
            lines.append(f"  Search result: {_b.STEP_RESULT_NAMES.get(int(r[7]), '?')}")
            for s in range(min(int(r[8]), 6)):
                a, f, c = r[9 + 3 * s], r[10 + 3 * s], r[11 + 3 * s]
                lines.append(f"  f({s}): {f:.16e}, c({s}): {c:.16e}, total: {f + pen * c:.16e}, alpha = {a:.10f}")
        lines.append(f"  termination = {_b.TERMINATION_NAMES.get(int(self.stats['termination_state']), '?')}")
        return "\n".join(lines) + "\n"


    # ---- the JSON log format of the reference's web front end --------------------------------------------------------
    def to_dict(self) -> dict:
        """Same object layout as nlohmann's `json(OptimizationOutputs)` in the reference: optimization/wasm.cc:19-27,46-65
        (OptimizationOutputs, SingleCartPoleState) and dependencies/mini_opt/source/serialization.cc:32-134
        (NLSSolverOutputs, NLSIteration, Errors, DirectionalDerivatives, LineSearchStep; enums by name).  The
        equality-only problem always uses the null-space QP, so `qp_outputs` is the QPNullSpaceTerminationState name
        and `qp_eigenvalues` is null (log_qp_eigenvalues is never enabled by Optimization)."""
        def state(s):
            return {"b_x": s.b_x, "th_1": s.th_1, "th_1_dot": s.th_1_dot, "b_x_dot": s.b_x_dot}

        iterations = []
        for i in range(int(self.stats["n_iterations"])):
            r = self._iter_log[i]
            iterations.append({
                "iteration": i,
                "optimizer_state": "NOMINAL" if int(r[0]) == 0 else "ATTEMPTING_RESTORE_LM",
                "lambda": float(r[1]),
                "errors_initial": {"f": float(r[2]), "equality": float(r[3])},
                "qp_outputs": "SUCCESS",
                "qp_eigenvalues": None,
                "directional_derivatives": {"d_f": float(r[4]), "d_equality": float(r[5])},
                "penalty": float(r[6]),
                "step_result": _b.STEP_RESULT_NAMES.get(int(r[7]), "?"),
                "line_search_steps": [
                    {"alpha": float(r[9 + 3 * k]), "errors": {"f": float(r[10 + 3 * k]), "equality": float(r[11 + 3 * k])}}
                    for k in range(min(int(r[8]), 6))],
            })
        return {
            "initial_state": state(self.initial_state),
            "previous_solution": [float(v) for v in self.previous_solution],
            "solver_outputs": {"termination_state": _b.TERMINATION_NAMES.get(int(self.stats["termination_state"]), "?"),
                               "iterations": iterations},
            "u": [float(v) for v in self.u],
            "predicted_states": [state(s) for s in self.predicted_states],
        }

    def to_json(self, **dumps_kwargs) -> str:
        import json
        return json.dumps(self.to_dict(), **dumps_kwargs)


def outputs_from_json(text) -> "OptimizationOutputs":
    """Inverse of OptimizationOutputs.to_json (optimization/wasm.cc:48-56 `from_json`): rebuilds an outputs object from
    a log written by either implementation, e.g. to replay it through solver_summary()."""
    import json
    j = json.loads(text) if isinstance(text, (str, bytes)) else text
    term_ids = {v: k for k, v in _b.TERMINATION_NAMES.items()}
    step_ids = {v: k for k, v in _b.STEP_RESULT_NAMES.items()}
    its = j["solver_outputs"]["iterations"]
    log = np.zeros((max(len(its), 1), 27))
    n_ls = n_failed = 0
    for i, it in enumerate(its):
        steps = it["line_search_steps"]
        log[i, :9] = [0.0 if it["optimizer_state"] == "NOMINAL" else 1.0, it["lambda"], it["errors_initial"]["f"],
                      it["errors_initial"]["equality"], it["directional_derivatives"]["d_f"],
                      it["directional_derivatives"]["d_equality"], it["penalty"], step_ids[it["step_result"]],
                      len(steps)]
        for k, st in enumerate(steps[:6]):
            log[i, 9 + 3 * k:12 + 3 * k] = [st["alpha"], st["errors"]["f"], st["errors"]["equality"]]
        n_ls += len(steps)
        n_failed += it["step_result"] not in ("SUCCESS", "FIRST_ORDER_SATISFIED")
    stats = {"termination_state": term_ids[j["solver_outputs"]["termination_state"]], "n_iterations": len(its),
             "n_line_search_steps": n_ls, "n_failed_line_searches": n_failed}
    mk = lambda d: SingleCartPoleState(d["b_x"], d["th_1"], d["b_x_dot"], d["th_1_dot"])
    return OptimizationOutputs(mk(j["initial_state"]), list(j["previous_solution"]), list(j["u"]),
                               [mk(d) for d in j["predicted_states"]], stats, log)


class Optimization:
    """wrapper.cc:87-90 / optimization/optimization.hpp:73-108 — a GPU batch of one."""

    def __init__(self, params: OptimizationParams):
        self._params = params
        self._impl = _b.BatchedOptimization(1, params=params._struct())
        self._impl.enable_iteration_log(True)

    def step(self, current_state: SingleCartPoleState, dynamics_params: SingleCartPoleParams,
             b_x_set_point: float) -> OptimizationOutputs:
        prev, has = self._impl.get_previous_solution()
        previous_solution = prev[0].tolist() if has[0] else []
        out = self._impl.step(current_state._array(), dynamics_params._array(), np.array([float(b_x_set_point)]))
        if int(out.stats[0]["termination_state"]) < 0:
            # the reference throws inside Solve() (nonlinear.cc:356-357; nanobind turns it into a RuntimeError); the
            # warm start is left as it was, exactly as there (optimization.cc:84-85 is never reached)
            raise RuntimeError("non-finite cost at the linearization point")
        log = self._impl.get_iteration_log()[0]
        predicted = [SingleCartPoleState(*row) for row in out.predicted_states[0]]
        return OptimizationOutputs(current_state, previous_solution, out.u[0].tolist(), predicted, out.stats[0], log)

    def reset(self) -> None:
        self._impl.reset()

    def set_previous_solution(self, guess) -> None:
        self._impl.set_previous_solution(np.asarray(guess, dtype=np.float64).reshape(1, -1))


class Simulator:
    """wrapper.cc:95-98 / optimization/simulator.hpp:10-29"""

    def __init__(self):
        self._state = np.array([[0.0, -np.pi / 2, 0.0, 0.0]], dtype=np.float64)

    def step(self, params: SingleCartPoleParams, dt: float, u: float, f_base: Vector2, f_mass: Vector2) -> None:
        self._state = _b.sim_step(self._state, params._array(), np.array([float(u)]), float(dt),
                                  np.array([[f_base.x, f_base.y]]), np.array([[f_mass.x, f_mass.y]]))

    def get_state(self) -> SingleCartPoleState:
        return SingleCartPoleState(*self._state[0])

    def set_state(self, state: SingleCartPoleState) -> None:
        self._state = state._array()
