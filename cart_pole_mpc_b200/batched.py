"""Batched MPC controller over the C ABI: B independent pendulum::Optimization objects on one GPU.

Host side only (argument checking, buffer ownership); all arithmetic happens in libcpmpc.so's CUDA kernels.
The single-instance, reference-shaped API lives in pypendulum.py.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional

import numpy as np

from . import _capi

# mini_opt::NLSTerminationState by declaration order (mini_opt/structs.hpp:232-248); -1 = rejected instance
TERMINATION_NAMES = {
    0: "MAX_ITERATIONS", 1: "SATISFIED_ABSOLUTE_TOL", 2: "SATISFIED_RELATIVE_TOL", 3: "SATISFIED_FIRST_ORDER_TOL",
    4: "MAX_LAMBDA", 5: "QP_INDEFINITE", 6: "USER_CALLBACK", -1: "REJECTED_NON_FINITE",
}
STEP_RESULT_NAMES = {
    0: "SUCCESS", 1: "MAX_ITERATIONS", 2: "FIRST_ORDER_SATISFIED", 3: "POSITIVE_DERIVATIVE",
    4: "FAILURE_NON_FINITE_COST", 5: "FAILURE_INVALID_ALPHA",
}

PARAM_FIELDS = tuple(name for name, _ in _capi.CpmpcParams._fields_)


def make_params(**overrides) -> _capi.CpmpcParams:
    """OptimizationParams{} defaults (optimization/optimization.hpp:12-48) with keyword overrides."""
    p = _capi.CpmpcParams()
    _capi.lib().cpmpc_default_params(C.byref(p))
    for k, v in overrides.items():
        if k not in PARAM_FIELDS:
            raise KeyError(f"unknown OptimizationParams field {k!r}")
        setattr(p, k, v)
    return p


def _host(a, shape, name):
    arr = np.ascontiguousarray(np.asarray(a, dtype=np.float64))
    if arr.shape != shape:
        raise ValueError(f"{name} must have shape {shape}, got {arr.shape}")
    return arr


@dataclass
class StepOutputs:
    u: np.ndarray                 # [B, W]     OptimizationOutputs::u
    predicted_states: np.ndarray  # [B, W, 4]  OptimizationOutputs::predicted_states (None if skipped)
    stats: np.ndarray             # [B] structured array, fields of cpmpc_stats

    @property
    def merit(self) -> np.ndarray:
        """f + penalty * equality at the final variables (Errors::Total, mini_opt/structs.hpp:177)."""
        return self.stats["f"] + self.stats["penalty"] * self.stats["equality"]


class BatchedOptimization:
    """B cart-pole MPC controllers stepped together on one GPU.

    Mirrors pendulum::Optimization (optimization/optimization.hpp:73-108) per instance: `step` == Step,
    `reset` == Reset, `set_previous_solution` == SetPreviousSolution; the warm start lives on the device.
    """

    def __init__(self, batch: int, params: Optional[_capi.CpmpcParams] = None, device: int = 0,
                 low_latency: bool = False, min_latency: bool = False, split_launch: Optional[bool] = None,
                 **overrides):
        self._h = C.c_void_p()
        self._lib = _capi.lib()
        self.params = params if params is not None else make_params(**overrides)
        if params is not None and overrides:
            raise ValueError("pass either params or keyword overrides")
        flags = (_capi.FLAG_LOW_LATENCY if low_latency else 0) | (_capi.FLAG_MIN_LATENCY if min_latency else 0)
        # launch shape (include/cpmpc.h): None = by batch size; the results are bit-identical either way
        if split_launch is not None:
            flags |= _capi.FLAG_SPLIT_LAUNCH if split_launch else _capi.FLAG_FUSED_LAUNCH
        _capi.check(self._lib.cpmpc_create_ex(C.byref(self.params), int(batch), int(device), flags,
                                              C.byref(self._h)))
        self.batch = int(batch)
        self.device = int(device)
        self.W = int(self.params.window_length)
        self.S = int(self.params.state_spacing)
        self.NS = self.W // self.S + 1
        self.N = int(self._lib.cpmpc_solution_dim(self._h))
        self.max_iterations = int(self.params.max_iterations)

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            self._lib.cpmpc_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- host-buffer path (numpy) -----------------------------------------------------------------------------
    def step(self, states, dyn_params, set_points, weights=None, want_predicted: bool = True) -> StepOutputs:
        B, W = self.batch, self.W
        states = _host(states, (B, 4), "states")
        dyn = _host(dyn_params, (B, 9), "dyn_params")
        sp = _host(set_points, (B,), "set_points")
        wts = _host(weights, (B, _capi.WEIGHTS_STRIDE), "weights") if weights is not None else None
        u = np.empty((B, W), dtype=np.float64)
        pred = np.empty((B, W, 4), dtype=np.float64) if want_predicted else None
        stats = np.empty(B, dtype=_capi.STATS_DTYPE)
        _capi.check(self._lib.cpmpc_step(
            self._h, states.ctypes.data, dyn.ctypes.data, sp.ctypes.data,
            wts.ctypes.data if wts is not None else None, u.ctypes.data,
            pred.ctypes.data if pred is not None else None, stats.ctypes.data))
        return StepOutputs(u, pred, stats)

    # ---- pipelined host-buffer path ---------------------------------------------------------------------------------
    def step_async(self, states, dyn_params, set_points, u_out, weights=None, pred_out=None, stats_out=None,
                   reset: bool = False) -> None:
        """cpmpc_step_async: enqueue one Step over caller-owned (ideally pinned) host arrays and return; at most two
        steps are in flight.  The arrays must stay alive and untouched until the `wait()` that retires the step.
        `reset=True` makes it a cold start (Reset() ordered before the solve)."""
        B, W = self.batch, self.W
        for name, a, shape in (("states", states, (B, 4)), ("dyn_params", dyn_params, (B, 9)),
                               ("set_points", set_points, (B,)), ("u_out", u_out, (B, W))):
            if not (isinstance(a, np.ndarray) and a.dtype == np.float64 and a.flags["C_CONTIGUOUS"] and a.shape == shape):
                raise ValueError(f"{name} must be a C-contiguous float64 array of shape {shape}")
        _capi.check(self._lib.cpmpc_step_async(
            self._h, states.ctypes.data, dyn_params.ctypes.data, set_points.ctypes.data,
            weights.ctypes.data if weights is not None else None, u_out.ctypes.data,
            pred_out.ctypes.data if pred_out is not None else None,
            stats_out.ctypes.data if stats_out is not None else None, 1 if reset else 0))

    def wait(self) -> None:
        """cpmpc_wait: block until the oldest step in flight is complete."""
        _capi.check(self._lib.cpmpc_wait(self._h))

    # ---- device-pointer path (torch tensors on this device, float64, contiguous) ----------------------------------
    def step_device(self, states, dyn_params, set_points, u_out, weights=None, pred_out=None, stats_out=None,
                    stream=None) -> None:
        """Asynchronous launch on `stream` (a torch.cuda.Stream or None for the current stream)."""
        import torch
        for name, t in (("states", states), ("dyn_params", dyn_params), ("set_points", set_points), ("u_out", u_out)):
            if not (t.is_cuda and t.dtype == torch.float64 and t.is_contiguous()):
                raise ValueError(f"{name} must be a contiguous float64 CUDA tensor")
        s = stream if stream is not None else torch.cuda.current_stream(self.device)
        _capi.check(self._lib.cpmpc_step_device(
            self._h, states.data_ptr(), dyn_params.data_ptr(), set_points.data_ptr(),
            weights.data_ptr() if weights is not None else None, u_out.data_ptr(),
            pred_out.data_ptr() if pred_out is not None else None,
            stats_out.data_ptr() if stats_out is not None else None, s.cuda_stream))

    def closed_loop(self, n_steps: int, states, dyn_params, set_points, weights=None, want_controls: bool = True):
        """n_steps x {Step; Simulator::Step(control_dt, u[0])} on the device (optimization_test.cc:38-61).

        Returns (final_states[B,4], controls[B,n_steps] or None, counts[4] = solves, sqp iterations, line-search
        steps, bad terminations)."""
        B = self.batch
        st = _host(states, (B, 4), "states").copy()
        dyn = _host(dyn_params, (B, 9), "dyn_params")
        sp = _host(set_points, (B,), "set_points")
        wts = _host(weights, (B, _capi.WEIGHTS_STRIDE), "weights") if weights is not None else None
        controls = np.empty((B, n_steps), dtype=np.float64) if want_controls else None
        counts = np.zeros(4, dtype=np.int64)
        _capi.check(self._lib.cpmpc_closed_loop(
            self._h, int(n_steps), st.ctypes.data, dyn.ctypes.data, sp.ctypes.data,
            wts.ctypes.data if wts is not None else None,
            controls.ctypes.data if controls is not None else None, counts.ctypes.data))
        return st, controls, counts

    def closed_loop_device(self, n_steps: int, states, dyn_params, set_points, weights=None, controls_out=None,
                           stats_sum=None, stream=None) -> None:
        import torch
        s = stream if stream is not None else torch.cuda.current_stream(self.device)
        _capi.check(self._lib.cpmpc_closed_loop_device(
            self._h, int(n_steps), states.data_ptr(), dyn_params.data_ptr(), set_points.data_ptr(),
            weights.data_ptr() if weights is not None else None,
            controls_out.data_ptr() if controls_out is not None else None,
            stats_sum.data_ptr() if stats_sum is not None else None, s.cuda_stream))

    # ---- warm-start state -----------------------------------------------------------------------------------------
    def reset(self, mask=None) -> None:
        m = np.ascontiguousarray(np.asarray(mask, dtype=np.uint8)) if mask is not None else None
        if m is not None and m.shape != (self.batch,):
            raise ValueError("mask must have shape (batch,)")
        _capi.check(self._lib.cpmpc_reset(self._h, m.ctypes.data if m is not None else None))

    def reset_device(self, stream=None) -> None:
        """Asynchronous Reset() of all instances on `stream` (torch stream or None for the current one)."""
        import torch
        s = stream if stream is not None else torch.cuda.current_stream(self.device)
        _capi.check(self._lib.cpmpc_reset_device(self._h, s.cuda_stream))

    def set_previous_solution(self, solutions, mask=None) -> None:
        sol = _host(solutions, (self.batch, self.N), "solutions")
        m = np.ascontiguousarray(np.asarray(mask, dtype=np.uint8)) if mask is not None else None
        _capi.check(self._lib.cpmpc_set_previous_solution(self._h, sol.ctypes.data,
                                                          m.ctypes.data if m is not None else None))

    def get_previous_solution(self):
        sol = np.empty((self.batch, self.N), dtype=np.float64)
        has = np.empty(self.batch, dtype=np.uint8)
        _capi.check(self._lib.cpmpc_get_previous_solution(self._h, sol.ctypes.data, has.ctypes.data))
        return sol, has.astype(bool)

    PHASE_NAMES = ("setup+initial guess", "linearise", "condensed qp", "riccati qp", "derivatives+penalty",
                   "line search", "update+exit tests", "outputs")

    def enable_phase_clocks(self, enable: bool = True) -> None:
        """cpmpc_enable_phase_clocks: launches made afterwards fill the per-phase entries of `warp_cycles` (1.5 % slower;
        same results)."""
        _capi.check(self._lib.cpmpc_enable_phase_clocks(self._h, 1 if enable else 0))

    def warp_cycles(self, with_instances: bool = False) -> np.ndarray:
        """SM cycles every warp of the last launch spent on its instances (cpmpc_get_warp_cycles); with
        `with_instances` the full [n, 10] records: cycles, first instance of the warp's last batch, then the cycles per
        solver phase (PHASE_NAMES)."""
        n = int(self._lib.cpmpc_get_warp_cycles(self._h, None, 0))
        out = np.zeros((max(n, 0), 10), dtype=np.uint64)
        if n > 0 and self._lib.cpmpc_get_warp_cycles(self._h, out.ctypes.data, n) != n:
            raise _capi.CpmpcError("cpmpc_get_warp_cycles failed")
        return out if with_instances else np.ascontiguousarray(out[:, 0])

    # ---- debug view inside the first SQP iteration ----------------------------------------------------------------
    def debug_first_qp(self, states, dyn_params, set_points, weights=None, lam: float = 0.0, qp_method: int = 0):
        """Blocks of the first linearisation and its QP step (cpmpc_debug_first_qp): a dict of arrays
        Phi [B,nseg,4,4], Gam [B,4,W], e [B,nseg,4], dxs [B,NS,4], du [B,W], qp_ok [B], f0 [B], eq0 [B].
        The warm start is honoured and left untouched."""
        B, W = self.batch, self.W
        nseg = W // self.S
        NS = nseg + 1
        st = _host(states, (B, 4), "states")
        dyn = _host(dyn_params, (B, 9), "dyn_params")
        sp = _host(set_points, (B,), "set_points")
        wts = _host(weights, (B, _capi.WEIGHTS_STRIDE), "weights") if weights is not None else None
        stride = int(self._lib.cpmpc_debug_stride(self._h))
        raw = np.empty((B, stride), dtype=np.float64)
        _capi.check(self._lib.cpmpc_debug_first_qp(
            self._h, st.ctypes.data, dyn.ctypes.data, sp.ctypes.data, wts.ctypes.data if wts is not None else None,
            float(lam), int(qp_method), raw.ctypes.data))
        o = 0
        out = {}
        for name, shape in (("Phi", (nseg, 4, 4)), ("Gam", (4, W)), ("e", (nseg, 4)), ("dxs", (NS, 4)), ("du", (W,))):
            n = int(np.prod(shape))
            out[name] = raw[:, o:o + n].reshape((B,) + shape)
            o += n
        out["qp_ok"] = raw[:, o] != 0.0
        out["f0"], out["eq0"], out["lambda"] = raw[:, o + 1], raw[:, o + 2], raw[:, o + 3]
        return out

    # ---- per-iteration log ----------------------------------------------------------------------------------------
    def enable_iteration_log(self, enable: bool = True) -> None:
        _capi.check(self._lib.cpmpc_enable_iteration_log(self._h, 1 if enable else 0))

    def get_iteration_log(self) -> np.ndarray:
        log = np.empty((self.batch, self.max_iterations, _capi.ITERLOG_STRIDE), dtype=np.float64)
        _capi.check(self._lib.cpmpc_get_iteration_log(self._h, log.ctypes.data))
        return log


def sim_step(states, dyn_params, u, dt: float, f_base=None, f_mass=None, device: int = 0) -> np.ndarray:
    """Batched Simulator::Step (optimization/simulator.cc:11-36); returns the new states [B,4]."""
    st = np.ascontiguousarray(np.asarray(states, dtype=np.float64)).copy()
    B = st.shape[0]
    dyn = _host(dyn_params, (B, 9), "dyn_params")
    uu = _host(u, (B,), "u")
    fb = _host(f_base, (B, 2), "f_base") if f_base is not None else None
    fm = _host(f_mass, (B, 2), "f_mass") if f_mass is not None else None
    _capi.check(_capi.lib().cpmpc_sim_step(int(device), B, st.ctypes.data, dyn.ctypes.data, uu.ctypes.data, float(dt),
                                           fb.ctypes.data if fb is not None else None,
                                           fm.ctypes.data if fm is not None else None))
    return st


def measure_fp64_peak(device: int = 0) -> float:
    """Measured DFMA throughput of the device in TFLOP/s (roofline denominator)."""
    out = C.c_double(0.0)
    _capi.check(_capi.lib().cpmpc_measure_fp64_peak(int(device), C.byref(out)))
    return out.value


def debug_fastmath(x, device: int = 0) -> np.ndarray:
    """{sin, cos, tanh, 1/x, sqrt|x|, 1/sqrt|x|} of every x as the DEVICE evaluates them (csrc/fastmath.cuh): [n, 6]."""
    xx = np.ascontiguousarray(x, dtype=np.float64).ravel()
    out = np.empty((xx.size, 6), dtype=np.float64)
    _capi.check(_capi.lib().cpmpc_debug_fastmath(int(device), xx.size, xx.ctypes.data, out.ctypes.data))
    return out


def launch_count() -> int:
    return int(_capi.lib().cpmpc_launch_count())
