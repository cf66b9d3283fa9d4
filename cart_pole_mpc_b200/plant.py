"""Plant-generic batched controller over the cpmpc_plant_* entry points (include/cpmpc.h): the cart-mounted double
pendulum of BASELINE.json configs[4] (physics spec: symbolic/dynamics_double.py:25-148 of the reference, which has no
optimiser for it), and the single pendulum through the same generic kernels as a cross-check.

Host side only; all arithmetic happens in libcpmpc.so (csrc/plant_solver_kernel.cuh).
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np

from . import _capi
from .batched import StepOutputs, _host

PLANTS = {"single": _capi.PLANT_SINGLE, "double": _capi.PLANT_DOUBLE}
_SCALAR_FIELDS = tuple(name for name, _ in _capi.CpmpcPlantParams._fields_ if name != "final_cost_weight")


def make_plant_params(plant: str = "double", final_cost_weight=None, **overrides) -> _capi.CpmpcPlantParams:
    """Defaults of OptimizationParams (optimization/optimization.hpp:12-48); terminal weights in state order, 150 on
    b_x and hard equalities (-1) on every angle and rate."""
    p = _capi.CpmpcPlantParams()
    _capi.lib().cpmpc_plant_default_params(PLANTS[plant], C.byref(p))
    for k, v in overrides.items():
        if k not in _SCALAR_FIELDS:
            raise KeyError(f"unknown parameter {k!r}")
        setattr(p, k, v)
    if final_cost_weight is not None:
        nx = _capi.lib().cpmpc_plant_state_dim(PLANTS[plant])
        w = [float(v) for v in final_cost_weight]
        if len(w) != nx:
            raise ValueError(f"final_cost_weight must have {nx} entries")
        for i, v in enumerate(w):
            p.final_cost_weight[i] = v
    return p


class PlantBatchedOptimization:
    """B controllers for `plant` ("double" or "single") stepped together on one GPU; same per-instance semantics as
    pendulum::Optimization (optimization/optimization.hpp:73-108) with NX states per shooting node."""

    def __init__(self, plant: str, batch: int, params: Optional[_capi.CpmpcPlantParams] = None, device: int = 0,
                 final_cost_weight=None, **overrides):
        if plant not in PLANTS:
            raise ValueError(f"unknown plant {plant!r}")
        self._h = C.c_void_p()
        self._lib = _capi.lib()
        self.plant = plant
        self.params = params if params is not None else make_plant_params(plant, final_cost_weight, **overrides)
        _capi.check(self._lib.cpmpc_plant_create(PLANTS[plant], C.byref(self.params), int(batch), int(device),
                                                 C.byref(self._h)))
        self.batch = int(batch)
        self.device = int(device)
        self.nx = int(self._lib.cpmpc_plant_state_dim(PLANTS[plant]))
        self.dyn_dim = int(self._lib.cpmpc_plant_dyn_dim(PLANTS[plant]))
        self.W = int(self.params.window_length)
        self.S = int(self.params.state_spacing)
        self.NS = self.W // self.S + 1
        self.N = int(self._lib.cpmpc_plant_solution_dim(self._h))
        self.max_iterations = int(self.params.max_iterations)

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            self._lib.cpmpc_plant_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def step(self, states, dyn_params, set_points, weights=None, want_predicted: bool = True) -> StepOutputs:
        B, W, nx = self.batch, self.W, self.nx
        states = _host(states, (B, nx), "states")
        dyn = _host(dyn_params, (B, self.dyn_dim), "dyn_params")
        sp = _host(set_points, (B,), "set_points")
        wts = _host(weights, (B, 2 + nx), "weights") if weights is not None else None
        u = np.empty((B, W), dtype=np.float64)
        pred = np.empty((B, W, nx), dtype=np.float64) if want_predicted else None
        stats = np.empty(B, dtype=_capi.STATS_DTYPE)
        _capi.check(self._lib.cpmpc_plant_step(
            self._h, states.ctypes.data, dyn.ctypes.data, sp.ctypes.data,
            wts.ctypes.data if wts is not None else None, u.ctypes.data,
            pred.ctypes.data if pred is not None else None, stats.ctypes.data))
        return StepOutputs(u, pred, stats)

    def step_device(self, states, dyn_params, set_points, u_out, weights=None, pred_out=None, stats_out=None,
                    stream=None) -> None:
        """Asynchronous launch on `stream` (a torch.cuda.Stream or None for the current stream); all tensors are
        contiguous float64 CUDA tensors on this handle's device."""
        import torch
        for name, t in (("states", states), ("dyn_params", dyn_params), ("set_points", set_points), ("u_out", u_out)):
            if not (t.is_cuda and t.dtype == torch.float64 and t.is_contiguous()):
                raise ValueError(f"{name} must be a contiguous float64 CUDA tensor")
        s = stream if stream is not None else torch.cuda.current_stream(self.device)
        _capi.check(self._lib.cpmpc_plant_step_device(
            self._h, states.data_ptr(), dyn_params.data_ptr(), set_points.data_ptr(),
            weights.data_ptr() if weights is not None else None, u_out.data_ptr(),
            pred_out.data_ptr() if pred_out is not None else None,
            stats_out.data_ptr() if stats_out is not None else None, s.cuda_stream))

    def reset(self) -> None:
        _capi.check(self._lib.cpmpc_plant_reset(self._h))

    def reset_device(self, stream=None) -> None:
        import torch
        s = stream if stream is not None else torch.cuda.current_stream(self.device)
        _capi.check(self._lib.cpmpc_plant_reset_device(self._h, s.cuda_stream))

    def set_previous_solution(self, solutions) -> None:
        sol = _host(solutions, (self.batch, self.N), "solutions")
        _capi.check(self._lib.cpmpc_plant_set_previous_solution(self._h, sol.ctypes.data))

    def get_previous_solution(self):
        sol = np.empty((self.batch, self.N), dtype=np.float64)
        has = np.empty(self.batch, dtype=np.uint8)
        _capi.check(self._lib.cpmpc_plant_get_previous_solution(self._h, sol.ctypes.data, has.ctypes.data))
        return sol, has.astype(bool)

    def enable_iteration_log(self, enable: bool = True) -> None:
        _capi.check(self._lib.cpmpc_plant_enable_iteration_log(self._h, 1 if enable else 0))

    def get_iteration_log(self) -> np.ndarray:
        log = np.empty((self.batch, self.max_iterations, _capi.ITERLOG_STRIDE), dtype=np.float64)
        _capi.check(self._lib.cpmpc_plant_get_iteration_log(self._h, log.ctypes.data))
        return log
