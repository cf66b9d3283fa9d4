"""ctypes binding of libcpmpc.so (include/cpmpc.h).  Fails loudly when the CUDA library is missing: there is no
CPU fallback anywhere in this package."""
from __future__ import annotations

import ctypes as C
import os

_PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG_DIR, "libcpmpc.so")

ITERLOG_STRIDE = 27
WEIGHTS_STRIDE = 6

# every symbol include/cpmpc.h declares
FLAG_LOW_LATENCY = 1  # CPMPC_FLAG_LOW_LATENCY
FLAG_MIN_LATENCY = 2  # CPMPC_FLAG_MIN_LATENCY
FLAG_SPLIT_LAUNCH = 4  # CPMPC_FLAG_SPLIT_LAUNCH
FLAG_FUSED_LAUNCH = 8  # CPMPC_FLAG_FUSED_LAUNCH

EXPORTED_SYMBOLS = (
    "cpmpc_last_error", "cpmpc_version", "cpmpc_default_params", "cpmpc_create", "cpmpc_create_ex", "cpmpc_destroy", "cpmpc_batch",
    "cpmpc_solution_dim", "cpmpc_step", "cpmpc_step_async", "cpmpc_wait", "cpmpc_step_device", "cpmpc_reset", "cpmpc_reset_device",
    "cpmpc_set_previous_solution",
    "cpmpc_get_previous_solution", "cpmpc_enable_iteration_log", "cpmpc_get_iteration_log", "cpmpc_sim_step_device",
    "cpmpc_sim_step", "cpmpc_closed_loop", "cpmpc_closed_loop_device", "cpmpc_launch_count",
    "cpmpc_measure_fp64_peak", "cpmpc_debug_fastmath", "cpmpc_debug_stride", "cpmpc_debug_first_qp", "cpmpc_get_warp_cycles", "cpmpc_enable_phase_clocks",
    "cpmpc_plant_state_dim", "cpmpc_plant_dyn_dim", "cpmpc_plant_default_params", "cpmpc_plant_create",
    "cpmpc_plant_destroy", "cpmpc_plant_solution_dim", "cpmpc_plant_step", "cpmpc_plant_step_device",
    "cpmpc_plant_reset", "cpmpc_plant_reset_device", "cpmpc_plant_set_previous_solution",
    "cpmpc_plant_get_previous_solution", "cpmpc_plant_enable_iteration_log", "cpmpc_plant_get_iteration_log",
)

PLANT_SINGLE = 0  # CPMPC_PLANT_SINGLE
PLANT_DOUBLE = 1  # CPMPC_PLANT_DOUBLE
PLANT_MAX_NX = 6


class CpmpcPlantParams(C.Structure):
    """struct cpmpc_plant_params: OptimizationParams with the terminal weights as an array in state order."""
    _fields_ = [
        ("control_dt", C.c_double),
        ("window_length", C.c_int64),
        ("state_spacing", C.c_int64),
        ("max_iterations", C.c_int64),
        ("relative_exit_tol", C.c_double),
        ("absolute_first_derivative_tol", C.c_double),
        ("equality_penalty_initial", C.c_double),
        ("u_guess_sinusoid_amplitude", C.c_double),
        ("u_cost_weight", C.c_double),
        ("u_derivative_cost_weight", C.c_double),
        ("final_cost_weight", C.c_double * PLANT_MAX_NX),
    ]


class CpmpcParams(C.Structure):
    """struct cpmpc_params == pendulum::OptimizationParams (optimization/optimization.hpp:12-48)."""
    _fields_ = [
        ("control_dt", C.c_double),
        ("window_length", C.c_int64),
        ("state_spacing", C.c_int64),
        ("max_iterations", C.c_int64),
        ("relative_exit_tol", C.c_double),
        ("absolute_first_derivative_tol", C.c_double),
        ("equality_penalty_initial", C.c_double),
        ("u_guess_sinusoid_amplitude", C.c_double),
        ("u_cost_weight", C.c_double),
        ("u_derivative_cost_weight", C.c_double),
        ("b_x_final_cost_weight", C.c_double),
        ("th_final_cost_weight", C.c_double),
        ("b_x_dot_final_cost_weight", C.c_double),
        ("th_dot_final_cost_weight", C.c_double),
    ]


class CpmpcStats(C.Structure):
    """struct cpmpc_stats (56 bytes)."""
    _fields_ = [
        ("termination_state", C.c_int32),
        ("n_iterations", C.c_int32),
        ("n_line_search_steps", C.c_int32),
        ("n_failed_line_searches", C.c_int32),
        ("last_step_result", C.c_int32),
        ("optimizer_state", C.c_int32),
        ("f", C.c_double),
        ("equality", C.c_double),
        ("penalty", C.c_double),
        ("lambda_", C.c_double),
    ]


STATS_DTYPE = [
    ("termination_state", "<i4"), ("n_iterations", "<i4"), ("n_line_search_steps", "<i4"),
    ("n_failed_line_searches", "<i4"), ("last_step_result", "<i4"), ("optimizer_state", "<i4"),
    ("f", "<f8"), ("equality", "<f8"), ("penalty", "<f8"), ("lambda", "<f8"),
]

_lib = None


class CpmpcError(RuntimeError):
    pass


def lib() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise CpmpcError(
            f"{LIB_PATH} is missing: build it with `python -m cart_pole_mpc_b200.build` (nvcc, sm_100a). "
            "There is no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    dp, vp = C.POINTER(C.c_double), C.c_void_p
    u8p, i64p = C.POINTER(C.c_uint8), C.POINTER(C.c_int64)
    L.cpmpc_last_error.restype = C.c_char_p
    L.cpmpc_last_error.argtypes = []
    L.cpmpc_version.restype = C.c_int
    L.cpmpc_default_params.restype = None
    L.cpmpc_default_params.argtypes = [C.POINTER(CpmpcParams)]
    L.cpmpc_create.restype = C.c_int
    L.cpmpc_create.argtypes = [C.POINTER(CpmpcParams), C.c_int64, C.c_int, C.POINTER(vp)]
    L.cpmpc_create_ex.restype = C.c_int
    L.cpmpc_create_ex.argtypes = [C.POINTER(CpmpcParams), C.c_int64, C.c_int, C.c_uint32, C.POINTER(vp)]
    L.cpmpc_destroy.restype = C.c_int
    L.cpmpc_destroy.argtypes = [vp]
    L.cpmpc_batch.restype = C.c_int64
    L.cpmpc_batch.argtypes = [vp]
    L.cpmpc_solution_dim.restype = C.c_int64
    L.cpmpc_solution_dim.argtypes = [vp]
    # pointers are passed as void* so both numpy (host) and torch data_ptr() (device) addresses can be given
    L.cpmpc_step.restype = C.c_int
    L.cpmpc_step.argtypes = [vp] * 8
    L.cpmpc_step_async.restype = C.c_int
    L.cpmpc_step_async.argtypes = [vp] * 8 + [C.c_uint32]
    L.cpmpc_wait.restype = C.c_int
    L.cpmpc_wait.argtypes = [vp]
    L.cpmpc_step_device.restype = C.c_int
    L.cpmpc_step_device.argtypes = [vp] * 9
    L.cpmpc_reset.restype = C.c_int
    L.cpmpc_reset.argtypes = [vp, vp]
    L.cpmpc_reset_device.restype = C.c_int
    L.cpmpc_reset_device.argtypes = [vp, vp]
    L.cpmpc_set_previous_solution.restype = C.c_int
    L.cpmpc_set_previous_solution.argtypes = [vp, vp, vp]
    L.cpmpc_get_previous_solution.restype = C.c_int
    L.cpmpc_get_previous_solution.argtypes = [vp, vp, vp]
    L.cpmpc_enable_iteration_log.restype = C.c_int
    L.cpmpc_enable_iteration_log.argtypes = [vp, C.c_int]
    L.cpmpc_get_iteration_log.restype = C.c_int
    L.cpmpc_get_iteration_log.argtypes = [vp, vp]
    L.cpmpc_sim_step_device.restype = C.c_int
    L.cpmpc_sim_step_device.argtypes = [C.c_int64, vp, vp, vp, C.c_double, vp, vp, vp]
    L.cpmpc_sim_step.restype = C.c_int
    L.cpmpc_sim_step.argtypes = [C.c_int, C.c_int64, vp, vp, vp, C.c_double, vp, vp]
    L.cpmpc_closed_loop.restype = C.c_int
    L.cpmpc_closed_loop.argtypes = [vp, C.c_int64, vp, vp, vp, vp, vp, vp]
    L.cpmpc_closed_loop_device.restype = C.c_int
    L.cpmpc_closed_loop_device.argtypes = [vp, C.c_int64, vp, vp, vp, vp, vp, vp, vp]
    L.cpmpc_launch_count.restype = C.c_int64
    L.cpmpc_launch_count.argtypes = []
    L.cpmpc_measure_fp64_peak.restype = C.c_int
    L.cpmpc_measure_fp64_peak.argtypes = [C.c_int, dp]
    L.cpmpc_debug_fastmath.restype = C.c_int
    L.cpmpc_debug_fastmath.argtypes = [C.c_int, C.c_int64, C.c_void_p, C.c_void_p]
    L.cpmpc_enable_phase_clocks.restype = C.c_int
    L.cpmpc_enable_phase_clocks.argtypes = [vp, C.c_int]
    L.cpmpc_get_warp_cycles.restype = C.c_int64
    L.cpmpc_get_warp_cycles.argtypes = [vp, vp, C.c_int64]
    L.cpmpc_debug_stride.restype = C.c_int64
    L.cpmpc_debug_stride.argtypes = [vp]
    L.cpmpc_debug_first_qp.restype = C.c_int
    L.cpmpc_debug_first_qp.argtypes = [vp, vp, vp, vp, vp, C.c_double, C.c_int, vp]
    L.cpmpc_plant_state_dim.restype = C.c_int
    L.cpmpc_plant_state_dim.argtypes = [C.c_int]
    L.cpmpc_plant_dyn_dim.restype = C.c_int
    L.cpmpc_plant_dyn_dim.argtypes = [C.c_int]
    L.cpmpc_plant_default_params.restype = None
    L.cpmpc_plant_default_params.argtypes = [C.c_int, C.POINTER(CpmpcPlantParams)]
    L.cpmpc_plant_create.restype = C.c_int
    L.cpmpc_plant_create.argtypes = [C.c_int, C.POINTER(CpmpcPlantParams), C.c_int64, C.c_int, C.POINTER(vp)]
    L.cpmpc_plant_destroy.restype = C.c_int
    L.cpmpc_plant_destroy.argtypes = [vp]
    L.cpmpc_plant_solution_dim.restype = C.c_int64
    L.cpmpc_plant_solution_dim.argtypes = [vp]
    L.cpmpc_plant_step.restype = C.c_int
    L.cpmpc_plant_step.argtypes = [vp] * 8
    L.cpmpc_plant_step_device.restype = C.c_int
    L.cpmpc_plant_step_device.argtypes = [vp] * 9
    L.cpmpc_plant_reset.restype = C.c_int
    L.cpmpc_plant_reset.argtypes = [vp]
    L.cpmpc_plant_reset_device.restype = C.c_int
    L.cpmpc_plant_reset_device.argtypes = [vp, vp]
    L.cpmpc_plant_set_previous_solution.restype = C.c_int
    L.cpmpc_plant_set_previous_solution.argtypes = [vp, vp]
    L.cpmpc_plant_get_previous_solution.restype = C.c_int
    L.cpmpc_plant_get_previous_solution.argtypes = [vp, vp, vp]
    L.cpmpc_plant_enable_iteration_log.restype = C.c_int
    L.cpmpc_plant_enable_iteration_log.argtypes = [vp, C.c_int]
    L.cpmpc_plant_get_iteration_log.restype = C.c_int
    L.cpmpc_plant_get_iteration_log.argtypes = [vp, vp]
    _lib = L
    return L


def check(rc: int) -> None:
    if rc != 0:
        msg = lib().cpmpc_last_error().decode("utf-8", "replace")
        if rc == 1:
            raise ValueError(msg)  # precondition failures: the reference throws assert::default_error
        raise CpmpcError(f"cpmpc error {rc}: {msg}")
