"""cart_pole_mpc_b200 — B200-native batched shooting-method MPC for the cart-pole (hot path of
gareth-cross/cart-pole-mpc: pendulum::Optimization::Step).

Layout: csrc/ holds the CUDA kernels and the C ABI (include/cpmpc.h -> libcpmpc.so); batched.py is the batched
host API over that ABI; pypendulum.py mirrors the reference's Python module; host/ holds the C++ facade that keeps
the reference's struct/class API.  Importing this package never touches oracle/.
"""
from .batched import (BatchedOptimization, StepOutputs, TERMINATION_NAMES, STEP_RESULT_NAMES, make_params,
                      measure_fp64_peak, sim_step, launch_count, debug_fastmath)

from .plant import PlantBatchedOptimization, make_plant_params

__all__ = ["PlantBatchedOptimization", "make_plant_params", "BatchedOptimization", "StepOutputs", "TERMINATION_NAMES", "STEP_RESULT_NAMES", "make_params",
           "measure_fp64_peak", "sim_step", "launch_count", "debug_fastmath"]
