"""Synthetic batch generators for the benchmark configurations.

Counter-based RNG (splitmix64 of ``seed * 2**32 + instance * 16 + field``) so that any process — any rank, the
CPU baseline, a test — regenerates exactly the same per-instance inputs from (seed, instance id) alone.

Parameter ranges follow the web app's sliders (viz/src/application.ts:223-294) as tabulated in SURVEY.md 8(d).
"""
from __future__ import annotations

import math

import numpy as np

# application.ts:61-71 / model/scratch.py:23 — the default dynamics parameters
DEFAULT_DYN = (1.0, 0.1, 0.25, 9.81, 0.05, 0.1, 0.02, 0.8, 100.0)
# optimization_test.cc:22
TEST_DYN = (1.0, 0.1, 0.25, 9.81, 0.03, 0.1, 0.13, 0.8, 100.0)

SEED_CONFIG2 = 20241101


def splitmix64(x: np.ndarray) -> np.ndarray:
    """One splitmix64 output per 64-bit counter (vectorised, wrapping arithmetic)."""
    x = np.asarray(x, dtype=np.uint64)
    with np.errstate(over="ignore"):
        z = x + np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return z


def uniform(seed: int, first: int, count: int, field: int, lo: float, hi: float) -> np.ndarray:
    """U[lo, hi) for instances [first, first+count), stream ``field`` (0..15)."""
    idx = np.arange(first, first + count, dtype=np.uint64)
    with np.errstate(over="ignore"):
        ctr = np.uint64(seed) * np.uint64(1 << 32) + idx * np.uint64(16) + np.uint64(field)
    u01 = (splitmix64(ctr) >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)
    return lo + (hi - lo) * u01


def swingup_batch(batch: int, seed: int = SEED_CONFIG2, first: int = 0, dyn=DEFAULT_DYN):
    """Config 2: randomised initial states around hanging-down and randomised cart / pole masses.

    Returns (states[B,4], dyn_params[B,9], set_points[B]) as C-contiguous float64 arrays.
    """
    states = np.empty((batch, 4), dtype=np.float64)
    states[:, 0] = uniform(seed, first, batch, 0, -0.5, 0.5)
    states[:, 1] = -math.pi / 2 + uniform(seed, first, batch, 1, -0.5, 0.5)
    states[:, 2] = uniform(seed, first, batch, 2, -0.5, 0.5)
    states[:, 3] = uniform(seed, first, batch, 3, -1.0, 1.0)
    params = np.tile(np.asarray(dyn, dtype=np.float64), (batch, 1))
    params[:, 0] = uniform(seed, first, batch, 4, 0.1, 2.0)   # m_b
    params[:, 1] = uniform(seed, first, batch, 5, 0.1, 1.0)   # m_1
    set_points = np.zeros(batch, dtype=np.float64)
    return np.ascontiguousarray(states), np.ascontiguousarray(params), set_points


def sweep_batch(batch: int, seed: int = SEED_CONFIG2 + 2, first: int = 0, dyn=DEFAULT_DYN):
    """Config 4: dynamics-parameter and cost-weight sweep.

    Returns (states, dyn_params, set_points, weights[B,6]); weights columns are
    (u, u_derivative, b_x_final, th_final, b_x_dot_final, th_dot_final) with the last three left at -1 (equality).
    """
    states, params, _ = swingup_batch(batch, seed, first, dyn)
    params[:, 2] = uniform(seed, first, batch, 6, 0.05, 0.5)    # l_1
    params[:, 4] = uniform(seed, first, batch, 7, 0.01, 0.5)    # mu_b
    params[:, 6] = uniform(seed, first, batch, 8, 0.01, 0.15)   # c_d_1
    weights = np.empty((batch, 6), dtype=np.float64)
    weights[:, 0] = uniform(seed, first, batch, 9, 0.01, 1.0)
    weights[:, 1] = uniform(seed, first, batch, 10, 0.01, 1.0)
    weights[:, 2] = uniform(seed, first, batch, 11, 0.0, 200.0)
    weights[:, 3:] = -1.0
    set_points = uniform(seed, first, batch, 12, -1.0, 1.0)
    return states, params, np.ascontiguousarray(set_points), weights


# symbolic/dynamics_double.py:13-22 field order {m_b, m_1, m_2, l_1, l_2, g}; both poles like the single pendulum's
DEFAULT_DOUBLE_DYN = (1.0, 0.1, 0.1, 0.25, 0.25, 9.81)


def double_swingup_batch(batch: int, seed: int = SEED_CONFIG2 + 4, first: int = 0, dyn=DEFAULT_DOUBLE_DYN):
    """Config 5: cart-mounted double pendulum, randomised states around hanging-down (both angles near -pi/2) and
    randomised cart / pole masses (the single pendulum's slider ranges, application.ts:223-240, for both poles).

    Returns (states[B,6] = [b_x, th_1, th_2, b_x', th_1', th_2'], dyn_params[B,6], set_points[B]).
    """
    states = np.empty((batch, 6), dtype=np.float64)
    states[:, 0] = uniform(seed, first, batch, 0, -0.5, 0.5)
    states[:, 1] = -math.pi / 2 + uniform(seed, first, batch, 1, -0.5, 0.5)
    states[:, 2] = -math.pi / 2 + uniform(seed, first, batch, 2, -0.5, 0.5)
    states[:, 3] = uniform(seed, first, batch, 3, -0.5, 0.5)
    states[:, 4] = uniform(seed, first, batch, 4, -1.0, 1.0)
    states[:, 5] = uniform(seed, first, batch, 5, -1.0, 1.0)
    params = np.tile(np.asarray(dyn, dtype=np.float64), (batch, 1))
    params[:, 0] = uniform(seed, first, batch, 6, 0.1, 2.0)   # m_b
    params[:, 1] = uniform(seed, first, batch, 7, 0.1, 1.0)   # m_1
    params[:, 2] = uniform(seed, first, batch, 8, 0.1, 1.0)   # m_2
    set_points = np.zeros(batch, dtype=np.float64)
    return np.ascontiguousarray(states), np.ascontiguousarray(params), set_points
