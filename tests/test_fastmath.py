"""Host-side accuracy check of the branch-free FP64 helpers used by the CUDA dynamics (csrc/fastmath.cuh):
the Newton / polynomial stages are compiled for the host with g++ and compared against libm."""
import os
import subprocess

from conftest import ROOT


def test_fastmath_accuracy(tmp_path):
    exe = tmp_path / "fastmath_check"
    src = os.path.join(ROOT, "tests", "native", "fastmath_check.cc")
    subprocess.run(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-o", str(exe), src], check=True)
    proc = subprocess.run([str(exe)], capture_output=True, text=True, timeout=300)
    assert proc.returncode == 0, proc.stdout


import numpy as np
import pytest


@pytest.mark.gpu
def test_fastmath_accuracy_on_the_device():
    """The same functions evaluated ON THE DEVICE (cpmpc_debug_fastmath: hardware rcp / rsqrt seed instructions instead of
    the host build's float-precision seeds) against numpy's libm results: sin / cos within 2 units of 2^-53 in absolute
    terms on |x| < 40, tanh within 4e-16 absolute, the reciprocal within 1 ulp, the square root within 1 ulp and the
    reciprocal square root within 2 ulp over 60 binades."""
    from cart_pole_mpc_b200 import debug_fastmath
    rng = np.random.default_rng(7)
    n = 1 << 20
    ang = rng.uniform(-40.0, 40.0, n)
    out = debug_fastmath(ang)
    u = 2.0 ** -53
    assert np.abs(out[:, 0] - np.sin(ang)).max() <= 2.0 * u
    assert np.abs(out[:, 1] - np.cos(ang)).max() <= 2.0 * u
    vel = np.concatenate([rng.uniform(-60.0, 60.0, n // 2), rng.uniform(-60.0, 60.0, n // 2) * 1e-6])
    assert np.abs(debug_fastmath(vel)[:, 2] - np.tanh(vel)).max() < 4e-16
    x = np.exp(rng.uniform(-30.0, 30.0, n)) * rng.choice([-1.0, 1.0], n)
    out = debug_fastmath(x)
    ulp = lambda got, want: np.abs(got - want) / np.spacing(np.abs(want))
    assert ulp(out[:, 3], 1.0 / x).max() <= 1.0
    assert ulp(out[:, 4], np.sqrt(np.abs(x))).max() <= 1.0
    assert ulp(out[:, 5], 1.0 / np.sqrt(np.abs(x))).max() <= 2.0
    edge = debug_fastmath(np.array([0.0, 1e300, -1e300, np.inf]))
    assert edge[0, 0] == 0.0 and edge[0, 1] == 1.0 and edge[0, 2] == 0.0
    assert edge[1, 2] == 1.0 and edge[2, 2] == -1.0 and np.isnan(edge[3, 0]) and np.isnan(edge[3, 1])
