"""Host-side check of the DEVICE dynamics source (csrc/dynamics.cuh compiled with g++) against the oracle's restatement
of the generated dynamics and of integration.hpp, which tests/test_reference_pinning.py holds bit-equal to the compiled
reference: accelerations, their Jacobian rows, RK4 steps and RK4 step Jacobians on 200000 random states with every
model term active (tests/native/dynamics_check.cc states the bounds)."""
import os
import subprocess

from conftest import ROOT


def test_device_dynamics_source_against_oracle(tmp_path):
    exe = tmp_path / "dynamics_check"
    srcs = [os.path.join(ROOT, "tests", "native", "dynamics_check.cc"), os.path.join(ROOT, "oracle", "cartpole.cc"),
            os.path.join(ROOT, "oracle", "sqp.cc")]
    env = dict(os.environ)
    env.pop("CXX", None)
    subprocess.run(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-o", str(exe)] + srcs, check=True, env=env)
    proc = subprocess.run([str(exe)], capture_output=True, text=True, timeout=300)
    assert proc.returncode == 0, proc.stdout
