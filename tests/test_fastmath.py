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
