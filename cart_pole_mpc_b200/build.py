"""Builds libcpmpc.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

The library is deliberately NOT a torch extension: the drop-in boundary is a plain C ABI (include/cpmpc.h) that the
reference's C++ / nanobind side can link without Python.  nvcc cross-compiles without a GPU.
"""
from __future__ import annotations

import os
import shutil
import subprocess

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
LIB_PATH = os.path.join(PKG_DIR, "libcpmpc.so")
SOURCES = ["cpmpc_capi.cu", "cpmpc_plant_capi.cu"]
HEADERS = ["solver_kernel.cuh", "plant_solver_kernel.cuh", "double_pendulum.cuh", "dynamics.cuh", "fastmath.cuh",
           "capi_common.hpp", os.path.join("..", "..", "include", "cpmpc.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC",
]


def _nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; libcpmpc.so cannot be built")


def is_stale() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    lib_m = os.path.getmtime(LIB_PATH)
    for f in SOURCES + HEADERS:
        if os.path.getmtime(os.path.join(CSRC, f)) > lib_m:
            return True
    return False


def build_library(force: bool = False, verbose: bool = False) -> str:
    """Compile the library if it is missing or older than its sources; returns its path."""
    if not force and not is_stale():
        return LIB_PATH
    env = dict(os.environ)
    # the image's CXX points at a wrapper without OpenMP; nvcc only needs a plain host g++
    env.pop("CXX", None)
    env.pop("CC", None)
    # one nvcc per translation unit, side by side; objects are only rebuilt when older than any source or header
    newest = max(os.path.getmtime(os.path.join(CSRC, f)) for f in SOURCES + HEADERS)
    objs, procs = [], []
    for src in SOURCES:
        obj = os.path.join(CSRC, src.replace(".cu", ".o"))
        objs.append(obj)
        if not force and os.path.exists(obj) and os.path.getmtime(obj) > newest:
            continue
        cmd = [_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", "-o", obj, src]
        procs.append(subprocess.Popen(cmd, cwd=CSRC, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True,
                                      env=env))
    for proc in procs:
        out, _ = proc.communicate()
        if proc.returncode != 0:
            raise RuntimeError("nvcc failed:\n" + out)
        if verbose:
            print(out)
    proc = subprocess.run([_nvcc(), "-shared", "-o", LIB_PATH] + objs, cwd=CSRC, capture_output=True, text=True,
                          env=env)
    if proc.returncode != 0:
        raise RuntimeError("nvcc link failed:\n" + proc.stdout + proc.stderr)
    return LIB_PATH


NATIVE_DIR = os.path.join(PKG_DIR, "native")
PYPENDULUM_SRC = os.path.join(PKG_DIR, "host", "pypendulum_ext.cc")


def pypendulum_path() -> str:
    import sysconfig
    return os.path.join(NATIVE_DIR, "pypendulum" + sysconfig.get_config_var("EXT_SUFFIX"))


def build_pypendulum(force: bool = False) -> str:
    """Compile the native `pypendulum` module (pybind11 over host/optimization.hpp, the binding list of the reference's
    wrapper/wrapper.cc:40-102) into cart_pole_mpc_b200/native/.  g++ only; it links libcpmpc.so by relative rpath."""
    import sysconfig
    import pybind11
    out = pypendulum_path()
    deps = [PYPENDULUM_SRC, os.path.join(PKG_DIR, "host", "optimization.hpp"), os.path.join(PKG_DIR, "host", "structs.hpp"),
            os.path.join(PKG_DIR, "..", "include", "cpmpc.h")]
    if not force and os.path.exists(out) and all(os.path.getmtime(out) > os.path.getmtime(d) for d in deps):
        return out
    build_library()
    os.makedirs(NATIVE_DIR, exist_ok=True)
    env = dict(os.environ)
    env.pop("CXX", None)
    cmd = ["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-fvisibility=hidden", "-I" + pybind11.get_include(),
           "-I" + sysconfig.get_paths()["include"], "-o", out, PYPENDULUM_SRC, "-L" + PKG_DIR, "-lcpmpc",
           "-Wl,-rpath,$ORIGIN/.."]
    proc = subprocess.run(cmd, capture_output=True, text=True, env=env)
    if proc.returncode != 0:
        raise RuntimeError("building pypendulum failed:\n" + proc.stdout + proc.stderr)
    return out


if __name__ == "__main__":
    print(build_library(force=True, verbose=True))
    print(build_pypendulum(force=True))
