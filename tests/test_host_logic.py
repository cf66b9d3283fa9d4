"""CPU tests of the host side: the C-ABI library loads and exports every declared symbol, argument validation
mirrors the reference's F_ASSERTs, the product never falls back to the CPU, synthetic inputs are reproducible, and
the sharded (N > 1) path works under gloo with world_size 2.  No compute call needs a GPU here."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT


def _declared_functions():
    text = open(os.path.join(ROOT, "include", "cpmpc.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(cpmpc_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(cuda_lib):
    from cart_pole_mpc_b200 import _capi
    declared = _declared_functions()
    assert declared, "no functions parsed from include/cpmpc.h"
    assert sorted(_capi.EXPORTED_SYMBOLS) == declared
    for name in declared:
        assert getattr(cuda_lib, name) is not None
    assert cuda_lib.cpmpc_version() == 1


def test_library_is_sm100a_and_has_no_oracle_dependency(cuda_lib):
    from cart_pole_mpc_b200 import _capi
    out = subprocess.run(["cuobjdump", "--list-elf", _capi.LIB_PATH], capture_output=True, text=True)
    if out.returncode == 0:
        assert "sm_100a" in out.stdout
    ldd = subprocess.run(["ldd", _capi.LIB_PATH], capture_output=True, text=True).stdout
    assert "oracle" not in ldd


def test_default_params_match_reference_defaults(cuda_lib):
    from cart_pole_mpc_b200 import make_params
    p = make_params()
    # optimization/optimization.hpp:12-48
    expect = dict(control_dt=0.01, window_length=40, state_spacing=10, max_iterations=8, relative_exit_tol=1e-5,
                  absolute_first_derivative_tol=1e-6, equality_penalty_initial=1.0, u_guess_sinusoid_amplitude=10.0,
                  u_cost_weight=0.1, u_derivative_cost_weight=0.1, b_x_final_cost_weight=150.0,
                  th_final_cost_weight=-1.0, b_x_dot_final_cost_weight=-1.0, th_dot_final_cost_weight=-1.0)
    for k, v in expect.items():
        assert getattr(p, k) == v


@pytest.mark.parametrize("bad", [
    dict(control_dt=0.0), dict(window_length=0), dict(window_length=40, state_spacing=7), dict(max_iterations=0),
    dict(u_cost_weight=-0.1), dict(u_derivative_cost_weight=-1.0), dict(state_spacing=0),
])
def test_invalid_params_are_rejected_like_the_reference(cuda_lib, bad):
    """optimization.cc:13-22 F_ASSERTs -> CPMPC_ERR_INVALID_ARGUMENT (ValueError in Python), checked before any CUDA
    call so it also holds on a machine without a GPU."""
    from cart_pole_mpc_b200 import BatchedOptimization
    with pytest.raises(ValueError):
        BatchedOptimization(4, **bad)


def test_stats_struct_layout(cuda_lib):
    from cart_pole_mpc_b200 import _capi
    assert C.sizeof(_capi.CpmpcStats) == 56 == np.dtype(_capi.STATS_DTYPE).itemsize
    assert C.sizeof(_capi.CpmpcParams) == 14 * 8


def test_no_cpu_fallback_without_gpu(cuda_lib):
    """Without a usable device the product path must fail loudly, never compute on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from cart_pole_mpc_b200 import BatchedOptimization, _capi
    with pytest.raises(_capi.CpmpcError):
        BatchedOptimization(4)


def test_product_package_never_imports_oracle():
    code = ("import sys; sys.path.insert(0, %r); import cart_pole_mpc_b200, cart_pole_mpc_b200.pypendulum, "
            "cart_pole_mpc_b200.sharding, cart_pole_mpc_b200.synthetic; "
            "assert not any(m == 'oracle' or m.startswith('oracle.') for m in sys.modules), 'oracle imported'") % ROOT
    subprocess.run([sys.executable, "-c", code], check=True)
    for dirpath, _, files in os.walk(os.path.join(ROOT, "cart_pole_mpc_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".h", ".cc")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle_py" not in text and "liboracle" not in text and "oracle/" not in text.replace(
                    "touches oracle/", ""), f


def test_synthetic_batches_are_counter_based():
    from cart_pole_mpc_b200 import synthetic
    s, d, p = synthetic.swingup_batch(64)
    s2, d2, _ = synthetic.swingup_batch(16, first=48)
    assert np.array_equal(s[48:], s2) and np.array_equal(d[48:], d2)
    assert np.all((d[:, 0] >= 0.1) & (d[:, 0] < 2.0)) and np.all((d[:, 1] >= 0.1) & (d[:, 1] < 1.0))
    assert np.all(np.abs(s[:, 0]) <= 0.5) and np.all(np.abs(s[:, 1] + np.pi / 2) <= 0.5)
    assert np.all(d[:, 2:] == np.asarray(synthetic.DEFAULT_DYN)[2:])
    st, dy, sp, w = synthetic.sweep_batch(32)
    assert w.shape == (32, 6) and np.all(w[:, 3:] == -1.0) and np.all((w[:, 0] >= 0.01) & (w[:, 0] < 1.0))
    # splitmix64 reference value (x = 0 -> first output of the canonical generator seeded with 0)
    assert int(synthetic.splitmix64(np.array([0], dtype=np.uint64))[0]) == 0xE220A8397B1DCDAF


def test_flop_model_matches_survey_numbers():
    sys.path.insert(0, ROOT)
    import bench
    # cold solve (n_iter 8, n_ls 19) ~ 1.27 Mflop; warm closed-loop (3.3, 3.4) ~ 0.47 Mflop
    assert abs(bench.algorithmic_flops(8, 19) / 1e6 - 1.27) < 0.02
    assert abs(bench.algorithmic_flops(3.3, 3.4) / 1e6 - 0.47) < 0.02
    assert 2500 < bench.hbm_bytes_per_solve() < 3000


def test_shard_ranges_cover_the_batch():
    from cart_pole_mpc_b200.sharding import shard_range
    for total in (1, 7, 4096, 1_000_000):
        for world in (1, 2, 3, 8):
            spans = [shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and sum(c for _, c in spans) == total
            for (f0, c0), (f1, _) in zip(spans, spans[1:]):
                assert f0 + c0 == f1
            assert max(c for _, c in spans) - min(c for _, c in spans) <= 1


_WORKER = r"""
import os, sys
sys.path.insert(0, %(root)r)
import numpy as np, torch, torch.distributed as dist
from cart_pole_mpc_b200 import synthetic
from cart_pole_mpc_b200.sharding import shard_range, all_gather_results
from oracle import oracle_py
dist.init_process_group(backend="gloo")
rank, world = dist.get_rank(), dist.get_world_size()
TOTAL = 37   # ragged on purpose
first, count = shard_range(TOTAL, rank, world)
st, dy, sp = synthetic.swingup_batch(count, first=first)
res = oracle_py.batch_step(oracle_py.opt_array(), st, dy, sp, n_threads=2, want_iter_log=False)   # stand-in for the per-rank GPU solve
u_all = all_gather_results(torch.from_numpy(res.u), TOTAL)
stats_all = all_gather_results(torch.from_numpy(res.stats), TOTAL)
if rank == 0:
    st, dy, sp = synthetic.swingup_batch(TOTAL)
    full = oracle_py.batch_step(oracle_py.opt_array(), st, dy, sp, n_threads=2, want_iter_log=False)
    assert u_all.shape == (TOTAL, 40)
    assert np.array_equal(u_all.numpy(), full.u), "sharded + gathered controls differ from the unsharded batch"
    assert np.array_equal(stats_all.numpy(), full.stats)
    print("GLOO_SHARD_OK")
dist.barrier()
dist.destroy_process_group()
"""


def test_sharded_batch_world_size_2_gloo(oracle, tmp_path):
    """The N>1 path: each rank regenerates and solves only its shard (no data-path collective), the epilogue
    all-gather reassembles results in instance order.  world_size 2, gloo, ragged shards."""
    script = tmp_path / "worker.py"
    script.write_text(_WORKER % {"root": ROOT})
    env = dict(os.environ, OMP_NUM_THREADS="2")
    proc = subprocess.run(
        [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
         "127.0.0.1", "--master-port", "29533", str(script)],
        capture_output=True, text=True, timeout=300, env=env)
    assert proc.returncode == 0, proc.stdout[-2000:] + proc.stderr[-2000:]
    assert "GLOO_SHARD_OK" in proc.stdout


def test_reference_arm_of_bench_prints_contract_line(oracle):
    import json
    proc = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                           "--warmup", "1", "--batch", "128"], capture_output=True, text=True, timeout=300)
    assert proc.returncode == 0, proc.stderr[-2000:]
    line = json.loads(proc.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "solves/s" and line["value"] > 0
    # "reference" when oracle/_ref/libref.so (the compiled reference) is present, otherwise the restatement
    assert line["cpu_baseline"]["kind"] == ("reference" if oracle.ref_available() else "port")
    assert line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["higher_is_better"] is True


def _build_facade_test(tmp_path):
    from cart_pole_mpc_b200 import _capi
    exe = tmp_path / "facade_closed_loop"
    pkg = os.path.dirname(_capi.LIB_PATH)
    subprocess.run(["g++", "-std=c++17", "-O2", "-Wall", "-Wextra", "-Werror", "-o", str(exe),
                    os.path.join(ROOT, "tests", "native", "facade_closed_loop.cc"), "-L" + pkg, "-lcpmpc",
                    "-Wl,-rpath," + pkg], check=True)
    return exe


def test_cpp_facade_compiles_against_the_c_abi(cuda_lib, tmp_path):
    """The reference-shaped C++ API (host/optimization.hpp) builds with -Wall -Wextra -Werror and links to
    libcpmpc.so; without a GPU the program must stop with the library's CUDA error, not compute anything."""
    import torch
    exe = _build_facade_test(tmp_path)
    if not torch.cuda.is_available():
        proc = subprocess.run([str(exe)], capture_output=True, text=True)
        assert proc.returncode != 0 and "FACADE_CLOSED_LOOP_OK" not in proc.stdout


def test_json_log_format_matches_the_reference_schema_and_round_trips():
    """The web front end's log format (optimization/wasm.cc:19-27,46-65; mini_opt/source/serialization.cc:32-134):
    key names, enum names and nesting, written from a solver log of the COMPILED REFERENCE (golden fixture), read
    back, and rendered by the ToString-style summary."""
    import json
    import numpy as np
    from conftest import load_golden
    from cart_pole_mpc_b200 import pypendulum as pp
    g = load_golden("ref_config1_default.npz")
    st = g["stats"][0]
    stats = {"termination_state": int(st[0]), "n_iterations": int(st[1]), "n_line_search_steps": int(st[2])}
    out = pp.OptimizationOutputs(pp.SingleCartPoleState(*g["states"][0]), [], g["u"][0].tolist(),
                                 [pp.SingleCartPoleState(*r) for r in g["predicted"][0]], stats, g["iter_log"][0])
    j = json.loads(out.to_json())
    assert list(j) == ["initial_state", "previous_solution", "solver_outputs", "u", "predicted_states"]
    assert set(j["initial_state"]) == {"b_x", "th_1", "th_1_dot", "b_x_dot"}
    assert j["solver_outputs"]["termination_state"] == "MAX_ITERATIONS"
    its = j["solver_outputs"]["iterations"]
    assert len(its) == 8 and len(j["u"]) == 40 and len(j["predicted_states"]) == 40
    assert list(its[0]) == ["iteration", "optimizer_state", "lambda", "errors_initial", "qp_outputs", "qp_eigenvalues",
                            "directional_derivatives", "penalty", "step_result", "line_search_steps"]
    assert its[0]["optimizer_state"] == "NOMINAL" and its[0]["qp_outputs"] == "SUCCESS" and its[0]["qp_eigenvalues"] is None
    assert set(its[0]["errors_initial"]) == {"f", "equality"}
    assert set(its[0]["directional_derivatives"]) == {"d_f", "d_equality"}
    assert set(its[0]["line_search_steps"][0]) == {"alpha", "errors"} and its[0]["line_search_steps"][0]["alpha"] == 1.0
    assert sum(len(it["line_search_steps"]) for it in its) == int(st[2])
    # doubles survive the text round trip bit for bit (repr round-trips IEEE doubles), and so does the summary
    back = pp.outputs_from_json(out.to_json())
    assert back.to_dict() == out.to_dict()
    assert back.solver_summary() == out.solver_summary()
    assert np.array_equal(np.array(back.u), g["u"][0])
    assert "termination = MAX_ITERATIONS" in back.solver_summary()


def test_header_is_plain_c99_and_links_from_c(cuda_lib, tmp_path):
    """include/cpmpc.h is the drop-in boundary: it must be consumable from plain C (what cgo / JNI / ctypes generators
    see), and a C program linked against libcpmpc.so must resolve its entry points.  Without a device the library
    fails loudly (status code + message), it never falls back to a CPU path."""
    from cart_pole_mpc_b200 import _capi
    src = os.path.join(ROOT, "tests", "native", "abi_c_check.c")
    inc = os.path.join(ROOT, "include")
    subprocess.run(["gcc", "-std=c99", "-pedantic", "-Wall", "-Werror", "-fsyntax-only", "-I" + inc, src], check=True)
    pkg = os.path.dirname(_capi.LIB_PATH)
    exe = tmp_path / "abi_c_check"
    subprocess.run(["gcc", "-std=c99", "-I" + inc, src, "-L" + pkg, "-lcpmpc", "-Wl,-rpath," + pkg, "-o", str(exe)],
                   check=True)
    proc = subprocess.run([str(exe)], capture_output=True, text=True, timeout=60)
    assert proc.returncode == 0 and "ABI_C_OK version" in proc.stdout, proc.stdout + proc.stderr


def test_bench_clock_sampler_window_and_thread_count():
    """bench.py helpers that decide what the JSON line says: samples inside the timed window win, an empty window
    falls back to the samples around it (and says so), throttle reasons are collected, and the CPU legs ask for every
    host thread explicitly (torchrun exports OMP_NUM_THREADS=1)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)

    class FakeProc:
        def terminate(self):
            pass

        def wait(self, timeout=None):
            return 0

    s = bench.ClockSampler(0)
    s.proc = FakeProc()
    s.lines = ["1200, 1965, 300.0, Not Active, Not Active, Not Active, Not Active",
               "1965, 1965, 700.0, Not Active, Not Active, Not Active, Active",
               "1950, 1965, 710.0, Not Active, Not Active, Not Active, Not Active"]
    s.stamps = [1.0, 2.0, 2.5]
    inside = s.stop(window=(1.5, 3.0))
    assert inside["sm_mhz"] == 1957.5 and inside["samples"] == 2 and inside["scope"] == "timed region"
    assert inside["reasons"] == ["sw_power_cap"] and inside["sm_max_mhz"] == 1965.0
    s.proc = FakeProc()
    empty = s.stop(window=(5.0, 6.0))
    assert empty["samples"] == 3 and "shorter than one sample" in empty["scope"]
    old = os.environ.get("OMP_NUM_THREADS")
    os.environ["OMP_NUM_THREADS"] = "1"
    try:
        assert bench.host_threads() == len(os.sched_getaffinity(0)) >= 1
    finally:
        if old is None:
            del os.environ["OMP_NUM_THREADS"]
        else:
            os.environ["OMP_NUM_THREADS"] = old
    # the flop model the roofline uses reproduces the survey's two reference points
    assert abs(bench.algorithmic_flops(8, 19) - 1.27e6) < 0.02e6
    assert abs(bench.algorithmic_flops(3.3, 3.4) - 0.47e6) < 0.02e6


def test_reference_arm_under_torchrun_runs_on_rank_zero_with_all_threads(oracle):
    """The driver launches `bench.py --impl reference` like the GPU arm (torchrun for N > 1): rank 0 alone prints the
    line, the other ranks exit 0 without work, and the CPU run is not throttled to one thread by torchrun's
    OMP_NUM_THREADS=1."""
    import json
    proc = subprocess.run(
        [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
         "127.0.0.1", "--master-port", "29541", os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
         "--steps", "1", "--warmup", "1", "--batch", "128"], capture_output=True, text=True, timeout=300)
    assert proc.returncode == 0, proc.stdout[-1500:] + proc.stderr[-1500:]
    lines = [ln for ln in proc.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1, lines
    line = json.loads(lines[0])
    assert line["impl"] == "reference" and line["n_gpus"] == 2 and line["value"] > 0
    assert line["cpu_baseline"]["cores"] == len(os.sched_getaffinity(0))
