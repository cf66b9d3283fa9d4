#!/usr/bin/env python
"""A/B harness for kernel changes that must not change results: runs the same seeded workloads through two builds of
libcpmpc.so (each in its own process), times the solve kernel on the device and compares every output bit for bit.

    python profiles/ab_compare.py run cart_pole_mpc_b200/libcpmpc_rl.so gpurun_out/ab_old.npz
    python profiles/ab_compare.py run cart_pole_mpc_b200/libcpmpc.so    gpurun_out/ab_new.npz
    python profiles/ab_compare.py diff gpurun_out/ab_old.npz gpurun_out/ab_new.npz

Workloads: config 2 (4096 cold swing-ups), an LM-heavy variant (u_cost_weight = 0: many damped iterations through the
Riccati QP), config 4 (W = 160, Riccati on every iteration, 8192 instances), config 5 (double pendulum, 8192), and the
two config-2 shards that contain a Levenberg-Marquardt instance (ranks 5 and 6 of the 8-GPU run).
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def run(lib_path, out_path):
    import torch
    from cart_pole_mpc_b200 import _capi
    _capi.LIB_PATH = os.path.abspath(lib_path)
    from cart_pole_mpc_b200 import BatchedOptimization, synthetic
    from cart_pole_mpc_b200.plant import PlantBatchedOptimization

    dev = torch.device("cuda", 0)
    results, times = {}, {}

    def timed(name, solver, st, dy, sp, wt=None, nx=4, reps=5):
        B, W = st.shape[0], solver.W if hasattr(solver, "W") else 40
        to = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
        st_d, dy_d, sp_d = to(st), to(dy), to(sp)
        wt_d = to(wt) if wt is not None else None
        u_d = torch.empty((B, W), dtype=torch.float64, device=dev)
        pred_d = torch.empty((B, W, nx), dtype=torch.float64, device=dev)
        stats_d = torch.empty((B, 56), dtype=torch.uint8, device=dev)
        ms = []
        for _ in range(reps):
            solver.reset_device()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            solver.step_device(st_d, dy_d, sp_d, u_d, weights=wt_d, pred_out=pred_d, stats_out=stats_d)
            e1.record()
            torch.cuda.synchronize()
            ms.append(e0.elapsed_time(e1))
        times[name] = float(np.median(ms[1:]))
        results[name + "/u"] = u_d.cpu().numpy()
        results[name + "/pred"] = pred_d.cpu().numpy()
        results[name + "/stats"] = stats_d.cpu().numpy()

    st, dy, sp = synthetic.swingup_batch(4096)
    s = BatchedOptimization(4096, device=0)
    timed("config2", s, st, dy, sp)
    s.close()
    for r in (5, 6):
        st, dy, sp = synthetic.swingup_batch(4096, first=r * 4096)
        s = BatchedOptimization(4096, device=0)
        timed(f"config2_shard{r}", s, st, dy, sp)
        s.close()
    st, dy, sp = synthetic.swingup_batch(4096, seed=13)
    s = BatchedOptimization(4096, device=0, u_cost_weight=0.0)
    timed("lm_heavy", s, st, dy, sp)
    s.close()
    st, dy, sp, wt = synthetic.sweep_batch(8192)
    s = BatchedOptimization(8192, device=0, window_length=160)
    timed("config4_w160", s, st, dy, sp, wt=wt, reps=3)
    s.close()
    st, dy, sp = synthetic.double_swingup_batch(8192)
    s = PlantBatchedOptimization("double", 8192, device=0)
    timed("config5_double", s, st, dy, sp, nx=6, reps=3)
    s.close()
    np.savez_compressed(out_path, times=json.dumps(times), **results)
    print(lib_path, json.dumps(times))


def diff(a_path, b_path):
    a, b = np.load(a_path), np.load(b_path)
    ta, tb = json.loads(str(a["times"])), json.loads(str(b["times"]))
    for k in ta:
        print(f"{k:20s} {ta[k]:9.3f} ms -> {tb[k]:9.3f} ms   x{ta[k] / tb[k]:.3f}")
    bad = 0
    for k in sorted(a.files):
        if k == "times":
            continue
        x, y = a[k], b[k]
        same = np.array_equal(x, y, equal_nan=True) if x.dtype.kind == "f" else np.array_equal(x, y)
        if not same:
            bad += 1
            if x.dtype.kind == "f":
                rows = np.where((x != y).reshape(x.shape[0], -1).any(axis=1))[0]
                print(f"  {k}: {rows.size} of {x.shape[0]} instances differ, max abs diff {np.nanmax(np.abs(x - y)):.3e}")
            else:
                rows = np.where((x != y).reshape(x.shape[0], -1).any(axis=1))[0]
                print(f"  {k}: {rows.size} of {x.shape[0]} stat records differ")
    print("BIT-IDENTICAL" if bad == 0 else f"{bad} arrays differ")
    return bad


if __name__ == "__main__":
    if sys.argv[1] == "run":
        run(sys.argv[2], sys.argv[3])
    else:
        sys.exit(1 if diff(sys.argv[2], sys.argv[3]) else 0)
