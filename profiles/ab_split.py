#!/usr/bin/env python
"""Fused launch against split launch (CPMPC_SPLIT=0 / 1, csrc/cpmpc_capi.cu Launch): the same seeded workloads through
both, every output compared bit for bit, the step timed on the device (CUDA events on the launch stream).

    python profiles/ab_split.py [config2 config2_warm config4 closed_loop ...]

Workloads: config 2 at 16384 / 65536 / 262144 cold instances, a warm-started second step, config 4 (W = 160, 32768
instances, per-instance weights) and the closed loop (16384 plants x 6 ticks).
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from cart_pole_mpc_b200 import BatchedOptimization, _capi, synthetic  # noqa: E402

DEV = torch.device("cuda:0")


def to(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


def step_workload(B, split, reps=5, warm=False, **kw):
    os.environ["CPMPC_SPLIT"] = "1" if split else "0"
    if kw.get("window_length", 40) == 160:
        st, dy, sp, wt = synthetic.sweep_batch(B)
        wt_d = to(wt)
    else:
        st, dy, sp = synthetic.swingup_batch(B)
        wt_d = None
    s = BatchedOptimization(B, device=0, **kw)
    W = s.W
    st_d, dy_d, sp_d = to(st), to(dy), to(sp)
    u = torch.empty((B, W), dtype=torch.float64, device=DEV)
    pred = torch.empty((B, W, 4), dtype=torch.float64, device=DEV)
    stats = torch.empty(B * np.dtype(_capi.STATS_DTYPE).itemsize, dtype=torch.uint8, device=DEV)
    ms = []
    for _ in range(reps):
        s.reset_device()
        if warm:
            s.step_device(st_d, dy_d, sp_d, u, weights=wt_d, pred_out=pred, stats_out=stats)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        s.step_device(st_d, dy_d, sp_d, u, weights=wt_d, pred_out=pred, stats_out=stats)
        e1.record()
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    prev, has = s.get_previous_solution()
    out = {"u": u.cpu().numpy(), "pred": pred.cpu().numpy(), "stats": stats.cpu().numpy(), "prev": prev, "has": has}
    s.close()
    return float(np.median(ms[1:])), out


def plant_workload(B, split, reps=3):
    from cart_pole_mpc_b200.plant import PlantBatchedOptimization
    os.environ["CPMPC_SPLIT"] = "1" if split else "0"
    st, dy, sp = synthetic.double_swingup_batch(B)
    s = PlantBatchedOptimization("double", B, device=0)
    st_d, dy_d, sp_d = to(st), to(dy), to(sp)
    u = torch.empty((B, 40), dtype=torch.float64, device=DEV)
    pred = torch.empty((B, 40, 6), dtype=torch.float64, device=DEV)
    stats = torch.empty(B * np.dtype(_capi.STATS_DTYPE).itemsize, dtype=torch.uint8, device=DEV)
    ms = []
    for _ in range(reps):
        s.reset_device()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        s.step_device(st_d, dy_d, sp_d, u, pred_out=pred, stats_out=stats)
        e1.record()
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    out = {"u": u.cpu().numpy(), "pred": pred.cpu().numpy(), "stats": stats.cpu().numpy()}
    s.close()
    return float(np.median(ms[1:])), out


def loop_workload(B, ticks, split, reps=3):
    os.environ["CPMPC_SPLIT"] = "1" if split else "0"
    st, dy, sp = synthetic.swingup_batch(B)
    s = BatchedOptimization(B, device=0)
    dy_d, sp_d = to(dy), to(sp)
    ctl = torch.empty((B, ticks), dtype=torch.float64, device=DEV)
    cnt = torch.zeros(4, dtype=torch.int64, device=DEV)
    ms = []
    for _ in range(reps):
        s.reset_device()
        st_d = to(st)
        cnt.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        s.closed_loop_device(ticks, st_d, dy_d, sp_d, controls_out=ctl, stats_sum=cnt)
        e1.record()
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    prev, has = s.get_previous_solution()
    out = {"states": st_d.cpu().numpy(), "controls": ctl.cpu().numpy(), "counts": cnt.cpu().numpy(), "prev": prev, "has": has}
    s.close()
    return float(np.median(ms[1:])), out


WORKLOADS = {
    "config2_8192": lambda sp: step_workload(8192, sp),
    "config2_10240": lambda sp: step_workload(10240, sp),
    "config2_12288": lambda sp: step_workload(12288, sp),
    "config2_16384": lambda sp: step_workload(16384, sp),
    "config2_24576": lambda sp: step_workload(24576, sp),
    "closed_loop_32768x10": lambda sp: loop_workload(32768, 10, sp),
    "config2_65536": lambda sp: step_workload(65536, sp),
    "config2_262144": lambda sp: step_workload(262144, sp, reps=3),
    "config2_warm_65536": lambda sp: step_workload(65536, sp, warm=True),
    "config4_w160_32768": lambda sp: step_workload(32768, sp, reps=3, window_length=160),
    "closed_loop_16384x6": lambda sp: loop_workload(16384, 6, sp),
    "closed_loop_65536x10": lambda sp: loop_workload(65536, 10, sp),
    "config5_double_65536": lambda sp: plant_workload(65536, sp),
    "config5_double_262144": lambda sp: plant_workload(262144, sp),
}

if __name__ == "__main__":
    names = sys.argv[1:] or list(WORKLOADS)
    for name in names:
        t0, a = WORKLOADS[name](False)
        t1, b = WORKLOADS[name](True)
        diffs = {k: int((a[k] != b[k]).sum()) if a[k].dtype.kind != "f"
                 else int((~((a[k] == b[k]) | (np.isnan(a[k]) & np.isnan(b[k])))).sum()) for k in a}
        print(json.dumps({"workload": name, "fused_ms": round(t0, 4), "split_ms": round(t1, 4),
                          "speedup": round(t0 / t1, 4), "differing_elements": diffs,
                          "bit_identical": all(v == 0 for v in diffs.values())}), flush=True)
