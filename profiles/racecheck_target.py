#!/usr/bin/env python
"""Small runs for compute-sanitizer --tool racecheck / synccheck (shared-memory hazards between the lanes of a group):
roomy and compact layout, fused and split launch, the Levenberg-Marquardt-heavy variant (Riccati recursion) and W = 160.
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cart_pole_mpc_b200 import BatchedOptimization, synthetic  # noqa: E402

st, dy, sp = synthetic.swingup_batch(96, seed=13)
for name, env, kw in (("roomy fused", {}, dict(split_launch=False)),
                      ("compact split", {"CPMPC_ZGLOBAL": "1", "CPMPC_STAGE_KNOTS": "6"}, dict(split_launch=True)),
                      ("lm-heavy compact", {"CPMPC_ZGLOBAL": "1"}, dict(u_cost_weight=0.0))):
    for k in ("CPMPC_ZGLOBAL", "CPMPC_STAGE_KNOTS"):
        os.environ.pop(k, None)
    os.environ.update(env)
    s = BatchedOptimization(96, **kw)
    out = s.step(st, dy, sp)
    s.step(st + 0.01, dy, sp)
    print(name, "ok", float(out.stats["n_iterations"].mean()), flush=True)
    s.close()
for k in ("CPMPC_ZGLOBAL", "CPMPC_STAGE_KNOTS"):
    os.environ.pop(k, None)
st, dy, sp, wt = synthetic.sweep_batch(48)
s = BatchedOptimization(48, window_length=160)
out = s.step(st, dy, sp, weights=wt)
print("W=160 ok", float(out.stats["n_iterations"].mean()), flush=True)
s.close()
