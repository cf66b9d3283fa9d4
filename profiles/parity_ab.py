#!/usr/bin/env python
"""Parity fractions of a given build of libcpmpc.so against the COMPILED REFERENCE on the variant list of
tests/test_gpu_parity.py (REF_VARIANTS), on a larger sample than the test uses, to tell a real shift from sampling noise
when the device arithmetic changes.  Test infrastructure (imports oracle/).

    python profiles/parity_ab.py cart_pole_mpc_b200/libcpmpc.so [B]
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cart_pole_mpc_b200 import _capi  # noqa: E402

_capi.LIB_PATH = os.path.abspath(sys.argv[1])
from cart_pole_mpc_b200 import BatchedOptimization, synthetic  # noqa: E402
from oracle import oracle_py as O  # noqa: E402

B = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
VARIANTS = [
    dict(), dict(state_spacing=5, max_iterations=10), dict(state_spacing=40), dict(state_spacing=1, max_iterations=4),
    dict(window_length=20, state_spacing=5), dict(window_length=80, state_spacing=10), dict(control_dt=0.02),
    dict(u_guess_sinusoid_amplitude=0.0), dict(equality_penalty_initial=10.0),
    dict(th_final_cost_weight=50.0, b_x_dot_final_cost_weight=20.0, th_dot_final_cost_weight=20.0),
    dict(b_x_final_cost_weight=-1.0), dict(b_x_final_cost_weight=0.0), dict(u_derivative_cost_weight=0.0),
    dict(window_length=160),
]
for over in VARIANTS:
    st, dy, sp = synthetic.swingup_batch(B, seed=17)
    sp = sp + np.linspace(-0.3, 0.3, B)
    with O.using_ref():
        ref = O.batch_step(O.opt_array(**over), st, dy, sp, want_iter_log=False)
    s = BatchedOptimization(B, **over)
    out = s.step(st, dy, sp)
    s.close()
    g = out.stats["f"] + out.stats["penalty"] * out.stats["equality"]
    o = ref.stats[:, 3] + ref.stats[:, 5] * ref.stats[:, 4]
    rel = np.abs(g - o) / np.maximum(np.abs(o), 1e-300)
    print(json.dumps({"over": over, "B": B, "within_1e-8": round(float((rel <= 1e-8).mean()), 4),
                      "within_1e-6": round(float((rel <= 1e-6).mean()), 4),
                      "iter_match": round(float((out.stats["n_iterations"] == ref.stats[:, 1].astype(int)).mean()), 4)}),
          flush=True)
