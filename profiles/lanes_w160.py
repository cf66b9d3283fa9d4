import os, sys, json, numpy as np, torch
sys.path.insert(0, os.getcwd())
from cart_pole_mpc_b200 import BatchedOptimization, synthetic
B=32768
st,dy,sp,wt=synthetic.sweep_batch(B)
dev=torch.device("cuda",0)
to=lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
s=BatchedOptimization(B,device=0,window_length=160)
W=160
st_d,dy_d,sp_d,wt_d=to(st),to(dy),to(sp),to(wt)
u=torch.empty((B,W),dtype=torch.float64,device=dev); pred=torch.empty((B,W,4),dtype=torch.float64,device=dev); stats=torch.empty((B,56),dtype=torch.uint8,device=dev)
ms=[]
for _ in range(3):
    s.reset_device(); e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record(); s.step_device(st_d,dy_d,sp_d,u,weights=wt_d,pred_out=pred,stats_out=stats); e1.record(); torch.cuda.synchronize(); ms.append(e0.elapsed_time(e1))
print(os.environ.get("CPMPC_LANES"), "ms", min(ms), "solves/s", B/min(ms)*1e3, "checksum", float(u.sum()))
