#!/bin/bash
python profiles/ab_compare.py run cart_pole_mpc_b200/libcpmpc_base.so gpurun_out/ab_old.npz > gpurun_out/t15_ab_old.log 2>&1; tail -1 gpurun_out/t15_ab_old.log
python profiles/ab_compare.py run cart_pole_mpc_b200/libcpmpc.so gpurun_out/ab_new.npz > gpurun_out/t15_ab_new.log 2>&1; tail -1 gpurun_out/t15_ab_new.log
python profiles/ab_compare.py diff gpurun_out/ab_old.npz gpurun_out/ab_new.npz
for i in 1 2; do
SWEEP_LIB=cart_pole_mpc_b200/libcpmpc_base.so python profiles/lane_sweep.py 4096 65536
python profiles/lane_sweep.py 4096 65536
done
rm -f gpurun_out/ab_old.npz gpurun_out/ab_new.npz
