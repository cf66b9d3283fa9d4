#!/bin/bash
# round 2, build with split launch + compact layout: launch list of a 65536-instance step (three kernels), one full
# capture of the solve kernel at 65536 instances (8 warps per SM) and one at the headline batch (4096, fused launch).
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/r02_y_launches_65536.csv \
    python profiles/prof_target.py 65536 3 > gpurun_out/r02_y_ncu_launch.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:cpmpc_solve_kernel -s 2 -c 1 -f -o gpurun_out/r02_y_prof_large \
    python profiles/prof_target.py 65536 3 > gpurun_out/r02_y_ncu_prof_large.log 2>&1
echo "full capture (65536) rc=$?"
ncu --set full --clock-control none --import-source on -k regex:cpmpc_solve_kernel -s 2 -c 1 -f -o gpurun_out/r02_y_prof \
    python profiles/prof_target.py 4096 3 > gpurun_out/r02_y_ncu_prof.log 2>&1
echo "full capture (4096) rc=$?"
grep -v "^==" gpurun_out/r02_y_launches_65536.csv | tail -12
