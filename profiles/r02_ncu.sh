#!/bin/bash
# round 2 ncu session (one GPU): launch list of the bench command, one full capture of the headline kernel, executed
# FP64 instruction counts of config 2 and config 4.  Numbers printed by runs under ncu are never bench values.
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-other-configs --no-cpu-baseline > gpurun_out/r02_ncu_bench.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:cpmpc_solve_kernel -s 2 -c 1 -f -o gpurun_out/r02_prof \
    python profiles/prof_target.py 4096 3 > gpurun_out/r02_ncu_prof.log 2>&1
echo "full capture rc=$?"
M=smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,gpu__time_duration.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum
ncu --metrics $M --clock-control none -k regex:cpmpc_solve_kernel -s 1 -c 1 --csv --log-file gpurun_out/r02_fp64_config2.csv \
    python profiles/prof_target.py 4096 2 > /dev/null 2>&1
ncu --metrics $M --clock-control none -k regex:cpmpc_solve_kernel -s 1 -c 1 --csv --log-file gpurun_out/r02_fp64_config2_large.csv \
    python profiles/prof_target.py 65536 2 > /dev/null 2>&1
ncu --metrics $M --clock-control none -k regex:cpmpc_solve_kernel -s 1 -c 1 --csv --log-file gpurun_out/r02_fp64_config4.csv \
    python profiles/prof_target_long.py 16384 2 > /dev/null 2>&1
echo "fp64 counts rc=$?"
tail -3 gpurun_out/r02_fp64_config2.csv gpurun_out/r02_fp64_config4.csv
ls -la gpurun_out/r02_prof.ncu-rep
