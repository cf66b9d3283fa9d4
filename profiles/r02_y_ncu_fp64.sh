#!/bin/bash
# executed FP64 instruction counts of the final build of round 2: config 2 at 4096 (fused launch: one kernel per step)
# and 65536 instances (split launch: three kernels per step), config 4 at 16384 (split)
mkdir -p gpurun_out
M=smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,gpu__time_duration.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum
ncu --metrics $M --clock-control none -k regex:cpmpc_ -s 1 -c 1 --csv --log-file gpurun_out/r02_y_fp64_config2.csv \
    python profiles/prof_target.py 4096 2 > /dev/null 2>&1
ncu --metrics $M --clock-control none -k regex:cpmpc_ -s 3 -c 3 --csv --log-file gpurun_out/r02_y_fp64_config2_large.csv \
    python profiles/prof_target.py 65536 2 > /dev/null 2>&1
ncu --metrics $M --clock-control none -k regex:cpmpc_ -s 3 -c 3 --csv --log-file gpurun_out/r02_y_fp64_config4.csv \
    python profiles/prof_target_long.py 16384 2 > /dev/null 2>&1
python profiles/fp64_ops.py gpurun_out/r02_y_fp64_ops.json config2=gpurun_out/r02_y_fp64_config2.csv:4096 \
    config2_large=gpurun_out/r02_y_fp64_config2_large.csv:65536 config4=gpurun_out/r02_y_fp64_config4.csv:16384
