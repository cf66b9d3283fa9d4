#!/bin/bash
# round 2, run a: where the CUDA path parts from the compiled reference (teacher-forced closed loop, W=160, first-QP
# blocks), lane-mapping sweep of the unchanged round-1 kernel, and a baseline bench line on this box.
mkdir -p gpurun_out
python tests/parity_diagnose.py --plants 256 --ticks 100 --long 512 --blocks 48 --json gpurun_out/r02_a_parity_diagnose.json > gpurun_out/r02_a_diag.log 2>&1
echo "diag rc=$?"
tail -12 gpurun_out/r02_a_diag.log | cut -c1-1500
for L in 4 8; do CPMPC_LANES=$L python profiles/lane_sweep.py 4096 65536 >> gpurun_out/r02_a_lane_sweep.jsonl 2>&1; done
for Wp in 3 4 5; do CPMPC_WARPS_PER_BLOCK=$Wp python profiles/lane_sweep.py 65536 >> gpurun_out/r02_a_lane_sweep.jsonl 2>&1; done
cat gpurun_out/r02_a_lane_sweep.jsonl
python bench.py > gpurun_out/r02_a_bench.json 2> gpurun_out/r02_a_bench.err
echo "bench rc=$?"
python -c "
import json
d=json.loads(open('gpurun_out/r02_a_bench.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','e2e','roofline','parity') if k in d})
"
