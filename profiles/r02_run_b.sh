#!/bin/bash
# round 2, run b: parity diagnosis (fixed), first build with the warp-cooperative Riccati recursion for narrow groups
mkdir -p gpurun_out
python tests/parity_diagnose.py --plants 256 --ticks 100 --long 512 --blocks 48 --json gpurun_out/r02_b_parity_diagnose.json > gpurun_out/r02_b_diag.log 2>&1
echo "diag rc=$?"
grep -v '^\[blocks\]' gpurun_out/r02_b_diag.log | cut -c1-1800 | tail -12
python -m pytest tests -m gpu -x -q > gpurun_out/r02_b_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r02_b_pytest.log
python bench.py > gpurun_out/r02_b_bench.json 2> gpurun_out/r02_b_bench.err
echo "bench rc=$?"
python -c "
import json
d=json.loads(open('gpurun_out/r02_b_bench.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','parity') if k in d}, d['e2e']['value'], d['roofline']['frac'])
"
