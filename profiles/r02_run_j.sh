#!/bin/bash
for L in 16 8 4 32; do
CPMPC_LANES=$L python bench.py --workload config4 --no-cpu-baseline --batch 32768 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('config4 lanes $L', round(d['value']), d['ms_per_step'])"
done
